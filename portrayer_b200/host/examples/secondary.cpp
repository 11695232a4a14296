// BASELINE.json configs[3]: the secondary-ray-bound scenes.
//   examples/water-glass.rs, examples/glossy-reflection.rs, examples/soft-shadows.rs
#include "examples.hpp"
using namespace portrayer;

static SceneNode room() {
    auto brick = ImageTexture::open("assets/Brick_Wall_013_COLOR.jpg");
    auto brick_normals = NormalMap::open("assets/Brick_Wall_013_NORM.jpg");
    // diffuse comes from texture
    auto mat_wall = Arc(Material{.specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .texture = brick, .normals = brick_normals});

    auto wood = ImageTexture::open("assets/Wood_018_basecolor_cubemap.jpg");
    auto wood_normals = NormalMap::open("assets/Wood_018_normal_cubemap.jpg");
    auto mat_table = Arc(Material{.specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .reflectivity = 0.2,
                                  .glossy_side_length = 2.0, .texture = wood, .normals = wood_normals});

    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Plane{}, mat_wall))
            .scaled(10.0).rotated_x(Radians::from_degrees(90.0)).translated({0.0, 1.0, -2.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_table)).scaled({8.0, 0.4, 4.0}).translated({0.0, 0.0, -0.2}).into(),
    });
}

static SceneNode drink() {
    auto mat_water = Arc(Material{.diffuse = {0.0, 0.0, 0.1}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0,
                                  .reflectivity = 0.9, .refraction_index = WATER_REFRACTION_INDEX});
    auto mat_straw = Arc(Material{.diffuse = {0.8, 0.0, 0.0}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});

    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cylinder{}, mat_water)).scaled({1.0, 1.4, 1.0}).translated({0.0, 0.7, 0.0}).into(),
        SceneNode::from(Geometry(Cylinder{}, mat_straw))
            .scaled({0.1, 2.0, 0.1}).rotated_z(Radians::from_degrees(28.4282)).translated({-0.165556, 0.911109, 0.1}).into(),
    });
}

PORTRAYER_EXAMPLE(water_glass, "water-glass") {
    ExampleScene ex;
    ex.name = "water-glass";
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            room().into(),
            drink().translated({0.0, 0.2, 0.0}).into(),
        }).into(),
        .lights = {Light{.position = {0.0, 27.0, 5.0}, .color = {0.5, 0.5, 0.5}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 3.2, 7.151111}, .center = {0.0, 0.091525, -5.719519}, .up = Vec3::up(),
                            .fovy = Radians::from_degrees(23.0)};
    ex.width = 910;
    ex.height = 512;
    ex.background = sky_gradient;
    return ex;
}

PORTRAYER_EXAMPLE(glossy_reflection, "glossy-reflection") {
    Material non_glossy{.diffuse = {0.146505, 0.314666, 0.170564}, .specular = {0.3, 0.3, 0.3}, .shininess = 100.0,
                        .reflectivity = 0.4};
    auto non_glossy_ball = Arc(non_glossy);
    Material glossy = non_glossy;  // ..(*non_glossy_ball).clone()
    glossy.glossy_side_length = 2.0;
    auto glossy_ball = Arc(glossy);
    auto center_ball = Arc(Material{.diffuse = {0.8, 0.0, 0.023362}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto table = Arc(Material{.diffuse = {1.0, 0.6, 0.1}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});

    ExampleScene ex;
    ex.name = "glossy-reflection";
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Sphere{}, non_glossy_ball)).translated({-1.1, 1.3, 0.0}).into(),
            SceneNode::from(Geometry(Sphere{}, glossy_ball)).translated({1.1, 1.3, 0.0}).into(),
            SceneNode::from(Geometry(Sphere{}, center_ball)).scaled(0.5).translated({0.0, 0.8, 1.8}).into(),
            SceneNode::from(Geometry(Cube{}, table)).scaled({10.0, 0.6, 5.0}).into(),
        }).into(),
        .lights = {
            Light{.position = {0.0, 6.0, 3.0}, .color = {0.9, 0.9, 0.9}},
            Light{.position = {0.0, 1.0, 12.0}, .color = {0.7, 0.7, 0.7}},
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 2.562834, 8.863271}, .center = {0.0, -1.083779, -11.817695},
                            .up = Vec3::up(), .fovy = Radians::from_degrees(20.0)};
    ex.width = 910;
    ex.height = 512;
    ex.background = sky_gradient;
    return ex;
}

PORTRAYER_EXAMPLE(soft_shadows, "soft-shadows") {
    auto mat_cow = Arc(Material{.diffuse = {0.37168, 0.236767, 0.692066}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto mat_wall_floor = Arc(Material{.diffuse = {0.627459, 0.8, 0.589836}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});

    auto cow_mesh = MeshData::load_obj("assets/cow.obj");

    ExampleScene ex;
    ex.name = "soft-shadows";
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            // Walls + Floor
            SceneNode::from(Geometry(Plane{}, mat_wall_floor)).scaled(30.0).into(),
            SceneNode::from(Geometry(Cube{}, mat_wall_floor)).scaled({0.2, 20.0, 20.0}).translated({0.0, 8.0, 8.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat_wall_floor)).scaled({30.0, 30.0, 0.4}).translated({0.0, 8.0, -2.0}).into(),
            // Objects
            SceneNode::from(Geometry(Mesh(cow_mesh, Shading::Smooth), mat_cow))
                .scaled(0.5).rotated_y(Radians::from_degrees(-15.0)).translated({-4.2, 1.8, 4.0}).into(),
            SceneNode::from(Geometry(Mesh(cow_mesh, Shading::Smooth), mat_cow))
                .scaled(0.5).rotated_y(Radians::from_degrees(195.0)).translated({4.2, 1.8, 4.0}).into(),
        }).into(),
        .lights = {
            // Left - Point Light
            Light{.position = {-2.0, 2.0, 16.0}, .color = {0.5, 0.5, 0.5}},
            // Right - Area Light
            Light{.position = {2.0, 2.0, 16.0}, .color = {0.5, 0.5, 0.5},
                  .area = Parallelogram{.a = {0.0, 0.5, 0.0}, .b = {0.5, 0.0, 0.0}}},
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 5.04746, 24.827951}, .center = {0.012231, -0.459716, -15.800501},
                            .up = Vec3::up(), .fovy = Radians::from_degrees(25.0)};
    ex.width = 910;
    ex.height = 512;
    ex.background = sky_gradient;
    return ex;
}
