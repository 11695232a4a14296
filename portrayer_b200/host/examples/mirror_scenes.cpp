// examples/cube-mapping.rs, examples/entering-the-mirror-dimension.rs, examples/transmission-refraction.rs —
// restated object for object.  Mirrors (reflectivity 1.0 / 0.9), with_child / with_children hierarchies, rotated_xzy
// with three different angles, textured KDMesh instances inside a dielectric water cube behind a dielectric glass
// pane, normal-mapped cubes.  (assets/earth_cube.png is missing upstream: the stand-in of texture-mapping is used.)
#include "examples.hpp"
using namespace portrayer;

namespace {
MaterialRef phong(Rgb diffuse, Rgb specular, double shininess, double reflectivity = 0.0) {
    return Arc(Material{.diffuse = diffuse, .specular = specular, .shininess = shininess, .reflectivity = reflectivity});
}
Radians deg(double d) { return Radians::from_degrees(d); }
ExampleScene finish(const char* name, HierScene scene, CameraSettings cam, size_t w, size_t h) {
    ExampleScene ex;
    ex.name = name;
    ex.scene = std::move(scene);
    ex.cam = cam;
    ex.width = w;
    ex.height = h;
    ex.background = sky_gradient;
    return ex;
}
}  // namespace

// examples/cube-mapping.rs
PORTRAYER_EXAMPLE(cube_mapping, "cube-mapping") {
    auto mat_mirror = phong({0.0, 0.0, 0.0}, {0.6, 0.6, 0.6}, 1000.0, 1.0);
    auto mat_wood = phong({0.545, 0.353, 0.169}, {0.5, 0.7, 0.5}, 25.0);
    auto earth = ImageTexture::open("assets/earth.jpg");
    auto mat_tex = Arc(Material{.diffuse = {0.506, 0.78, 0.518}, .specular = {0.5, 0.5, 0.5}, .shininess = 25.0, .texture = earth});
    auto earth_cubemap = ImageTexture::open("assets/earth_cube.png");
    auto mat_tex_cube = Arc(Material{.diffuse = {0.506, 0.78, 0.518}, .specular = {0.5, 0.5, 0.5}, .shininess = 25.0,
                                     .texture = earth_cubemap});
    NodeRef mirror = SceneNode::from(Geometry(Cube{}, mat_wood))
        .scaled({9.0, 0.5, 6.0})
        .rotated_x(deg(10.0))
        .with_child(SceneNode::from(Geometry(Cube{}, mat_mirror))
                        .scaled({8.1 / 9.0, 0.05 / 0.5, 5.4 / 6.0}).translated({0.0, 0.27 / 0.5, 0.0}).into())
        .into();
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            mirror,
            SceneNode::from(Geometry(Plane{}, mat_tex)).scaled({8.0, 1.0, 2.0}).rotated_x(deg(90.0)).translated({0.0, 2.0, -2.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat_tex_cube)).scaled(1.5).translated({-3.75, 2.0, 0.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat_tex_cube)).scaled(1.5).rotated_y(deg(-90.0)).translated({-1.25, 2.0, 0.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat_tex_cube)).scaled(1.5).rotated_y(deg(180.0)).translated({1.25, 2.0, 0.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat_tex_cube)).scaled(1.5).rotated_y(deg(-270.0)).translated({3.75, 2.0, 0.0}).into(),
        }).into(),
        .lights = {Light{.position = {-6.0, 5.0, 4.0}, .color = {0.5, 0.5, 0.5}},
                   Light{.position = {6.0, 5.0, 4.0}, .color = {0.5, 0.5, 0.5}},
                   Light{.position = {0.0, 1.0, -4.0}, .color = {0.5, 0.5, 0.5}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("cube-mapping", std::move(scene),
                  CameraSettings{.eye = {0.0, 10.15667, 11.579666}, .center = {0.0, -5.913023, -7.571445}, .up = Vec3::up(),
                                 .fovy = deg(25.0)}, 910, 512);
}

// examples/entering-the-mirror-dimension.rs
PORTRAYER_EXAMPLE(mirror_dimension, "entering-the-mirror-dimension") {
    auto mat_mirror_frame = phong({0.29, 0.204, 0.145}, {0.0, 0.0, 0.0}, 1.0);
    auto mat_mirror = phong({0.0, 0.0, 0.0}, {0.8, 0.8, 0.8}, 1000.0, 1.0);
    auto mat_floor = phong({0.016, 0.384, 0.0}, {0.8, 0.8, 0.8}, 25.0);
    auto mat_body = phong({0.906, 0.22, 0.282}, {0.8, 0.8, 0.8}, 25.0);
    auto mat_head = phong({0.086, 0.671, 0.906}, {0.8, 0.8, 0.8}, 50.0);
    auto mat_eyes = phong({0.3, 0.3, 0.3}, {0.8, 0.8, 0.8}, 1000.0, 0.9);
    auto mat_arms = phong({0.345, 0.588, 0.906}, {0.8, 0.8, 0.8}, 1.0);
    auto monkey = MeshData::load_obj("assets/monkey.obj");
    auto plane = MeshData::load_obj("assets/plane.obj");

    NodeRef mirror = SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_mirror_frame)).scaled({3.96, 5.5, 0.4}).translated({0.0, 2.75, 0.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_mirror)).scaled({3.6, 5.0, 0.1}).translated({0.0, 2.75, 0.2}).into(),
    }).translated({0.0, 0.0, -1.3}).into();

    auto arm = [&](Vec3 scale, Vec3 angles, Vec3 at) {
        return SceneNode::from(Geometry(Sphere{}, mat_arms)).scaled(scale)
            .rotated_xzy(deg(angles.x), deg(angles.y), deg(angles.z)).translated(at).into();
    };
    NodeRef monkey_character = SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_body)).scaled({0.545055, 2.6, 0.545055}).translated({0.0, 1.3, 0.0}).into(),
        SceneNode::from(Geometry(Mesh(monkey, Shading::Flat), mat_head))
            .scaled({1.0, 1.0, 1.0}).rotated_y(deg(180.0)).translated({0.0, 2.7, 0.0})
            .with_children({
                SceneNode::from(Geometry(Sphere{}, mat_eyes)).scaled({0.1, 0.1, 0.05}).translated({0.35, 0.24, 0.8}).into(),
                SceneNode::from(Geometry(Sphere{}, mat_eyes)).scaled({0.1, 0.1, 0.05}).translated({-0.35, 0.24, 0.8}).into(),
            })
            .into(),
        arm({0.2, 0.63, 0.2}, {161.156, 107.062, -133.944}, {-0.388703, 1.715599, -0.2}),
        arm({0.2, 0.56, 0.2}, {127.221, 42.0695, -104.823}, {-0.711297, 1.284401, -1.0}),
        SceneNode::from(Geometry(Sphere{}, mat_mirror)).scaled({0.5, 0.5, 0.3}).translated({-0.711297, 1.284401, -1.20}).into(),
        arm({0.2, 0.63, 0.2}, {92.3684, -57.6199, 38.2278}, {0.581161, 1.984976, -0.2}),
        arm({0.2, 0.56, 0.2}, {91.5166, -11.239, 28.419}, {1.118839, 2.015024, -1.0}),
        SceneNode::from(Geometry(Sphere{}, mat_mirror)).scaled({0.5, 0.5, 0.3}).translated({1.118839, 2.015024, -1.20}).into(),
    }).into();
    NodeRef floor = SceneNode::from(Geometry(Mesh(plane, Shading::Flat), mat_floor)).scaled(20.0).into();

    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{mirror, floor, monkey_character}).into(),
        .lights = {Light{.position = {2.5, 3.5, -1.0}, .color = {0.9, 0.9, 0.9}},
                   Light{.position = {10.0, 10.0, 0.0}, .color = {0.9, 0.9, 0.9}},
                   Light{.position = {-9.0, 4.0, 0.0}, .color = {0.406471, 0.901283, 1.0}}},
        .ambient = {0.2, 0.2, 0.2},
    };
    return finish("entering-the-mirror-dimension", std::move(scene),
                  CameraSettings{.eye = {5.545485, 2.966984, 1.795613}, .center = {-4.348584, 2.148794, -3.057839}, .up = Vec3::up(),
                                 .fovy = deg(30.0)}, 800, 600);
}

// examples/transmission-refraction.rs
namespace {
SceneNode tr_room() {
    auto mat_walls = phong({0.607917, 0.8, 0.551884}, {0.3, 0.3, 0.3}, 25.0);
    auto wood = ImageTexture::open("assets/Wood_018_basecolor_cubemap.jpg");
    auto wood_normals = NormalMap::open("assets/Wood_018_normal_cubemap.jpg");
    auto mat_table = Arc(Material{.specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .texture = wood, .normals = wood_normals});
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_table)).scaled({20.0, 5.0, 2.5}).translated({0.0, -2.0, 1.3}).into(),
        SceneNode::from(Geometry(Plane{}, mat_walls)).scaled({20.0, 1.0, 20.0}).rotated_x(deg(90.0)).translated({0.0, 3.0, -10.0}).into(),
        SceneNode::from(Geometry(Plane{}, mat_walls)).scaled({20.0, 1.0, 12.0}).rotated_z(deg(90.0)).translated({10.0, 3.0, -6.0}).into(),
        SceneNode::from(Geometry(Plane{}, mat_walls)).scaled({20.0, 1.0, 12.0}).rotated_z(deg(-90.0)).translated({-10.0, 3.0, -6.0}).into(),
        SceneNode::from(Geometry(Plane{}, mat_walls)).scaled({12.1, 1.0, 20.0}).rotated_x(deg(90.0)).translated({16.0, 3.0, 0.0}).into(),
        SceneNode::from(Geometry(Plane{}, mat_walls)).scaled({12.1, 1.0, 20.0}).rotated_x(deg(90.0)).translated({-16.0, 3.0, 0.0}).into(),
    });
}
SceneNode tr_tank() {
    auto tiles = ImageTexture::open("assets/Tiles_017_basecolor_cubemap.jpg");
    auto tiles_normals = NormalMap::open("assets/Tiles_017_normal_cubemap.jpg");
    auto mat_tank = Arc(Material{.specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .texture = tiles, .normals = tiles_normals});
    std::vector<NodeRef> nodes;
    for (int i = 0; i < 4; ++i) {
        nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_tank)).scaled({5.0, 5.0, 0.2})
                            .translated({(double)i * 5.0 - 7.5, -2.0, -10.0}).into());
        nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_tank)).scaled({5.0, 5.0, 0.2})
                            .translated({(double)i * 5.0 - 7.5, -2.0, 0.0}).into());
    }
    for (int i = 0; i < 2; ++i) {
        nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_tank)).scaled({0.2, 5.0, 5.0})
                            .translated({-10.0, -2.0, -((double)i * 5.0 + 2.5)}).into());
        nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_tank)).scaled({0.2, 5.0, 5.0})
                            .translated({10.0, -2.0, -((double)i * 5.0 + 2.5)}).into());
    }
    for (int x = 0; x < 4; ++x)
        for (int y = 0; y < 2; ++y)
            nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_tank)).scaled({5.0, 0.2, 5.0})
                                .translated({(double)x * 5.0 - 7.5, -4.0, -((double)y * 5.0 + 2.5)}).into());
    return SceneNode::from(std::move(nodes));
}
SceneNode tr_water() {
    auto mat_water = Arc(Material{.diffuse = {0.0, 0.0, 0.1}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .reflectivity = 0.9,
                                  .refraction_index = WATER_REFRACTION_INDEX});
    auto fish_skin = ImageTexture::open("assets/fish.png");
    auto mat_fish = Arc(Material{.diffuse = {0.8, 0.8, 0.8}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .texture = fish_skin});
    auto fish_model = MeshData::load_obj("assets/fish.obj");
    KDMesh fish_mesh(*fish_model, Shading::Smooth);
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_water)).scaled({19.799999, 3.8, 9.8}).translated({0.0, -2.0, -5.0}).into(),
        SceneNode::from(Geometry(fish_mesh, mat_fish))
            .rotated_xzy(deg(0.0), deg(-71.8181), deg(30.8927)).translated({-4.798946, -0.970323, -5.246493}).into(),
        SceneNode::from(Geometry(fish_mesh, mat_fish))
            .rotated_xzy(deg(0.0), deg(108.666), deg(-23.084)).translated({3.110451, -2.562474, -6.838645}).into(),
    });
}
SceneNode tr_drink() {
    auto mat_water = Arc(Material{.diffuse = {0.0, 0.0, 0.1}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .reflectivity = 0.9,
                                  .refraction_index = WATER_REFRACTION_INDEX});
    auto mat_straw = phong({0.8, 0.0, 0.0}, {0.3, 0.3, 0.3}, 25.0);
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cylinder{}, mat_water)).scaled({1.0, 1.4, 1.0}).translated({-7.4, 1.2, 1.2}).into(),
        SceneNode::from(Geometry(Cylinder{}, mat_straw)).scaled({0.1, 2.0, 0.1}).rotated_z(deg(28.4282))
            .translated({-7.565556, 1.411109, 1.1}).into(),
    });
}
}  // namespace

PORTRAYER_EXAMPLE(transmission_refraction, "transmission-refraction") {
    auto mat_glass = Arc(Material{.diffuse = {0.0, 0.0, 0.0}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .reflectivity = 1.0,
                                  .refraction_index = WINDOW_GLASS_REFRACTION_INDEX});
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Cube{}, mat_glass)).scaled({20.0, 10.0, 0.2}).translated({0.0, 5.0, 0.0}).into(),
            tr_room().into(), tr_tank().into(), tr_water().into(), tr_drink().into(),
        }).into(),
        .lights = {Light{.position = {0.0, 27.0, 5.0}, .color = {0.5, 0.5, 0.5}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("transmission-refraction", std::move(scene),
                  CameraSettings{.eye = {0.0, 14.658033, 27.19817}, .center = {0.0, -6.058867, -24.828854}, .up = Vec3::up(),
                                 .fovy = deg(23.0)}, 910, 512);
}
