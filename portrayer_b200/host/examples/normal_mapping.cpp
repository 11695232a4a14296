// examples/normal-mapping.rs — BASELINE.json configs[1]; three images that
// differ only in the light position (normal-mapping.rs:139-143).
#include "examples.hpp"
using namespace portrayer;

static ExampleScene build_normal_mapping(const std::string& name, Vec3 light_pos) {
    auto tex_map_plane = ImageTexture::open("assets/Terracotta_Tiles_002_Base_Color.jpg");
    auto norm_map_plane = NormalMap::open("assets/Terracotta_Tiles_002_Normal.jpg");
    auto mat_tex_plane = Arc(Material{.diffuse = {0.37168, 0.236767, 0.692066}, .specular = {0.4, 0.4, 0.4},
                                      .shininess = 25.0, .texture = tex_map_plane});
    auto mat_tex_plane_norm = Arc(Material{.diffuse = {0.37168, 0.236767, 0.692066}, .specular = {0.4, 0.4, 0.4},
                                           .shininess = 25.0, .texture = tex_map_plane, .normals = norm_map_plane});

    auto tex_map_sphere = ImageTexture::open("assets/Rock_033_baseColor_2.jpg");
    auto norm_map_sphere = NormalMap::open("assets/Rock_033_normal_2.jpg");
    auto mat_tex_sphere = Arc(Material{.diffuse = {0.37168, 0.236767, 0.692066}, .specular = {0.6, 0.6, 0.6},
                                       .shininess = 25.0, .texture = tex_map_sphere});
    auto mat_tex_sphere_norm = Arc(Material{.diffuse = {0.37168, 0.236767, 0.692066}, .specular = {0.6, 0.6, 0.6},
                                            .shininess = 25.0, .texture = tex_map_sphere, .normals = norm_map_sphere});

    auto tex_map_cube = ImageTexture::open("assets/Stone_Wall_007_COLOR_cubemap.jpg");
    auto norm_map_cube = NormalMap::open("assets/Stone_Wall_007_NORM_cubemap.jpg");
    auto mat_tex_cube = Arc(Material{.diffuse = {0.37168, 0.236767, 0.692066}, .specular = {0.3, 0.3, 0.3},
                                     .shininess = 25.0, .texture = tex_map_cube});
    auto mat_tex_cube_norm = Arc(Material{.diffuse = {0.37168, 0.236767, 0.692066}, .specular = {0.3, 0.3, 0.3},
                                          .shininess = 25.0, .texture = tex_map_cube, .normals = norm_map_cube});

    auto mat_wall_floor = Arc(Material{.diffuse = {0.424858, 0.531206, 0.8}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});

    NodeRef scene_root = SceneNode::from(std::vector<NodeRef>{
        // Floor
        SceneNode::from(Geometry(Plane{}, mat_wall_floor)).scaled(40.0).translated({0.0, -1.0, 0.0}).into(),
        // Left - Texture Only
        SceneNode::from(Geometry(Plane{}, mat_tex_plane))
            .scaled(6.0).rotated_x(Radians::from_degrees(90.0)).translated({-4.0, 2.0, -6.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_tex_cube)).scaled(2.0).translated({-7.0, 0.0, -1.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_tex_sphere)).translated({-7.0, 2.0, -1.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_tex_cube)).scaled(2.0).translated({-2.0, 0.0, 3.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_tex_sphere)).translated({-2.0, 2.0, 3.0}).into(),
        // Right - Normal + Texture
        SceneNode::from(Geometry(Plane{}, mat_tex_plane_norm))
            .scaled(6.0).rotated_x(Radians::from_degrees(90.0)).translated({4.0, 2.0, -6.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_tex_cube_norm)).scaled(2.0).translated({7.0, 0.0, -1.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_tex_sphere_norm)).translated({7.0, 2.0, -1.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_tex_cube_norm)).scaled(2.0).translated({2.0, 0.0, 3.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_tex_sphere_norm)).translated({2.0, 2.0, 3.0}).into(),
    }).into();

    ExampleScene ex;
    ex.name = name;
    ex.scene = HierScene{
        .root = scene_root,
        .lights = {Light{.position = light_pos, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.2, 0.2, 0.2},
    };
    ex.cam = CameraSettings{.eye = {0.0, 8.07551, 23.078941}, .center = {0.0, -2.854475, -16.437334},
                            .up = Vec3::up(), .fovy = Radians::from_degrees(22.0)};
    ex.width = 910;
    ex.height = 512;
    ex.background = sky_gradient;
    return ex;
}

PORTRAYER_EXAMPLE(normal_mapping, "normal-mapping") { return build_normal_mapping("normal-mapping", {0.0, 8.0, 10.0}); }
PORTRAYER_EXAMPLE(normal_mapping_left, "normal-mapping-left") {
    return build_normal_mapping("normal-mapping-left", {-8.0, 8.0, 10.0});
}
PORTRAYER_EXAMPLE(normal_mapping_right, "normal-mapping-right") {
    return build_normal_mapping("normal-mapping-right", {8.0, 8.0, 10.0});
}
