// examples/monkeys-making-monkeys.rs — restated object for object: a room of three planes, a textured painting,
// a glossy normal-mapped desk (cube-mapped wood), a cube-mapped CPU tower, 8 linear Meshes (3 uses of monkey.obj:
// flat dielectric "hologram", smooth head; screen, text, torso, teapot), a glass ball, glossy teapot / golf ball,
// TWO AREA lights.  1920x1080.  `assets/cpu_cubemap.png` is missing upstream (.MISSING_LARGE_BLOBS): the
// texture registry hands out its procedural stand-in (portrayer_b200/assets.py).
#include "examples.hpp"
using namespace portrayer;

namespace {
Radians deg(double d) { return Radians::from_degrees(d); }

SceneNode room() {
    auto mat_floor = Arc(Material{.diffuse = {0.655758, 0.8, 0.753899}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto mat_walls = Arc(Material{.diffuse = {0.8, 0.680366, 0.555109}, .specular = {0.8, 0.8, 0.8}, .shininess = 25.0});
    return SceneNode::from(std::vector<NodeRef>{
        // ground, left wall, right wall
        SceneNode::from(Geometry(Plane{}, mat_floor)).scaled(16.0).translated({0.0, 0.0, 3.708507}).into(),
        SceneNode::from(Geometry(Plane{}, mat_walls)).scaled(16.0).rotated_z(deg(-90.0)).translated({-6.340487, 5.0, 4.199467}).into(),
        SceneNode::from(Geometry(Plane{}, mat_walls)).scaled(16.0).rotated_x(deg(90.0)).translated({0.0, 5.0, -3.2}).into(),
    });
}

SceneNode wall_decor() {
    auto mat_poster = Arc(Material{.diffuse = {0.8, 0.329194, 0.120657}, .specular = {0.8, 0.8, 0.8}, .shininess = 25.0});
    auto painting = ImageTexture::open("assets/four-shapes.png");
    auto mat_painting = Arc(Material{.specular = {0.2, 0.2, 0.2}, .shininess = 25.0, .texture = painting});
    auto mat_canvas = Arc(Material{.diffuse = {0.8, 0.8, 0.8}, .specular = {0.2, 0.2, 0.2}, .shininess = 25.0});
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Plane{}, mat_poster)).scaled(4.74905).rotated_z(deg(-90.0)).translated({-6.330487, 8.043096, 3.401992}).into(),
        SceneNode::from(Geometry(Plane{}, mat_painting)).scaled({6.0, 1.0, 1.6}).rotated_x(deg(90.0)).translated({-1.0, 10.2, -3.095}).into(),
        SceneNode::from(Geometry(Cube{}, mat_canvas)).scaled({6.0, 1.6, 0.2}).translated({-1.0, 10.2, -3.2}).into(),
    });
}

SceneNode desk() {
    auto wood = ImageTexture::open("assets/Wood_018_basecolor_cubemap.jpg");
    auto wood_normals = NormalMap::open("assets/Wood_018_normal_cubemap.jpg");
    auto mat_desk = Arc(Material{.specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .reflectivity = 0.2, .glossy_side_length = 2.0,
                                 .texture = wood, .normals = wood_normals});
    std::vector<NodeRef> nodes;
    nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_desk)).scaled({8.0, 0.5, 6.0}).translated({0.0, 5.0, 0.0}).into());
    for (double x : {-3.5, 3.5})
        for (double z : {-2.517656, 2.517656}) {
            const double y = 2.54158;
            nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_desk)).scaled({0.470548, 4.8, 0.470548}).translated({x, y, z}).into());
        }
    return SceneNode::from(std::move(nodes));
}

SceneNode computer(const std::shared_ptr<const MeshData>& monkey_mesh) {
    auto cpu = ImageTexture::open("assets/cpu_cubemap.png");
    auto mat_cpu = Arc(Material{.texture = cpu});
    auto mat_computer = Arc(Material{.diffuse = {0.043232, 0.043232, 0.043232}, .specular = {0.3, 0.3, 0.3}, .shininess = 10.0});
    auto mat_screen = Arc(Material{.diffuse = {0.655925, 0.655925, 0.655925}, .specular = {0.3, 0.3, 0.3}, .shininess = 10.0});
    auto mat_screen_text = Arc(Material{.diffuse = {0.8, 0.8, 0.8}, .specular = {0.3, 0.3, 0.3}, .shininess = 10.0});
    auto mat_hologram = Arc(Material{.diffuse = {0.479036, 0.8, 0.518124}, .reflectivity = 0.6, .refraction_index = WATER_REFRACTION_INDEX});
    auto computer_screen_base_mesh = MeshData::load_obj("assets/computer_screen_base.obj");
    auto computer_edge_display_mesh = MeshData::load_obj("assets/computer_edge_display.obj");
    auto screen_text_mesh = MeshData::load_obj("assets/text_monkey.3d.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_cpu)).scaled({1.6, 3.0, 2.0}).translated({-3.0, 6.74, 0.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_computer)).scaled({0.28, 0.12, 0.4}).translated({1.411292, 5.327119, 1.857835}).into(),
        SceneNode::from(Geometry(Mesh(computer_screen_base_mesh, Shading::Smooth), mat_computer)).translated({0.0, 5.25, 0.0}).into(),
        SceneNode::from(Geometry(Mesh(computer_edge_display_mesh, Shading::Flat), mat_screen)).translated({0.0, 7.256888, 0.0}).into(),
        SceneNode::from(Geometry(Mesh(screen_text_mesh, Shading::Flat), mat_screen_text)).translated({0.0, 9.081371, 0.01}).into(),
        // holographic monkey
        SceneNode::from(Geometry(Mesh(monkey_mesh, Shading::Flat), mat_hologram))
            .scaled(1.5).rotated_xzy(deg(-33.2668), deg(8.17821), deg(-8.17821)).translated({0.0, 7.0, 0.0}).into(),
    });
}

SceneNode chair() {
    auto mat_chair = Arc(Material{.diffuse = {0.032075, 0.032075, 0.032075}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Sphere{}, mat_chair)).scaled({1.283107, 1.537732, 0.425492}).translated({0.0, 5.334378, 5.404959}).into(),
    });
}

SceneNode character(const std::shared_ptr<const MeshData>& monkey_mesh) {
    auto mat_torso = Arc(Material{.diffuse = {0.077701, 0.075793, 0.125964}, .specular = {0.8, 0.8, 0.8}, .shininess = 25.0});
    auto mat_head = Arc(Material{.diffuse = {0.064598, 0.270305, 0.716789}, .specular = {0.8, 0.8, 0.8}, .shininess = 25.0});
    auto monkey_torso_mesh = MeshData::load_obj("assets/monkey_torso.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Mesh(monkey_mesh, Shading::Smooth), mat_head)).rotated_y(deg(180.0)).translated({0.0, 7.0, 4.0}).into(),
        SceneNode::from(Geometry(Mesh(monkey_torso_mesh, Shading::Smooth), mat_torso)).translated({0.0, 5.148612, 4.23546}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_torso))
            .scaled({0.282782, 1.299079, 0.282782}).rotated_z(deg(19.0)).translated({0.984683, 5.126376, 4.344858}).into(),
    });
}

SceneNode desk_objects() {
    auto mat_teapot = Arc(Material{.diffuse = {0.314666, 0.314666, 0.314666}, .specular = {0.8, 0.8, 0.8}, .shininess = 25.0,
                                   .reflectivity = 0.3, .glossy_side_length = 1.0});
    auto mat_glass = Arc(Material{.diffuse = {0.0, 0.0, 0.0}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .reflectivity = 1.0,
                                  .refraction_index = OPTICAL_GLASS_REFRACTION_INDEX});
    auto mat_apple = Arc(Material{.diffuse = {0.8, 0.0, 0.0}});
    auto mat_golf_ball = Arc(Material{.diffuse = {0.8, 0.8, 0.8}, .specular = {0.8, 0.8, 0.8}, .shininess = 25.0,
                                      .reflectivity = 0.3, .glossy_side_length = 1.0});
    auto mat_cone = Arc(Material{.diffuse = {0.368949, 0.335492, 0.8}});
    auto teapot_mesh = MeshData::load_obj("assets/teapot.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Mesh(teapot_mesh, Shading::Smooth), mat_teapot)).scaled(0.030).translated({2.43888, 5.241134, -0.617814}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_glass)).scaled(0.5).translated({2.768083, 5.751237, -1.475317}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_apple)).scaled(0.28).translated({3.369787, 5.538453, -0.782367}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_golf_ball)).scaled(0.14).translated({3.03616, 5.384166, -0.381234}).into(),
        SceneNode::from(Geometry(Cone{}, mat_cone)).scaled({0.64963, 1.106842, 0.64963}).translated({3.182365, 5.777666, -2.332999}).into(),
    });
}
}  // namespace

PORTRAYER_EXAMPLE(monkeys_making_monkeys, "monkeys-making-monkeys") {
    ExampleScene ex;
    ex.name = "monkeys-making-monkeys";
    auto monkey_mesh = MeshData::load_obj("assets/monkey.obj");
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            room().into(), wall_decor().into(), desk().into(), desk_objects().into(), computer(monkey_mesh).into(), chair().into(),
            character(monkey_mesh).into(),
        }).into(),
        .lights = {
            // overhead light, window
            Light{.position = {0.0, 13.0, 1.0}, .color = {0.9, 0.9, 0.9}, .area = Parallelogram{.a = {4.0, 0.0, 0.0}, .b = {0.0, 0.0, 4.0}}},
            Light{.position = {8.0, 8.0, 8.0}, .color = {0.4, 0.4, 0.4}, .area = Parallelogram{.a = {0.0, 0.0, 2.5}, .b = {0.0, 2.5, 0.0}}},
        },
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {10.626843, 11.525522, 15.875655}, .center = {-11.287256, 4.506533, -10.496798}, .up = Vec3::up(),
                            .fovy = deg(23.0)};
    ex.width = 1920;
    ex.height = 1080;
    ex.background = sky_gradient;
    return ex;
}
