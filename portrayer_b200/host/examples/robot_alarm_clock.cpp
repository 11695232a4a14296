// examples/robot-alarm-clock.rs — restated object for object: 9 KDMesh instances (4 of them through shared
// nodes), 14 linear Meshes, an AREA light, glossy reflective metal and table (normal-mapped), a dielectric display,
// a uv_trans-tiled wallpaper, instanced sub-trees (eyeballs, connectors, buttons).  1920x1080.
#include "examples.hpp"
using namespace portrayer;

namespace {
Radians deg(double d) { return Radians::from_degrees(d); }
std::shared_ptr<const MeshData> model(const char* name) { return MeshData::load_obj(std::string("assets/robot-alarm-clock/") + name); }
MaterialRef plain(Rgb diffuse) { return Arc(Material{.diffuse = diffuse, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0}); }

SceneNode room() {
    auto wallpaper = ImageTexture::open("assets/robot-alarm-clock/wallpaper.jpg");
    auto mat_wall = Arc(Material{.specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .texture = wallpaper,
                                 .uv_trans = Mat3::scaling_3d({3.0, 3.0, 3.0})});
    auto wood = ImageTexture::open("assets/Wood_018_basecolor_cubemap.jpg");
    auto wood_normals = NormalMap::open("assets/Wood_018_normal_cubemap.jpg");
    auto mat_table = Arc(Material{.specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .reflectivity = 0.2, .glossy_side_length = 2.0,
                                  .texture = wood, .normals = wood_normals});
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Plane{}, mat_wall)).scaled(20.0).rotated_x(deg(90.0)).translated({-2.0, 8.0, -5.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_table)).scaled({20.0, 1.0, 10.0}).translated({-2.0, 0.0, 0.0}).into(),
    });
}

SceneNode clock() {
    auto mat_clock_case = plain({1.0, 1.0, 1.0});
    auto mat_time_bg = Arc(Material{.diffuse = {0.059252, 0.059252, 0.059252}});
    auto mat_time = Arc(Material{.diffuse = {1.0, 0.0, 0.0}});
    auto clock_case_model = model("robot_base_clock_case.obj");
    auto clock_time_model = model("robot_base_clock_time.obj");
    const double angle = -6.62911;
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Mesh(clock_case_model, Shading::Smooth), mat_clock_case))
            .rotated_x(deg(angle)).translated({0.0, 1.228179, 0.350087}).into(),
        SceneNode::from(Geometry(Plane{}, mat_time_bg))
            .scaled({2.966855, 1.0, 0.684205}).rotated_x(deg(90.0 + angle)).translated({0.0, 1.294323, 0.919223}).into(),
        SceneNode::from(Geometry(Mesh(clock_time_model, Shading::Flat), mat_time))
            .rotated_x(deg(83.2518 - 90.0)).translated({0.0, 1.535768, 0.921095}).into(),
    });
}

SceneNode clock_buttons() {
    auto mat_clock_button = plain({0.8, 0.103095, 0.086502});
    auto clock_button_model = model("robot_base_clock_button.obj");
    NodeRef clock_button = SceneNode::from(Geometry(Mesh(clock_button_model, Shading::Smooth), mat_clock_button)).into();
    std::vector<NodeRef> nodes;
    for (double x : {-1.2, -0.4, 0.4, 1.2})
        nodes.push_back(SceneNode::from(clock_button).rotated_x(deg(15.0)).translated({x, 1.7, -0.2}).into());
    return SceneNode::from(std::move(nodes));
}

// `count` copies of one connector KDMesh stacked 0.2 apart from y_offset, at every x of `xs`
SceneNode connectors(const char* obj, const MaterialRef& mat_connector, double y_offset, int count, std::vector<double> xs) {
    const double height = 0.2;
    auto connector_model = model(obj);
    NodeRef connector = SceneNode::from(Geometry(KDMesh(*connector_model, Shading::Flat), mat_connector)).into();
    std::vector<NodeRef> nodes;
    for (double x : xs)
        for (int i = 0; i < count; ++i) {
            const double y = y_offset + (double)i * height;
            nodes.push_back(SceneNode::from(connector).translated({x, y, -0.712655}).into());
        }
    return SceneNode::from(std::move(nodes));
}

SceneNode robot_base(const MaterialRef& mat_robot_metal, const MaterialRef& mat_connector) {
    auto robot_base_model = model("robot_base.obj");
    auto robot_base_sides_model = model("robot_base_sides.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(KDMesh(*robot_base_model, Shading::Smooth), mat_robot_metal)).translated({0.0, 1.002795, -0.209603}).into(),
        SceneNode::from(Geometry(KDMesh(*robot_base_sides_model, Shading::Flat), mat_robot_metal)).translated({0.0, 1.002795, -0.209603}).into(),
        clock().into(),
        clock_buttons().into(),
        connectors("robot_base_connector.obj", mat_connector, 1.960454, 5, {0.0}).into(),
    });
}

SceneNode arm_sockets() {
    auto mat_arm_socket = plain({1.0, 1.0, 1.0});
    auto arm_socket_model = model("robot_arm_socket.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Mesh(arm_socket_model, Shading::Smooth), mat_arm_socket)).translated({2.1, 3.8, -0.7}).into(),
        SceneNode::from(Geometry(Mesh(arm_socket_model, Shading::Smooth), mat_arm_socket))
            .rotated_y(deg(180.0)).translated({-2.1, 3.8, -0.7}).into(),
    });
}

SceneNode arms(const MaterialRef& mat_robot_metal) {
    auto mat_hand = plain({1.0, 1.0, 1.0});
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Mesh(model("robot_arm_left.obj"), Shading::Smooth), mat_robot_metal)).translated({2.1, 3.8, -0.7}).into(),
        SceneNode::from(Geometry(Mesh(model("robot_arm_right.obj"), Shading::Smooth), mat_robot_metal)).translated({-2.1, 3.8, -0.7}).into(),
        SceneNode::from(Geometry(Mesh(model("robot_hand_left.obj"), Shading::Smooth), mat_hand)).translated({2.95, 5.45, -0.7}).into(),
        SceneNode::from(Geometry(Mesh(model("robot_hand_right.obj"), Shading::Smooth), mat_hand)).translated({-2.95, 5.45, -0.7}).into(),
    });
}

SceneNode robot_torso(const MaterialRef& mat_robot_metal, const MaterialRef& mat_connector) {
    auto mat_torso_display = Arc(Material{.diffuse = {0.204899, 0.066919, 0.086002}, .reflectivity = 0.1,
                                          .refraction_index = OPTICAL_GLASS_REFRACTION_INDEX});
    auto mat_torso_text = Arc(Material{.diffuse = {1.0, 0.0, 0.0}});
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(KDMesh(*model("robot_torso.obj"), Shading::Smooth), mat_robot_metal)).translated({0.0, 3.781665, -0.7}).into(),
        SceneNode::from(Geometry(KDMesh(*model("robot_torso_sides.obj"), Shading::Flat), mat_robot_metal)).translated({0.0, 3.781665, -0.7}).into(),
        SceneNode::from(Geometry(Mesh(model("robot_torso_display.obj"), Shading::Smooth), mat_torso_display))
            .translated({0.0, 3.828179, -0.255186}).into(),
        SceneNode::from(Geometry(Mesh(model("robot_torso_text.obj"), Shading::Flat), mat_torso_text))
            .translated({-0.016937, 3.806762, 0.040324}).into(),
        arm_sockets().into(),
        arms(mat_robot_metal).into(),
        connectors("robot_torso_connector.obj", mat_connector, 4.783508, 4, {0.0}).into(),
    });
}

SceneNode robot_head(const MaterialRef& mat_robot_metal, const MaterialRef& mat_connector) {
    auto mat_smile = plain({0.0, 0.0, 0.0});
    auto mat_eyeball = plain({1.0, 1.0, 1.0});
    auto mat_pupil = plain({0.0, 0.0, 0.0});
    NodeRef eyeball = SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Mesh(model("robot_eyeball.obj"), Shading::Smooth), mat_eyeball)).into(),
        SceneNode::from(Geometry(Mesh(model("robot_pupil.obj"), Shading::Smooth), mat_pupil)).into(),
    }).into();
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(KDMesh(*model("robot_head.obj"), Shading::Smooth), mat_robot_metal)).translated({0.0, 5.95, -0.7}).into(),
        SceneNode::from(Geometry(KDMesh(*model("robot_head_sides.obj"), Shading::Flat), mat_robot_metal)).translated({0.0, 5.95, -0.7}).into(),
        SceneNode::from(Geometry(Mesh(model("robot_smile.obj"), Shading::Smooth), mat_smile)).translated({0.0, 6.137964, -0.117689}).into(),
        connectors("robot_head_connector.obj", mat_connector, 6.583508, 3, {-0.6, 0.6}).into(),
        SceneNode::from(eyeball).translated({-0.6, 7.53, -0.7}).into(),
        SceneNode::from(eyeball).translated({0.6, 7.53, -0.7}).into(),
    });
}

// `metal`: the robot's diffuse colour — the example source ships green and keeps the other three as commented-out
// lines (examples/robot-alarm-clock.rs:98-101); upstream published a render of each (render/10_robot-alarm-clock*.png)
SceneNode robot(Rgb metal) {
    auto mat_robot_metal = Arc(Material{.diffuse = metal, .specular = {0.8, 0.8, 0.8}, .shininess = 100.0,
                                        .reflectivity = 0.3, .glossy_side_length = 2.0});
    auto mat_connector = plain({0.048247, 0.048247, 0.048247});
    return SceneNode::from(std::vector<NodeRef>{
        robot_base(mat_robot_metal, mat_connector).into(),
        robot_torso(mat_robot_metal, mat_connector).into(),
        robot_head(mat_robot_metal, mat_connector).into(),
    });
}
}  // namespace

ExampleScene robot_scene(const char* name, Rgb metal) {
    ExampleScene ex;
    ex.name = name;
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{room().into(), robot(metal).into()}).into(),
        .lights = {Light{.position = {-2.0, 15.0, 5.0}, .color = {0.9, 0.9, 0.9},
                         .area = Parallelogram{.a = {5.0, 0.0, 0.0}, .b = {0.0, 0.0, 5.0}}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {1.914036, 3.826548, 20.213762}, .center = {-3.201259, 4.146196, -14.407373}, .up = Vec3::up(),
                            .fovy = deg(23.0)};
    ex.width = 1920;
    ex.height = 1080;
    ex.background = [](Uv uv) { return Rgb{0.529, 0.808, 0.922} * (1.0 - uv.v) + Rgb{0.086, 0.38, 0.745} * uv.v; };
    return ex;
}

PORTRAYER_EXAMPLE(robot_alarm_clock, "robot-alarm-clock") { return robot_scene("robot-alarm-clock", {0.006449, 0.417885, 0.025384}); }  // green, as shipped
// the three other published colour variants (not among the reference's 28 example programs: same program, one line changed)
PORTRAYER_EXAMPLE(robot_alarm_clock_cyan, "robot-alarm-clock-cyan") { return robot_scene("robot-alarm-clock-cyan", {0.211857, 0.772537, 0.8971}); }
PORTRAYER_EXAMPLE(robot_alarm_clock_dark_blue, "robot-alarm-clock-dark-blue") { return robot_scene("robot-alarm-clock-dark-blue", {0.006512, 0.08022, 0.417885}); }
PORTRAYER_EXAMPLE(robot_alarm_clock_red, "robot-alarm-clock-red") { return robot_scene("robot-alarm-clock-red", {0.417885, 0.006501, 0.006501}); }
