// The small example programs of the reference, restated object for object (same transforms in the same order,
// same cameras, image sizes and background closures):
//   examples/simple.rs, hier.rs, instance.rs, nonhier2.rs, macho-cows.rs, simple-cows.rs, single-triangle.rs,
//   smooth-shading.rs, primitives-simple.rs, four-shapes.rs, antialiasing.rs, fish.rs, graphics-poster.rs
// They widen the parity suite beyond BASELINE.json's configs: deep hierarchies, instancing through shared
// Arc<SceneNode>, a Triangle as a top-level primitive, flat + smooth meshes side by side, a textured smooth mesh,
// a dielectric + glossy mesh, constant backgrounds.
#include "examples.hpp"
using namespace portrayer;

namespace {

MaterialRef phong(Rgb diffuse, Rgb specular, double shininess) {
    return Arc(Material{.diffuse = diffuse, .specular = specular, .shininess = shininess});
}
Rgb white_background(Uv) { return Rgb::white(); }
CameraSettings camera(Vec3 eye, Vec3 center, double fovy_degrees) {
    return CameraSettings{.eye = eye, .center = center, .up = Vec3::up(), .fovy = Radians::from_degrees(fovy_degrees)};
}
ExampleScene finish(const char* name, HierScene scene, CameraSettings cam, size_t w, size_t h,
                    std::function<Rgb(Uv)> background = sky_gradient) {
    ExampleScene ex;
    ex.name = name;
    ex.scene = std::move(scene);
    ex.cam = cam;
    ex.width = w;
    ex.height = h;
    ex.background = std::move(background);
    return ex;
}

// the stone arc of hier.rs / instance.rs / macho-cows.rs (examples/instance.rs:42-57)
SceneNode arc_of(const MaterialRef& mat) {
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat)).scaled({0.8, 4.0, 0.8}).translated({-2.0, 2.0, 0.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat)).scaled({0.8, 4.0, 0.8}).translated({2.0, 2.0, 0.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat)).scaled({4.0, 0.6, 0.6}).translated({0.0, 4.0, 0.0}).into(),
    });
}

}  // namespace

// examples/simple.rs
PORTRAYER_EXAMPLE(simple, "simple") {
    auto mat1 = phong({0.7, 1.0, 0.7}, {0.5, 0.7, 0.5}, 25.0);
    auto mat2 = phong({0.5, 0.5, 0.5}, {0.5, 0.7, 0.5}, 25.0);
    auto mat3 = phong({1.0, 0.6, 0.1}, {0.5, 0.7, 0.5}, 25.0);
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(100.0).translated({0.0, 0.0, -400.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(150.0).translated({200.0, 50.0, -100.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat2)).scaled(1000.0).translated({0.0, -1200.0, -500.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat3)).scaled(50.0).translated({-100.0, 25.0, -300.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(25.0).translated({0.0, 100.0, -250.0}).into(),
        }).into(),
        .lights = {Light{.position = {-100.0, 150.0, 400.0}, .color = {0.9, 0.9, 0.9}},
                   Light{.position = {400.0, 100.0, 150.0}, .color = {0.7, 0.0, 0.7}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("simple", std::move(scene), camera({0.0, 0.0, 800.0}, {0.0, 0.0, 0.0}, 50.0), 256, 256);
}

// examples/hier.rs
PORTRAYER_EXAMPLE(hier, "hier") {
    auto gold = phong({0.9, 0.8, 0.4}, {0.8, 0.8, 0.4}, 25.0);
    auto grass = phong({0.1, 0.7, 0.1}, {0.0, 0.0, 0.0}, 0.0);
    auto blue = phong({0.7, 0.6, 1.0}, {0.5, 0.4, 0.8}, 25.0);
    auto plane = MeshData::load_obj("assets/plane.obj");
    auto dodeca = MeshData::load_obj("assets/dodeca.obj");

    NodeRef arc = arc_of(gold).translated({0.0, 0.0, -10.0}).rotated_y(Radians::from_degrees(60.0)).into();
    NodeRef floor = SceneNode::from(Geometry(Mesh(plane, Shading::Flat), grass)).scaled(30.0).into();
    NodeRef poly = SceneNode::from(Geometry(Mesh(dodeca, Shading::Flat), blue)).translated({-2.0, 1.618034, 0.0}).into();

    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{arc, floor, poly})
                    .rotated_x(Radians::from_degrees(23.0)).translated({6.0, -2.0, -15.0}).into(),
        .lights = {Light{.position = {200.0, 200.0, 400.0}, .color = {0.8, 0.8, 0.8}},
                   Light{.position = {0.0, 5.0, -20.0}, .color = {0.4, 0.4, 0.8}}},
        .ambient = {0.4, 0.4, 0.4},
    };
    return finish("hier", std::move(scene), camera({0.0, 0.0, 0.0}, {0.0, 0.0, -1.0}, 50.0), 256, 256);
}

// examples/instance.rs — one Arc<SceneNode> instanced six times
PORTRAYER_EXAMPLE(instance, "instance") {
    auto stone = phong({0.8, 0.7, 0.7}, {0.0, 0.0, 0.0}, 0.0);
    auto grass = phong({0.1, 0.7, 0.1}, {0.0, 0.0, 0.0}, 0.0);
    auto plane = MeshData::load_obj("assets/plane.obj");

    NodeRef arc = arc_of(stone).translated({0.0, 0.0, -10.0}).into();
    std::vector<NodeRef> nodes;
    for (int i = 1; i <= 6; ++i)
        nodes.push_back(SceneNode::from(arc).rotated_y(Radians::from_degrees(60.0 * (double)i)).into());
    nodes.push_back(SceneNode::from(Geometry(Mesh(plane, Shading::Flat), grass)).scaled(30.0).into());
    nodes.push_back(SceneNode::from(Geometry(Sphere{}, stone)).scaled(2.5).into());

    HierScene scene{
        .root = SceneNode::from(std::move(nodes)).rotated_x(Radians::from_degrees(23.0)).into(),
        .lights = {Light{.position = {200.0, 202.0, 430.0}, .color = {0.8, 0.8, 0.8}}},
        .ambient = {0.4, 0.4, 0.4},
    };
    return finish("instance", std::move(scene), camera({0.0, 2.0, 30.0}, {0.0, 2.0, 29.0}, 50.0), 256, 256);
}

// examples/nonhier2.rs — nonhier with the whole scene pushed back through the root's transform
PORTRAYER_EXAMPLE(nonhier2, "nonhier2") {
    auto mat1 = phong({0.7, 1.0, 0.7}, {0.5, 0.7, 0.5}, 25.0);
    auto mat2 = phong({0.5, 0.5, 0.5}, {0.5, 0.7, 0.5}, 25.0);
    auto mat3 = phong({1.0, 0.6, 0.1}, {0.5, 0.7, 0.5}, 25.0);
    auto mat4 = phong({0.7, 0.6, 1.0}, {0.5, 0.4, 0.8}, 25.0);
    auto monkey = MeshData::load_obj("assets/monkey.obj");
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(100.0).translated({0.0, 0.0, -400.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(150.0).translated({200.0, 50.0, -100.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat2)).scaled(1000.0).translated({0.0, -1200.0, -500.0}).into(),
            SceneNode::from(Geometry(Cube{}, mat4)).scaled(100.0).translated({-150.0, -75.0, 50.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat3)).scaled(50.0).translated({-100.0, 25.0, -300.0}).into(),
            SceneNode::from(Geometry(Sphere{}, mat1)).scaled(25.0).translated({0.0, 100.0, -250.0}).into(),
            SceneNode::from(Geometry(Mesh(monkey, Shading::Flat), mat3)).scaled(100.0).translated({-150.0, 200.0, -100.0}).into(),
        }).translated({0.0, 0.0, -800.0}).into(),
        .lights = {Light{.position = {-100.0, 150.0, -400.0}, .color = {0.9, 0.9, 0.9}},
                   Light{.position = {400.0, 100.0, -650.0}, .color = {0.7, 0.0, 0.7}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("nonhier2", std::move(scene), camera({0.0, 0.0, 0.0}, {0.0, 0.0, -1.0}, 50.0), 256, 256);
}

namespace {
// the ring of arcs, the three cows, the floor and the altar shared by macho-cows.rs and simple-cows.rs
HierScene cow_scene(const MaterialRef& stone, const MaterialRef& grass, NodeRef arc, NodeRef cow) {
    auto plane = MeshData::load_obj("assets/plane.obj");
    auto buckyball = MeshData::load_obj("assets/buckyball.obj");
    std::vector<NodeRef> nodes;
    for (int i = 1; i <= 6; ++i)
        nodes.push_back(SceneNode::from(arc).rotated_y(Radians::from_degrees(60.0 * (double)(i - 1))).into());
    const std::pair<Vec3, double> cows[] = {{{1.0, 1.3, 14.0}, 20.0}, {{5.0, 1.3, -11.0}, 180.0}, {{-5.5, 1.3, -3.0}, -60.0}};
    for (const auto& [cow_pos, cow_rot] : cows)
        nodes.push_back(SceneNode::from(cow).scaled(1.4).rotated_y(Radians::from_degrees(cow_rot)).translated(cow_pos).into());
    nodes.push_back(SceneNode::from(Geometry(Mesh(plane, Shading::Flat), grass)).scaled(30.0).into());
    nodes.push_back(SceneNode::from(Geometry(Mesh(buckyball, Shading::Flat), stone)).scaled(1.5).into());
    return HierScene{
        .root = SceneNode::from(std::move(nodes)).rotated_x(Radians::from_degrees(23.0)).into(),
        .lights = {Light{.position = {200.0, 202.0, 430.0}, .color = {0.8, 0.8, 0.8}}},
        .ambient = {0.4, 0.4, 0.4},
    };
}
}  // namespace

// examples/macho-cows.rs — the 5 804-triangle cow as a linear Mesh, instanced three times
PORTRAYER_EXAMPLE(macho_cows, "macho-cows") {
    auto stone = phong({0.8, 0.7, 0.7}, {0.0, 0.0, 0.0}, 0.0);
    auto grass = phong({0.1, 0.7, 0.1}, {0.0, 0.0, 0.0}, 0.0);
    auto cow_hide = phong({0.84, 0.6, 0.53}, {0.3, 0.3, 0.3}, 20.0);
    auto cow_model = MeshData::load_obj("assets/cow.obj");
    NodeRef arc = arc_of(stone).translated({0.0, 0.0, -10.0}).into();
    NodeRef cow = SceneNode::from(Geometry(Mesh(cow_model, Shading::Flat), cow_hide))
                      .translated({0.0, 3.637, 0.0}).scaled(2.0 / (2.76 + 3.637)).translated({0.0, -1.0, 0.0}).into();
    return finish("macho-cows", cow_scene(stone, grass, arc, cow), camera({0.0, 2.0, 30.0}, {0.0, 2.0, 29.0}, 50.0), 256, 256);
}

// examples/simple-cows.rs — cows made of seven spheres; the arc's posts are translated BEFORE they are scaled
PORTRAYER_EXAMPLE(simple_cows, "simple-cows") {
    auto stone = phong({0.8, 0.7, 0.7}, {0.0, 0.0, 0.0}, 0.0);
    auto grass = phong({0.1, 0.7, 0.1}, {0.0, 0.0, 0.0}, 0.0);
    auto cow_hide = phong({0.84, 0.6, 0.53}, {0.3, 0.3, 0.3}, 20.0);
    NodeRef arc = SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, stone)).translated({-1.9, 0.5, 0.1}).scaled({0.8, 4.0, 0.8}).into(),
        SceneNode::from(Geometry(Cube{}, stone)).translated({2.1, 0.5, 0.1}).scaled({0.8, 4.0, 0.8}).into(),
        SceneNode::from(Geometry(Sphere{}, stone)).scaled({4.0, 0.6, 0.6}).translated({0.0, 4.0, 0.0}).into(),
    }).translated({0.0, 0.0, -10.0}).into();
    auto ball = [&](double scale, Vec3 at) { return SceneNode::from(Geometry(Sphere{}, cow_hide)).scaled(scale).translated(at).into(); };
    NodeRef cow = SceneNode::from(std::vector<NodeRef>{
        ball(1.0, {0.0, 0.0, 0.0}), ball(0.6, {0.9, 0.3, 0.0}), ball(0.2, {-0.94, 0.34, 0.0}), ball(0.3, {0.7, -0.7, -0.7}),
        ball(0.3, {-0.7, -0.7, -0.7}), ball(0.3, {0.7, -0.7, 0.7}), ball(0.3, {-0.7, -0.7, 0.7}),
    }).into();
    return finish("simple-cows", cow_scene(stone, grass, arc, cow), camera({0.0, 2.0, 30.0}, {0.0, 2.0, 29.0}, 50.0), 256, 256);
}

// examples/single-triangle.rs — a Triangle used directly as a primitive
PORTRAYER_EXAMPLE(single_triangle, "single-triangle") {
    auto mat1 = phong({0.541, 0.169, 0.886}, {0.5, 0.7, 0.5}, 25.0);
    Triangle triangle = Triangle::flat({-1.0, 0.0, 0.0}, {1.0, 0.0, 0.0}, {0.0, 1.5, 0.0});
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{SceneNode::from(Geometry(triangle, mat1)).into()}).into(),
        .lights = {Light{.position = {1.0, 1.0, 10.0}, .color = {0.5, 0.5, 0.5}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("single-triangle", std::move(scene), camera({0.0, 0.5, 4.0}, {0.0, 0.5, 0.0}, 50.0), 640, 480);
}

// examples/smooth-shading.rs — the same meshes flat (left) and smooth (right)
PORTRAYER_EXAMPLE(smooth_shading, "smooth-shading") {
    auto mat_rock = phong({0.256361, 0.256361, 0.256361}, {0.6, 0.6, 0.6}, 50.0);
    auto mat_cow = phong({0.692066, 0.477245, 0.293336}, {0.3, 0.3, 0.3}, 25.0);
    auto mat_monkey = phong({0.261829, 0.8, 0.310477}, {0.3, 0.3, 0.3}, 25.0);
    auto monkey_mesh = MeshData::load_obj("assets/monkey.obj");
    auto cow_mesh = MeshData::load_obj("assets/cow.obj");
    auto flat_rock_mesh = MeshData::load_obj("assets/flat_rock.obj");
    auto smooth_rock_mesh = MeshData::load_obj("assets/smooth_rock.obj");
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Mesh(monkey_mesh, Shading::Flat), mat_monkey))
                .rotated_y(Radians::from_degrees(45.0)).translated({-1.904434, 1.4, 0.0}).into(),
            SceneNode::from(Geometry(Mesh(cow_mesh, Shading::Flat), mat_cow))
                .scaled(0.5).rotated_y(Radians::from_degrees(-15.0)).translated({-4.2, 1.8, 4.0}).into(),
            SceneNode::from(Geometry(Mesh(flat_rock_mesh, Shading::Flat), mat_rock)).translated({-3.396987, -1.4, 2.286671}).into(),
            SceneNode::from(Geometry(Mesh(monkey_mesh, Shading::Smooth), mat_monkey))
                .rotated_y(Radians::from_degrees(-45.0)).translated({1.242585, 1.4, 0.0}).into(),
            SceneNode::from(Geometry(Mesh(cow_mesh, Shading::Smooth), mat_cow))
                .scaled(0.5).rotated_y(Radians::from_degrees(205.0)).translated({3.8, 1.8, 4.0}).into(),
            SceneNode::from(Geometry(Mesh(smooth_rock_mesh, Shading::Smooth), mat_rock))
                .translated({3.271008, -1.406423, 2.372513}).into(),
        }).into(),
        .lights = {Light{.position = {0.0, 5.0, 10.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("smooth-shading", std::move(scene),
                  camera({1.062382, 0.54746, 22.827951}, {-0.813817, 0.424462, -8.112782}, 24.0), 910, 512);
}

// examples/primitives-simple.rs
PORTRAYER_EXAMPLE(primitives_simple, "primitives-simple") {
    auto mat_grass = Arc(Material{.diffuse = {0.173224, 0.8, 0.226505}});
    auto mat_cylinder = phong({0.139339, 0.435762, 0.8}, {0.3, 0.3, 0.3}, 25.0);
    auto mat_cone = phong({0.8, 0.047361, 0.04305}, {0.3, 0.3, 0.3}, 25.0);
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Cylinder{}, mat_cylinder)).scaled(2.0).translated({-2.0, 1.0, 0.0}).into(),
            SceneNode::from(Geometry(Cone{}, mat_cone)).scaled(2.0).translated({2.0, 1.0, 0.0}).into(),
            SceneNode::from(Geometry(Plane{}, mat_grass)).scaled(10.0).into(),
        }).into(),
        .lights = {Light{.position = {0.0, 10.0, 9.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("primitives-simple", std::move(scene),
                  camera({0.760838, 8.095396, 10.50759}, {-0.41716, -3.477774, -5.761218}, 25.0), 910, 512);
}

// examples/four-shapes.rs — white background
PORTRAYER_EXAMPLE(four_shapes, "four-shapes") {
    auto glass_like = [](Rgb diffuse) { return phong(diffuse, {0.3, 0.3, 0.3}, 100.0); };
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Sphere{}, glass_like({0.8, 0.0, 0.0}))).translated({-4.0, 0.0, 0.0}).into(),
            SceneNode::from(Geometry(Cube{}, glass_like({0.0, 0.158481, 0.8})))
                .scaled(1.6).rotated_y(Radians::from_degrees(-17.5411)).translated({-1.1, 0.0, 0.0}).into(),
            SceneNode::from(Geometry(Cone{}, glass_like({0.064785, 0.8, 0.174433}))).scaled(1.8).translated({1.5, 0.2, 0.0}).into(),
            SceneNode::from(Geometry(Cylinder{}, glass_like({0.127564, 0.016029, 0.8}))).scaled(1.6).translated({4.0, 0.0, 0.0}).into(),
        }).into(),
        .lights = {Light{.position = {0.0, 3.0, 11.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.1, 0.1, 0.1},
    };
    return finish("four-shapes", std::move(scene), camera({0.0, 6.473007, 15.607252}, {0.0, -2.181935, -5.702181}, 10.0),
                  1920, 512, white_background);
}

// examples/antialiasing.rs — rendered twice by the reference (SAMPLES = 1 and 32, set by the program itself)
PORTRAYER_EXAMPLE(antialiasing, "antialiasing") {
    auto mat_monkey = phong({0.961, 0.573, 0.259}, {0.3, 0.3, 0.3}, 25.0);
    auto monkey_mesh = MeshData::load_obj("assets/monkey.obj");
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{SceneNode::from(Geometry(Mesh(monkey_mesh, Shading::Flat), mat_monkey)).into()}).into(),
        .lights = {Light{.position = {0.0, 0.0, 10.0}, .color = {0.5, 0.5, 0.5}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("antialiasing", std::move(scene), camera({0.0, 0.0, 6.5}, {0.0, 0.0, 0.0}, 20.0), 300, 250);
}

// examples/fish.rs — a textured, smooth-shaded mesh with quads (fan-triangulated by the loader)
PORTRAYER_EXAMPLE(fish, "fish") {
    auto fish_skin = ImageTexture::open("assets/fish.png");
    auto mat_fish = Arc(Material{.diffuse = {0.8, 0.8, 0.8}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0, .texture = fish_skin});
    auto fish_model = MeshData::load_obj("assets/fish.obj");
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Mesh(fish_model, Shading::Smooth), mat_fish)).rotated_y(Radians::from_degrees(30.0)).into(),
            SceneNode::from(Geometry(Mesh(fish_model, Shading::Smooth), mat_fish)).rotated_y(Radians::from_degrees(210.0)).into(),
        }).into(),
        .lights = {Light{.position = {0.0, 0.0, 10.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("fish", std::move(scene), camera({0.0, 0.0, 11.0}, {0.0, 0.0, 0.0}, 25.0), 910, 512);
}

// examples/graphics-poster.rs — a dielectric, glossy dodecahedron around a smooth cow; white background
PORTRAYER_EXAMPLE(graphics_poster, "graphics-poster") {
    auto mat_glass = Arc(Material{.diffuse = {0.003638, 0.017153, 0.048247}, .specular = {0.5, 0.5, 0.5}, .shininess = 100.0,
                                  .reflectivity = 0.8, .glossy_side_length = 0.5,
                                  .refraction_index = OPTICAL_GLASS_REFRACTION_INDEX});
    auto mat_cow = phong({0.725682, 0.501253, 0.8}, {0.3, 0.3, 0.3}, 25.0);
    auto dodeca_model = MeshData::load_obj("assets/dodeca.obj");
    auto cow_model = MeshData::load_obj("assets/cow.obj");
    HierScene scene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Mesh(dodeca_model, Shading::Flat), mat_glass)).rotated_y(Radians::from_degrees(90.0)).into(),
            SceneNode::from(Geometry(Mesh(cow_model, Shading::Smooth), mat_cow))
                .scaled(0.24).rotated_y(Radians::from_degrees(-60.0)).into(),
        }).into(),
        .lights = {Light{.position = {1.33223, 4.297232, 3.473453}, .color = {0.9, 0.9, 0.9}},
                   Light{.position = {0.8, 0.806596, 0.9}, .color = {0.3, 0.3, 0.3}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    return finish("graphics-poster", std::move(scene),
                  camera({4.482203, 3.038775, 4.350142}, {-7.387217, -4.572944, -6.838186}, 35.0), 256, 256, white_background);
}
