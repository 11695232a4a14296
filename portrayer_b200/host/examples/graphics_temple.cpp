// examples/graphics-temple.rs — restated object for object (the example is unfinished upstream: its first floor's
// maze generator fills cells but emits NO nodes, `examples/graphics-temple.rs:127-183`, so that floor is an empty
// group here too and the StdRng draws it makes have no observable effect).  5 KDMeshes (grass, underwater land,
// one puppet instanced 3x through a shared node, teapot, cow), one linear Mesh (monkey), a glossy dielectric lake
// cube, 16 instanced columns of 5 primitives each, idols built from one shared rotated cube.  533x300.
#include <algorithm>

#include "examples.hpp"
using namespace portrayer;

namespace {
Radians deg(double d) { return Radians::from_degrees(d); }
MaterialRef placeholder_red() { return Arc(Material{.diffuse = {1.0, 0.0, 0.0}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0}); }

SceneNode hills() {
    auto mat_grass = Arc(Material{.diffuse = {0.376, 0.502, 0.22}});
    auto grass_model = MeshData::load_obj("assets/tog_grass.obj");
    return SceneNode::from(Geometry(KDMesh(*grass_model, Shading::Smooth), mat_grass)).translated({1.958125, 16.093138, -86.113747});
}

SceneNode lake() {
    auto mat_water = Arc(Material{.diffuse = {0.0, 0.0, 0.1}, .specular = {0.5, 0.5, 0.5}, .shininess = 100.0, .reflectivity = 0.9,
                                  .glossy_side_length = 1.0, .refraction_index = WATER_REFRACTION_INDEX});
    auto mat_dirt = Arc(Material{.diffuse = {0.592, 0.671, 0.055}});
    auto underwater_land_model = MeshData::load_obj("assets/tog_underwater_land.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_water)).scaled({600.0, 200.0, 600.0}).translated({0.0, -107.0, 300.0}).into(),
        SceneNode::from(Geometry(KDMesh(*underwater_land_model, Shading::Flat), mat_dirt)).translated({0.0, -107.0, 300.0}).into(),
    });
}

// examples/graphics-temple.rs:127-183: the loops over the maze have empty bodies
SceneNode temple_floor_1() { return SceneNode::from(std::vector<NodeRef>{}); }

// a cylindrical column with its centre at the bottom middle (examples/graphics-temple.rs:385-413)
SceneNode cylinder_column(const MaterialRef& mat_column) {
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_column)).scaled({3.2, 1.0, 3.2}).translated({0.0, 3.8, 0.0}).into(),
        SceneNode::from(Geometry(Cube{}, mat_column)).scaled({3.2, 1.0, 3.2}).translated({0.0, -3.8, 0.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_column)).scaled({1.5, 0.5, 1.5}).translated({0.0, 3.0, 0.0}).into(),
        SceneNode::from(Geometry(Sphere{}, mat_column)).scaled({1.5, 0.5, 1.5}).translated({0.0, -3.0, 0.0}).into(),
        SceneNode::from(Geometry(Cylinder{}, mat_column)).scaled({2.0, 6.0, 2.0}).into(),
    }).translated({0.0, 4.3, 0.0});
}

SceneNode temple_floor_2() {
    const double floor_width = 168.0, floor_height = 20.0, floor_length = 32.0, floor_y_offset = 20.0;
    const double floor_front_z = floor_length / 2.0;
    const size_t sections = 4;
    const double section_width = 30.0;
    const double column_scale = 2.0;
    const double column_diameter = 3.2 * column_scale;
    const double column_height = 8.6 * column_scale;
    const double section_spacing = (floor_width - column_diameter - (double)sections * section_width) / (double)(sections - 1);

    std::vector<NodeRef> nodes;
    auto mat_column = placeholder_red();
    NodeRef column = cylinder_column(mat_column).into();
    for (size_t i = 0; i < sections * 2; ++i) {
        const double x = section_width * (double)((i + 1) / 2) + section_spacing * (double)(i / 2) - floor_width / 2.0 + column_diameter / 2.0;
        nodes.push_back(SceneNode::from(column).scaled(column_scale)
                            .translated({x, floor_y_offset, floor_front_z - column_diameter / 2.0}).into());
        nodes.push_back(SceneNode::from(column).scaled(column_scale)
                            .translated({x, floor_y_offset, -(floor_front_z - column_diameter / 2.0)}).into());
    }
    const double ceiling_height = floor_height - column_height;
    nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_column)).scaled({floor_width, ceiling_height, floor_length})
                        .translated({0.0, floor_y_offset + column_height + ceiling_height / 2.0, 0.0}).into());

    // one "idol" per section: a cube and the three kinds of transformation applied to it
    auto mat_idol = placeholder_red();
    const double extent = std::min(section_width, column_height);
    NodeRef base_idol = SceneNode::from(Geometry(Cube{}, mat_idol)).scaled(extent * 0.5).rotated_y(deg(30.0)).into();
    std::vector<SceneNode> idols;
    idols.push_back(SceneNode::from(base_idol));
    idols.push_back(SceneNode::from(base_idol).scaled({1.0, 0.4, 1.0}));
    idols.push_back(SceneNode::from(base_idol).rotated_z(deg(80.0)));
    idols.push_back(SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(base_idol).scaled(0.5).translated({-extent / 4.0, extent / 8.0, -floor_length / 8.0}).into(),
        SceneNode::from(base_idol).scaled(0.5).translated({extent / 4.0, -extent / 8.0, floor_length / 8.0}).into(),
    }));
    for (size_t i = 0; i < idols.size(); ++i) {
        const double x = section_width * (double)(i + 1) + section_spacing * (double)i - floor_width / 2.0 - section_width / 2.0 +
                         column_diameter / 2.0;
        nodes.push_back(std::move(idols[i]).translated({x, floor_y_offset + column_height / 2.0, 0.0}).into());
    }
    return SceneNode::from(std::move(nodes));
}

SceneNode temple_floor_3() {
    const double floor_width = 117.6, floor_length = 25.6, floor_height = 20.0, floor_y_offset = 40.0;
    const double puppet_height = 17.2, puppet_y_offset = 44.083061;
    const double ceiling_height = floor_height - puppet_height;
    const double ceiling_y_offset = floor_y_offset + puppet_height + ceiling_height / 2.0;
    auto mat_puppet = placeholder_red();
    auto puppet_model = MeshData::load_obj("assets/tog_puppet.obj");
    NodeRef puppet = SceneNode::from(Geometry(KDMesh(*puppet_model, Shading::Smooth), mat_puppet)).translated({0.0, puppet_y_offset, 0.0}).into();
    auto mat_ceiling = placeholder_red();
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cube{}, mat_ceiling)).scaled({floor_width, ceiling_height, floor_length}).translated({0.0, ceiling_y_offset, 0.0}).into(),
        SceneNode::from(puppet).rotated_y(deg(90.0)).translated({-55.1, 0.0, 0.0}).into(),
        SceneNode::from(puppet).translated({0.0, 0.0, -5.0}).into(),
        SceneNode::from(puppet).rotated_y(deg(-90.0)).translated({55.1, 0.0, 0.0}).into(),
    });
}

SceneNode temple_floor_4() {
    auto mat_crystal = placeholder_red();
    auto monkey_model = MeshData::load_obj("assets/monkey.obj");
    auto teapot_model = MeshData::load_obj("assets/teapot.obj");
    auto cow_model = MeshData::load_obj("assets/cow.obj");
    return SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Mesh(monkey_model, Shading::Smooth), mat_crystal))
            .scaled(8.0).rotated_xzy(deg(-34.9072), deg(25.0), deg(0.0)).translated({-30.0, 64.214905, 1.0}).into(),
        SceneNode::from(Geometry(KDMesh(*teapot_model, Shading::Smooth), mat_crystal))
            .scaled(0.6).rotated_y(deg(-55.0)).translated({0.0, 59.857296, 0.0}).into(),
        SceneNode::from(Geometry(KDMesh(*cow_model, Shading::Smooth), mat_crystal))
            .scaled(1.5).rotated_y(deg(-125.0)).translated({30.0, 65.31517, 0.0}).into(),
    });
}
}  // namespace

PORTRAYER_EXAMPLE(graphics_temple, "graphics-temple") {
    ExampleScene ex;
    ex.name = "graphics-temple";
    auto mat_temple_block = Arc(Material{.diffuse = {0.913099, 0.913099, 0.715694}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            SceneNode::from(Geometry(Cube{}, mat_temple_block)).scaled({240.0, 20.0, 40.0}).translated({0.0, 10.0, 0.0}).into(),
            hills().into(), lake().into(), temple_floor_1().into(), temple_floor_2().into(), temple_floor_3().into(), temple_floor_4().into(),
        }).into(),
        .lights = {Light{.position = {190.0, 98.0, 151.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 61.971188, 546.971191}, .center = {0.0, -13.390381, -585.524353}, .up = Vec3::up(), .fovy = deg(25.0)};
    ex.width = 533;
    ex.height = 300;
    ex.background = [](Uv uv) { return Rgb{0.529, 0.808, 0.922} * (1.0 - uv.v) + Rgb{0.086, 0.38, 0.745} * uv.v; };
    return ex;
}
