// examples/primitives.rs — BASELINE.json configs[1]
#include "examples.hpp"
using namespace portrayer;

static SceneNode make_castle() {
    auto mat_dome = Arc(Material{.diffuse = {0.609065, 0.731162, 0.8}, .specular = {0.5, 0.5, 0.5},
                                 .shininess = 1000.0, .reflectivity = 0.3});
    auto mat_castle = Arc(Material{.diffuse = {0.769051, 0.304112, 0.8}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto mat_castle_tower_top =
        Arc(Material{.diffuse = {0.352613, 0.42773, 0.8}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto mat_castle_door = Arc(Material{.diffuse = {0.176099, 0.115632, 0.054921}});
    auto mat_road = Arc(Material{.diffuse = {0.121484, 0.024035, 0.0}});

    std::vector<NodeRef> nodes;

    const double castle_width = 4.0;
    const double castle_length = castle_width;
    const double castle_height = 2.0;
    const double dome_radius = castle_width / 2.0;
    const double tower_height = castle_height * 1.5;
    const double tower_width = 1.5;
    const double tower_roof_height = 2.0;
    const double tower_roof_width = tower_width + 0.1;

    // Main castle body
    nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_castle))
                        .scaled({castle_width, castle_height, castle_length})
                        .translated({0.0, castle_height / 2.0, 0.0}).into());
    // Castle dome
    nodes.push_back(SceneNode::from(Geometry(Sphere{}, mat_dome))
                        .scaled({dome_radius, castle_height, dome_radius})
                        .translated({0.0, castle_height, 0.0}).into());
    // Castle door
    auto castle_door_model = MeshData::load_obj("assets/prim_castle_door.obj");
    nodes.push_back(SceneNode::from(Geometry(Mesh(castle_door_model, Shading::Smooth), mat_castle_door))
                        .translated({0.0, 1.1, castle_length / 2.0 + 0.1}).into());
    // Road
    nodes.push_back(SceneNode::from(Geometry(Cube{}, mat_road))
                        .scaled({2.0, 0.01, 4.0})
                        .translated({0.0, 0.0, castle_length / 2.0 + 2.0 - 0.3}).into());

    // All 4 towers
    NodeRef tower = SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cylinder{}, mat_castle))
            .scaled({tower_width, tower_height, tower_width})
            .translated({0.0, tower_height / 2.0, 0.0}).into(),
        SceneNode::from(Geometry(Cone{}, mat_castle_tower_top))
            .scaled({tower_roof_width, tower_roof_height, tower_roof_width})
            .translated({0.0, tower_height + tower_roof_height / 2.0, 0.0}).into(),
    }).into();

    for (double x : {-1.0, 1.0})
        for (double z : {-1.0, 1.0}) {
            Vec3 tower_pos{castle_width / 2.0 * x, 0.0, castle_length / 2.0 * z};
            nodes.push_back(SceneNode::from(tower).translated(tower_pos).into());
        }
    return SceneNode::from(std::move(nodes));
}

static SceneNode make_trees() {
    auto mat_tree_leaves = Arc(Material{.diffuse = {0.289596, 0.8, 0.308959}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto mat_tree_trunk = Arc(Material{.diffuse = {0.8, 0.441708, 0.115746}});

    NodeRef tree = SceneNode::from(std::vector<NodeRef>{
        SceneNode::from(Geometry(Cylinder{}, mat_tree_trunk)).scaled({0.3, 2.0, 0.3}).translated({0.0, 1.0, 0.0}).into(),
        SceneNode::from(Geometry(Cone{}, mat_tree_leaves)).scaled({1.0, 2.0, 1.0}).translated({0.0, 2.9, 0.0}).into(),
    }).into();

    const Vec3 tree_positions[] = {
        // Trees to the right of the camera
        {4.225878, 0.0, 3.695781}, {5.225877, 0.0, 2.895781}, {4.125877, 0.0, 2.395781}, {5.125877, 0.0, 1.595781},
        {6.525877, 0.0, 0.795781}, {5.125877, 0.0, 0.395781}, {5.925876, 0.0, -0.704219}, {4.725877, 0.0, -1.30422},
        {3.425877, 0.0, -0.804219}, {3.025877, 0.0, -2.204219}, {4.225877, 0.0, -2.30422}, {5.425877, 0.0, -2.50422},
        {6.525876, 0.0, -2.00422}, {6.925876, 0.0, -3.50422}, {5.825876, 0.0, -3.90422}, {4.625876, 0.0, -3.70422},
        {3.425876, 0.0, -3.40422}, {3.625876, 0.0, -4.80422}, {5.025876, 0.0, -5.10422}, {6.825876, 0.0, -5.00422},
        // Trees to the left of the camera
        {-3.374122, 0.0, 3.79578}, {-4.874123, 0.0, 3.29578}, {-2.874123, 0.0, 2.39578}, {-4.374123, 0.0, 2.19578},
        {-5.674122, 0.0, 1.79578}, {-5.974123, 0.0, 0.195781}, {-4.674122, 0.0, 0.395781}, {-3.574123, 0.0, 1.09578},
        {-3.274122, 0.0, -0.204219}, {-4.674122, 0.0, -1.00422}, {-5.874123, 0.0, -1.20422}, {-5.874123, 0.0, -2.40422},
        {-4.574122, 0.0, -2.40422}, {-3.474122, 0.0, -1.70422}, {-3.574123, 0.0, -3.30422}, {-5.374123, 0.0, -3.60422},
    };

    NodeRef fallen_tree = SceneNode::from(tree)
        .rotated_xzy(Radians::from_degrees(0.0), Radians::from_degrees(50.0), Radians::from_degrees(-80.0))
        .translated({2.285154, 0.13965, 2.474418}).into();

    std::vector<NodeRef> nodes;
    for (const Vec3& tree_pos : tree_positions) nodes.push_back(SceneNode::from(tree).translated(tree_pos).into());
    nodes.push_back(fallen_tree);
    return SceneNode::from(std::move(nodes));
}

PORTRAYER_EXAMPLE(primitives, "primitives") {
    auto mat_grass = Arc(Material{.diffuse = {0.177353, 0.334328, 0.169638}});

    NodeRef castle = make_castle().translated({0.0, 0.0, -1.6}).into();
    NodeRef trees = make_trees().into();

    ExampleScene ex;
    ex.name = "primitives";
    ex.scene = HierScene{
        .root = SceneNode::from(std::vector<NodeRef>{
            castle,
            trees,
            // Floor
            SceneNode::from(Geometry(Plane{}, mat_grass)).scaled(30.0).into(),
        }).into(),
        .lights = {Light{.position = {0.0, 10.0, 9.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 4.311144, 17.370693}, .center = {0.0, 2.133119, -7.534255},
                            .up = Vec3::up(), .fovy = Radians::from_degrees(25.0)};
    ex.width = 910;
    ex.height = 512;
    ex.background = sky_gradient;
    return ex;
}
