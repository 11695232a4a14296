// Known-answer scenes restating the reference's own unit tests for the hot
// path, so oracle and device can be run on exactly what those tests build:
//   kat-edge-case / kat-edge-case-flipped   src/kdtree/node.rs:219-352
//   kat-mesh-equivalence-{mesh,kdmesh}      src/kdtree/kdmesh.rs:99-166
#include "examples.hpp"
using namespace portrayer;

namespace {

std::unique_ptr<KDIndexTree> leaf_of(std::vector<uint32_t> items, const std::vector<FlatSceneNode>& nodes) {
    auto n = std::make_unique<KDIndexTree>();
    n->is_leaf = true;
    // leaf bounds do not matter currently (node.rs:276-277)
    n->leaf.bounds = BoundingBox(Vec3::zero(), Vec3::zero());
    for (uint32_t i : items)
        n->leaf.nodes.push_back(std::make_shared<const NodeBounds<uint32_t>>(NodeBounds<uint32_t>{nodes[i].bounds(), i}));
    return n;
}

// node.rs:219-293 (flipped = false) and node.rs:295-351 (flipped = true).
// Instance 0 = B (red), instance 1 = C (blue); the expected hit is B.
ExampleScene edge_case(bool flipped) {
    const double sgn = flipped ? -1.0 : 1.0;
    auto mat_b = Arc(Material{.diffuse = Rgb::red()});
    auto mat_c = Arc(Material{.diffuse = Rgb::blue()});

    Mat4 trans_b = Mat4::scaling_3d(2.0).rotated_x(Radians::from_degrees(sgn * 90.0).get()).translated_3d({0.0, 1.2, sgn * -0.4});
    Mat4 trans_c = Mat4::scaling_3d(2.0).rotated_x(Radians::from_degrees(sgn * 50.0).get()).translated_3d({0.0, 0.0, sgn * -0.3});

    auto kd = std::make_shared<KDTreeScene>();
    kd->nodes.emplace_back(Geometry(Plane{}, mat_b), trans_b);
    kd->nodes.emplace_back(Geometry(Plane{}, mat_c), trans_c);
    const uint32_t B = 0, C = 1;

    auto root = std::make_unique<KDIndexTree>();
    root->is_leaf = false;
    root->axis = 2;  // sep_plane normal = unit_z, point = zero
    root->plane = 0.0;
    std::vector<BoundingBox> both = {kd->nodes[B].bounds(), kd->nodes[C].bounds()};
    root->bounds = bounds_of(both.begin(), both.end(), [](const BoundingBox& b) { return b; });
    if (!flipped) {
        root->front_nodes = leaf_of({C}, kd->nodes);
        root->back_nodes = leaf_of({C, B}, kd->nodes);  // Force tree to check C again by putting it first
    } else {
        root->front_nodes = leaf_of({C, B}, kd->nodes);
        root->back_nodes = leaf_of({C}, kd->nodes);
    }
    kd->root = std::move(root);
    kd->ambient = {1.0, 1.0, 1.0};  // colour of a hit == the material's diffuse

    ExampleScene ex;
    ex.name = flipped ? "kat-edge-case-flipped" : "kat-edge-case";
    ex.prebuilt = kd;
    ex.cam = CameraSettings{.eye = {0.0, 0.5, sgn * 0.9}, .center = {0.0, 0.5, 0.0}, .up = Vec3::up(),
                            .fovy = Radians::from_degrees(1.0)};
    ex.width = 1;
    ex.height = 1;
    ex.background = [](Uv) { return Rgb::black(); };
    return ex;
}

// kdmesh.rs:99-166
ExampleScene mesh_equivalence(bool kd) {
    auto mat_castle_walls = Arc(Material{.diffuse = {1.0, 0.0, 0.0}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    auto model = MeshData::load_obj("assets/castle.obj");
    Primitive prim = kd ? Primitive(KDMesh(*model, Shading::Flat, MAX_TREE_DEPTH)) : Primitive(Mesh(model, Shading::Flat));

    ExampleScene ex;
    ex.name = kd ? "kat-mesh-equivalence-kdmesh" : "kat-mesh-equivalence-mesh";
    ex.scene = HierScene{
        .root = SceneNode::from(Geometry(prim, mat_castle_walls)).scaled(1.4).translated({0.0, 0.0, -229.0}).into(),
        .lights = {Light{.position = {50.0, 110.0, -120.0}, .color = {0.9, 0.9, 0.9}}},
        .ambient = {0.3, 0.3, 0.3},
    };
    ex.cam = CameraSettings{.eye = {0.0, 120.0, 240.0}, .center = {0.0, 100.0, -24.0}, .up = Vec3::up(),
                            .fovy = Radians::from_degrees(25.0)};
    ex.width = 533;
    ex.height = 300;
    ex.background = [](Uv) { return Rgb::black(); };
    return ex;
}

// Degenerate inputs the render loop must survive exactly like the reference:
//   edge-empty        no geometry at all: every ray returns the background (ray.rs:146); one light that is never used
//   edge-no-lights    geometry but no lights: ambient only, no shadow rays (material.rs:149 iterates an empty Vec)
//   edge-degenerate   a zero-area triangle, a plane scaled to zero width and a sphere scaled flat: the reference's
//                     arithmetic produces NaN / inf there (0 / 0 in Cramer's rule, singular inverses) and every
//                     comparison with them is false, i.e. "no hit" — the device must agree bit for bit
ExampleScene edge_scene(int which) {
    auto mat = Arc(Material{.diffuse = {0.8, 0.3, 0.2}, .specular = {0.3, 0.3, 0.3}, .shininess = 25.0});
    std::vector<NodeRef> nodes;
    std::vector<Light> lights = {Light{.position = {3.0, 5.0, 8.0}, .color = {0.9, 0.9, 0.9}}};
    if (which == 1) {
        nodes.push_back(SceneNode::from(Geometry(Sphere{}, mat)).scaled(1.5).into());
        nodes.push_back(SceneNode::from(Geometry(Cube{}, mat)).scaled({6.0, 0.2, 6.0}).translated({0.0, -1.6, 0.0}).into());
        lights.clear();
    } else if (which == 2) {
        nodes.push_back(SceneNode::from(Geometry(Triangle::flat({-1.0, 0.0, 0.0}, {1.0, 0.0, 0.0}, {3.0, 0.0, 0.0}), mat)).into());
        nodes.push_back(SceneNode::from(Geometry(Plane{}, mat)).scaled({0.0, 1.0, 4.0}).rotated_x(Radians::from_degrees(90.0)).into());
        nodes.push_back(SceneNode::from(Geometry(Sphere{}, mat)).scaled({1.0, 0.0, 1.0}).translated({0.0, 1.0, 0.0}).into());
        nodes.push_back(SceneNode::from(Geometry(Cylinder{}, mat)).scaled(0.8).translated({-2.0, 0.0, 0.0}).into());
        nodes.push_back(SceneNode::from(Geometry(Cone{}, mat)).scaled(0.8).translated({2.0, 0.0, 0.0}).into());
    }
    ExampleScene ex;
    ex.name = which == 0 ? "edge-empty" : which == 1 ? "edge-no-lights" : "edge-degenerate";
    ex.scene = HierScene{.root = SceneNode::from(std::move(nodes)).into(), .lights = lights, .ambient = {0.3, 0.3, 0.3}};
    ex.cam = CameraSettings{.eye = {0.0, 1.0, 9.0}, .center = {0.0, 0.0, 0.0}, .up = Vec3::up(), .fovy = Radians::from_degrees(40.0)};
    ex.width = 203;  // neither a multiple of the 32x32 tile nor of the 8x4 warp block
    ex.height = 117;
    ex.background = sky_gradient;
    return ex;
}

// Exact ties inside one linear Mesh: coplanar, overlapping triangles with small-integer coordinates in the plane
// z = -4, listed so that INDEX order and spatial (Morton) order disagree, plus 400 filler triangles behind them.  For an
// axis-aligned ray (0, 0, -1) from z = 0 every product in Triangle::ray_hit is exact, so all triangles covering a point
// return t == 4.0 bit for bit and the fold's rule decides: the FIRST listed wins (ray.rs:50-63: later candidates need
// t < e_now).  tests/: the oracle against the brute-force answer, the device fold (which visits triangles in Morton
// order) against the oracle.
ExampleScene mesh_ties() {
    std::vector<Vec3> pos;
    std::vector<std::array<size_t, 3>> tris;
    auto tri = [&](Vec3 a, Vec3 b, Vec3 c) {
        const size_t base = pos.size();
        pos.push_back(a); pos.push_back(b); pos.push_back(c);
        tris.push_back({base, base + 1, base + 2});
    };
    const double z = -4.0;
    tri({8.0, 8.0, z}, {12.0, 8.0, z}, {8.0, 12.0, z});      // 0: small, far corner (high Morton)
    tri({0.0, 0.0, z}, {16.0, 0.0, z}, {0.0, 16.0, z});      // 1: big, covers 0, 2, 3, 5
    tri({0.0, 0.0, z}, {4.0, 0.0, z}, {0.0, 4.0, z});        // 2: small, near corner (low Morton), listed AFTER the big one
    tri({2.0, 2.0, z}, {6.0, 2.0, z}, {2.0, 6.0, z});        // 3: small, overlaps 2 and 1
    tri({0.0, 0.0, z}, {16.0, 0.0, z}, {0.0, 16.0, z});      // 4: the big one again
    tri({1.0, 1.0, z}, {3.0, 1.0, z}, {1.0, 3.0, z});        // 5: inside 2
    tri({20.0, 0.0, z}, {24.0, 0.0, z}, {20.0, 4.0, z});     // 6: alone
    tri({20.0, 0.0, z}, {24.0, 0.0, z}, {20.0, 4.0, z});     // 7: its twin: loses to 6 everywhere
    for (int j = 0; j < 20; ++j)                              // fillers behind (z = -8), some under the tie region
        for (int i = 0; i < 20; ++i)
            tri({(double)i, (double)j, -8.0}, {(double)i + 1.0, (double)j, -8.0}, {(double)i, (double)j + 1.0, -8.0});
    auto data = std::make_shared<const MeshData>(std::move(pos), std::move(tris), std::vector<Vec3>{}, std::vector<Uv>{});
    auto mat = Arc(Material{.diffuse = {0.2, 0.6, 0.9}});
    ExampleScene ex;
    ex.name = "edge-mesh-ties";
    ex.scene = HierScene{.root = SceneNode::from(Geometry(Mesh(data, Shading::Flat), mat)).into(),
                         .lights = {Light{.position = {8.0, 8.0, 20.0}, .color = {0.9, 0.9, 0.9}}},
                         .ambient = {0.3, 0.3, 0.3}};
    ex.cam = CameraSettings{.eye = {10.0, 8.0, 30.0}, .center = {10.0, 8.0, -4.0}, .up = Vec3::up(), .fovy = Radians::from_degrees(45.0)};
    ex.width = 160;
    ex.height = 120;
    ex.background = sky_gradient;
    return ex;
}

}  // namespace

PORTRAYER_EXAMPLE(edge_mesh_ties, "edge-mesh-ties") { return mesh_ties(); }
PORTRAYER_EXAMPLE(edge_empty, "edge-empty") { return edge_scene(0); }
PORTRAYER_EXAMPLE(edge_no_lights, "edge-no-lights") { return edge_scene(1); }
PORTRAYER_EXAMPLE(edge_degenerate, "edge-degenerate") { return edge_scene(2); }
PORTRAYER_EXAMPLE(kat_edge_case, "kat-edge-case") { return edge_case(false); }
PORTRAYER_EXAMPLE(kat_edge_case_flipped, "kat-edge-case-flipped") { return edge_case(true); }
PORTRAYER_EXAMPLE(kat_mesh_eq_mesh, "kat-mesh-equivalence-mesh") { return mesh_equivalence(false); }
PORTRAYER_EXAMPLE(kat_mesh_eq_kdmesh, "kat-mesh-equivalence-kdmesh") { return mesh_equivalence(true); }
