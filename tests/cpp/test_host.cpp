// Host-side known-answer tests: the reference's own unit tests for the code
// that prepares the hot path's inputs, restated against the C++ host mirror.
//   src/bounding_box.rs:171-195   rotated_plane_bounds_90, rotated_cube_bounds_60
//   src/kdtree/leaf.rs:248-361    single_axis_center_partition, single_axis_uneven_partition
// plus self-checks of the vek-convention matrix algebra they rest on.
// Run by tests/test_host_kats.py; exits non-zero on the first failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "kdtree.hpp"
#include "pack.hpp"
#include "rand07.hpp"
#include "scene.hpp"

using namespace portrayer;

static int g_failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);      \
            ++g_failures;                                                    \
        }                                                                    \
    } while (0)

static double round_to(double x, double k) { return std::round(x * k) / k; }

// bounding_box.rs:171-182
static void rotated_plane_bounds_90() {
    BoundingBox bounds(Vec3{-0.5, 0.0, -0.5}, Vec3{0.5, 0.0, 0.5});
    Mat4 trans = Mat4::rotation_x(Radians::from_degrees(90.0).get());
    BoundingBox rb = trans * bounds;
    CHECK(round_to(rb.min().x, 10) == -0.5 && round_to(rb.min().y, 10) == -0.5 && round_to(rb.min().z, 10) == 0.0);
    CHECK(round_to(rb.max().x, 10) == 0.5 && round_to(rb.max().y, 10) == 0.5 && round_to(rb.max().z, 10) == 0.0);
}

// bounding_box.rs:184-195 — pins vek's composition order: scaling_3d(..).rotated_x(..) == Rx * S
static void rotated_cube_bounds_60() {
    BoundingBox bounds(Vec3{-0.5, -0.5, -0.5}, Vec3{0.5, 0.5, 0.5});
    Mat4 trans = Mat4::scaling_3d({8.0, 0.25, 5.0}).rotated_x(Radians::from_degrees(60.0).get());
    BoundingBox rb = trans * bounds;
    CHECK(round_to(rb.min().x, 1000) == -4.0 && round_to(rb.min().y, 1000) == -2.228 && round_to(rb.min().z, 1000) == -1.358);
    CHECK(round_to(rb.max().x, 1000) == 4.0 && round_to(rb.max().y, 1000) == 2.228 && round_to(rb.max().z, 1000) == 1.358);
}

// leaf.rs:263-267: a Plane rotated 90 degrees about z, translated to x
static FlatSceneNode plane_node_at(double x, const MaterialRef& mat) {
    return FlatSceneNode(Geometry(Plane{}, mat),
                         Mat4::rotation_z(Radians::from_degrees(90.0).get()).translated_3d({x, 0.0, 0.0}));
}

static std::vector<uint32_t> leaf_items(const KDIndexTree& n) {
    std::vector<uint32_t> out;
    for (const auto& nb : n.leaf.nodes) out.push_back(nb->node);
    return out;
}

static void partition_case(const std::vector<double>& xs, long merit, double expect_plane,
                           const std::vector<uint32_t>& expect_back, const std::vector<uint32_t>& expect_front) {
    auto mat = Arc(Material{});
    std::vector<FlatSceneNode> nodes;
    for (double x : xs) nodes.push_back(plane_node_at(x, mat));
    PartitionConfig conf{3, merit, 10};
    auto root = build_kdtree(nodes.size(), [&](size_t i) { return nodes[i].bounds(); }, 5, conf);
    CHECK(!root->is_leaf);
    if (root->is_leaf) return;
    CHECK(root->axis == 0);
    CHECK(root->plane == expect_plane);
    CHECK(root->front_nodes->is_leaf && root->back_nodes->is_leaf);
    CHECK(leaf_items(*root->front_nodes) == expect_front);
    CHECK(leaf_items(*root->back_nodes) == expect_back);
    // expected_root.bounds == nodes_bounds; leaf bounds == bounds of their members (leaf.rs:286-297)
    auto all = bounds_of(nodes.begin(), nodes.end(), [](const FlatSceneNode& n) { return n.bounds(); });
    CHECK(root->bounds == all);
    std::vector<BoundingBox> fb, bb;
    for (uint32_t i : expect_front) fb.push_back(nodes[i].bounds());
    for (uint32_t i : expect_back) bb.push_back(nodes[i].bounds());
    auto id = [](const BoundingBox& b) { return b; };
    CHECK(root->front_nodes->leaf.bounds == bounds_of(fb.begin(), fb.end(), id));
    CHECK(root->back_nodes->leaf.bounds == bounds_of(bb.begin(), bb.end(), id));
}

// leaf.rs:248-300: A B | C D E around S at x = 0
static void single_axis_center_partition() { partition_case({-8.0, -5.0, 3.0, 5.0, 8.0}, 3, 0.0, {0, 1}, {2, 3, 4}); }
// leaf.rs:302-360: plane found by bisection at x = 4.0
static void single_axis_uneven_partition() { partition_case({-8.0, 0.0, 3.0, 5.0, 8.0}, 2, 4.0, {0, 1, 2}, {3, 4}); }

static void matrix_algebra() {
    Mat4 m = Mat4::scaling_3d({2.0, 3.0, 0.5}).rotated_x(0.3).rotated_z(-1.1).rotated_y(2.0).translated_3d({5.0, -7.0, 11.0});
    Mat4 id = m * m.inverted();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) CHECK(std::fabs(id.m[i][j] - (i == j ? 1.0 : 0.0)) < 1e-12);
    // affine inverse keeps the last row exactly (0, 0, 0, ~1)
    Mat4 inv = m.inverted();
    CHECK(inv.m[3][0] == 0.0 && inv.m[3][1] == 0.0 && inv.m[3][2] == 0.0);
    // T(c) * Rx(b) * S(a): a point on the +y face of the unit cube ends where the composition says
    Mat4 t = Mat4::scaling_3d({1.0, 2.0, 1.0}).rotated_x(Radians::from_degrees(90.0).get()).translated_3d({0.0, 0.0, 10.0});
    Vec3 p = transformed_point(Vec3{0.0, 0.5, 0.0}, t);
    CHECK(std::fabs(p.x) < 1e-12 && std::fabs(p.y) < 1e-12 && std::fabs(p.z - 11.0) < 1e-12);
    // look_at_rh: the eye maps to the origin, the centre lies on -z
    Mat4 view = Mat4::look_at_rh({0.0, 0.0, 800.0}, {0.0, 0.0, 0.0}, Vec3::up());
    Vec3 e = transformed_point(Vec3{0.0, 0.0, 800.0}, view), c = transformed_point(Vec3{0.0, 0.0, 0.0}, view);
    CHECK(std::fabs(e.x) + std::fabs(e.y) + std::fabs(e.z) < 1e-9);
    CHECK(std::fabs(c.z + 800.0) < 1e-9);
}

static void flatten_order_and_instancing() {
    // BFS order with total_trans = parent * node (flat_scene.rs:18-46)
    auto mat = Arc(Material{});
    NodeRef shared = SceneNode::from(Geometry(Sphere{}, mat)).translated({1.0, 0.0, 0.0}).into();
    NodeRef left = SceneNode::from(shared).translated({0.0, 10.0, 0.0}).into();
    NodeRef right = SceneNode::from(std::vector<NodeRef>{shared, SceneNode::from(Geometry(Cube{}, mat)).into()})
                        .scaled(2.0).into();
    HierScene scene{SceneNode::from(std::vector<NodeRef>{left, right}).into(), {}, {}};
    FlatScene flat = FlatScene::from(scene);
    CHECK(flat.root.size() == 3);
    if (flat.root.size() != 3) return;
    CHECK(flat.root[0].geometry.primitive.kind == PrimKind::Sphere);  // left's child
    CHECK(flat.root[0].trans.m[0][3] == 1.0 && flat.root[0].trans.m[1][3] == 10.0);
    CHECK(flat.root[1].geometry.primitive.kind == PrimKind::Sphere);  // right's first child, scaled parent
    CHECK(flat.root[1].trans.m[0][3] == 2.0 && flat.root[1].trans.m[0][0] == 2.0);
    CHECK(flat.root[2].geometry.primitive.kind == PrimKind::Cube);
}

static void blob_roundtrip() {
    auto mat = Arc(Material{.diffuse = {0.1, 0.2, 0.3}, .reflectivity = 0.5});
    std::vector<NodeRef> kids;
    for (int i = 0; i < 20; ++i)
        kids.push_back(SceneNode::from(Geometry(i % 2 ? Primitive(Sphere{}) : Primitive(Cone{}), mat))
                           .translated({(double)i * 3.0, 0.0, 0.0}).into());
    HierScene scene{SceneNode::from(std::move(kids)).into(), {Light{.position = {0, 10, 0}, .color = {1, 1, 1}}}, {0.3, 0.3, 0.3}};
    KDTreeScene kd = KDTreeScene::from(FlatScene::from(scene), 10);
    std::vector<uint8_t> blob = pack_scene(kd);
    PtSceneDesc d;
    CHECK(pt_scene_unpack(blob.data(), blob.size(), &d) == PT_OK);
    CHECK(d.n_instances == 20 && d.n_materials == 1 && d.n_lights == 1);
    CHECK(d.tlas_extent == kd.root->extent());
    CHECK(d.n_tlas_nodes >= 3 && d.tlas_depth >= 1 && d.tlas_depth <= 10);
    // every instance appears in at least one leaf
    std::vector<int> seen(20, 0);
    for (uint32_t i = 0; i < d.n_tlas_items; ++i) seen[d.tlas_items[i]] = 1;
    for (int s : seen) CHECK(s == 1);
    // corrupting an index is rejected
    std::vector<uint8_t> bad = blob;
    PtBlobHeader h;
    std::memcpy(&h, bad.data(), sizeof h);
    uint32_t huge = 0x7fffffffu;
    std::memcpy(bad.data() + h.off_tlas_items, &huge, 4);
    CHECK(pt_scene_unpack(bad.data(), bad.size(), &d) == PT_ERR_INVALID);
    bad = blob;
    bad[0] ^= 0xFF;
    CHECK(pt_scene_unpack(bad.data(), bad.size(), &d) == PT_ERR_INVALID);
    CHECK(pt_scene_unpack(blob.data(), blob.size() / 2, &d) == PT_ERR_INVALID);
}

static void stdrng_stream() {
    // the generator must be a pure function of the seed, and u32/u64 draws must interleave without losing words
    StdRng a = StdRng::seed_from_u64(1234939301ull), b = StdRng::seed_from_u64(1234939301ull);
    for (int i = 0; i < 200; ++i) CHECK(a.next_u64() == b.next_u64());
    StdRng c = StdRng::seed_from_u64(7), d = StdRng::seed_from_u64(7);
    uint32_t lo = c.next_u32(), hi = c.next_u32();
    CHECK(d.next_u64() == ((uint64_t)hi << 32 | lo));
    for (int i = 0; i < 1000; ++i) {
        double f = a.gen_f64();
        CHECK(f >= 0.0 && f < 1.0);
        CHECK(a.gen_index(4) < 4);
    }
}

int main() {
    rotated_plane_bounds_90();
    rotated_cube_bounds_60();
    single_axis_center_partition();
    single_axis_uneven_partition();
    matrix_algebra();
    flatten_order_and_instancing();
    blob_roundtrip();
    stdrng_stream();
    if (g_failures) {
        std::printf("%d host check(s) FAILED\n", g_failures);
        return 1;
    }
    std::printf("all host checks passed\n");
    return 0;
}
