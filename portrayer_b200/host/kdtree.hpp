// k-d tree BUILD and scene flattening — the host-side preparation the
// reference runs before its pixel loop (src/render.rs:124-126).  In the target
// design this stays Rust; restated in C++ because no Rust toolchain exists in
// this image.  Must produce exactly the reference's trees: the traversal
// (device) reproduces the reference's quirks, so a different tree means
// different hit ids.
//
//   FlatScene::from / FlatSceneNode   src/flat_scene.rs:18-46,50-69,101-131
//   NodeBounds / PartitionConfig      src/kdtree/leaf.rs:12-67
//   KDLeaf::partitioned               src/kdtree/leaf.rs:89-231
//   KDTreeNode                        src/kdtree/node.rs:12-25
//   KDTreeScene::from                 src/kdtree/kdscene.rs:19-44
//   KDMesh::new                       src/kdtree/kdmesh.rs:37-58
#pragma once
#include <cstdlib>
#include <deque>
#include <memory>
#include <vector>

#include "scene.hpp"

namespace portrayer {

struct FlatSceneNode {
    Geometry geometry;
    Mat4 trans, invtrans, normal_trans;
    FlatSceneNode(Geometry g, const Mat4& t) : geometry(std::move(g)), trans(t) {
        invtrans = trans.inverted();
        normal_trans = invtrans.transposed();
    }
    BoundingBox bounds() const { return trans * geometry.primitive.bounds(); }  // flat_scene.rs:63-69
};

struct FlatScene {
    std::vector<FlatSceneNode> root;
    std::vector<Light> lights;
    Rgb ambient{};
    // Breadth-first flatten, total_trans = parent * node. flat_scene.rs:18-46
    static FlatScene from(const HierScene& hier) {
        FlatScene out;
        std::deque<std::pair<Mat4, NodeRef>> remaining;
        remaining.emplace_back(Mat4::identity(), hier.root);
        while (!remaining.empty()) {
            auto [parent_trans, node] = remaining.front();
            remaining.pop_front();
            Mat4 total_trans = parent_trans * node->trans();
            if (node->geometry()) out.root.emplace_back(*node->geometry(), total_trans);
            for (const auto& child : node->children()) remaining.emplace_back(total_trans, child);
        }
        out.lights = hier.lights;
        out.ambient = hier.ambient;
        return out;
    }
};

template <class T>
struct NodeBounds {
    BoundingBox bounds;
    T node;
};

struct PartitionConfig {
    size_t target_max_nodes = 3;
    long target_max_merit = 3;
    size_t max_tries = 10;
};

template <class T>
struct KDTreeNode;

template <class T>
struct KDLeaf {
    BoundingBox bounds;
    std::vector<std::shared_ptr<const NodeBounds<T>>> nodes;
    std::unique_ptr<KDTreeNode<T>> partitioned(int axis, size_t max_depth, const PartitionConfig& conf) &&;
};

template <class T>
struct KDTreeNode {
    bool is_leaf = true;
    // Split
    int axis = 0;        // sep_plane.normal = unit vector of this axis
    double plane = 0.0;  // sep_plane.point[axis] (other components are 0, leaf.rs:136-143)
    BoundingBox bounds;
    std::unique_ptr<KDTreeNode<T>> front_nodes, back_nodes;
    // Leaf
    KDLeaf<T> leaf;

    const BoundingBox& node_bounds() const { return is_leaf ? leaf.bounds : bounds; }
    double extent() const { return node_bounds().extent(); }  // node.rs:54-64
    size_t depth() const {
        if (is_leaf) return 0;
        size_t a = front_nodes->depth(), b = back_nodes->depth();
        return 1 + (a > b ? a : b);
    }
};

namespace detail {
enum class PlaneSide { Front, Back };
// InfinitePlane::which_side with an axis-unit normal: (p - point).dot(normal) >= 0
// reduces to the single component (the other two products are exact zeros).
// src/primitive/infinite_plane.rs:27-35
inline PlaneSide which_side(int axis, double plane, Vec3 p) {
    return (p[axis] - plane) >= 0.0 ? PlaneSide::Front : PlaneSide::Back;
}
enum class Partition { Front, Back, Shared };
template <class T>
Partition partition_node(const NodeBounds<T>& node, int axis, double plane) {  // leaf.rs:114-131
    PlaneSide a = which_side(axis, plane, node.bounds.min());
    PlaneSide b = which_side(axis, plane, node.bounds.max());
    if (a == PlaneSide::Front && b == PlaneSide::Front) return Partition::Front;
    if (a == PlaneSide::Back && b == PlaneSide::Back) return Partition::Back;
    return Partition::Shared;
}
}  // namespace detail

template <class T>
std::unique_ptr<KDTreeNode<T>> KDLeaf<T>::partitioned(int axis, size_t max_depth, const PartitionConfig& conf) && {
    auto out = std::make_unique<KDTreeNode<T>>();
    if (max_depth == 0 || nodes.size() <= conf.target_max_nodes) {  // leaf.rs:91-93
        out->is_leaf = true;
        out->leaf = std::move(*this);
        return out;
    }
    using detail::Partition;
    // centre of the bounding box along the axis. leaf.rs:135-148
    const double min_axis = bounds.min()[axis];
    const double max_axis = bounds.max()[axis];
    double plane = min_axis + (max_axis - min_axis) / 2.0;
    double plane_min = min_axis, plane_max = max_axis;

    for (size_t attempt = 0; attempt < conf.max_tries; ++attempt) {  // leaf.rs:156-201
        long front = 0, back = 0, shared = 0;
        for (const auto& n : nodes) {
            switch (detail::partition_node(*n, axis, plane)) {
                case Partition::Front: ++front; break;
                case Partition::Back: ++back; break;
                case Partition::Shared: ++shared; break;
            }
        }
        long merit = std::labs(front - back) + shared;
        if (merit <= conf.target_max_merit) break;
        if (front > back) {
            plane_min = plane;  // plane_range = (sep_plane.point, plane_max)
            plane = plane + (plane_max - plane) / 2.0;
        } else {
            plane_max = plane;  // plane_range = (plane_min, sep_plane.point)
            plane = plane_min + (plane - plane_min) / 2.0;
        }
    }

    KDLeaf<T> front_leaf, back_leaf;
    for (auto& n : nodes) {  // leaf.rs:204-215
        switch (detail::partition_node(*n, axis, plane)) {
            case Partition::Front: front_leaf.nodes.push_back(n); break;
            case Partition::Back: back_leaf.nodes.push_back(n); break;
            case Partition::Shared:
                front_leaf.nodes.push_back(n);
                back_leaf.nodes.push_back(n);
                break;
        }
    }
    auto get_bounds = [](const std::shared_ptr<const NodeBounds<T>>& n) { return n->bounds; };
    front_leaf.bounds = bounds_of(front_leaf.nodes.begin(), front_leaf.nodes.end(), get_bounds);
    back_leaf.bounds = bounds_of(back_leaf.nodes.begin(), back_leaf.nodes.end(), get_bounds);

    const int next = (axis + 1) % 3;  // (1,0,0) -> (0,1,0) -> (0,0,1). leaf.rs:97-103
    out->is_leaf = false;
    out->axis = axis;
    out->plane = plane;
    out->bounds = bounds;
    out->front_nodes = std::move(front_leaf).partitioned(next, max_depth - 1, conf);
    out->back_nodes = std::move(back_leaf).partitioned(next, max_depth - 1, conf);
    return out;
}

inline size_t env_depth(const char* name, size_t fallback) {  // kdscene.rs:36-38, kdmesh.rs:51-53
    const char* v = std::getenv(name);
    if (!v || !*v) return fallback;
    char* end = nullptr;
    unsigned long long d = std::strtoull(v, &end, 10);
    if (end == v || *end != '\0') return fallback;
    return static_cast<size_t>(d);
}

constexpr size_t MAX_TREE_DEPTH = 10;  // kdscene.rs:13, kdmesh.rs:15

// Build a tree over item INDICES 0..n-1 (bounds supplied per index).  The
// reference stores the items themselves (Arc<NodeBounds<T>>); indices keep the
// flat-instance / triangle ids that the device reports as hit ids.
template <class BoundsFn>
std::unique_ptr<KDTreeNode<uint32_t>> build_kdtree(size_t n, BoundsFn&& bounds_fn, size_t max_depth,
                                                   const PartitionConfig& conf = PartitionConfig{}) {
    KDLeaf<uint32_t> leaf;
    leaf.nodes.reserve(n);
    for (size_t i = 0; i < n; ++i)
        leaf.nodes.push_back(std::make_shared<const NodeBounds<uint32_t>>(
            NodeBounds<uint32_t>{bounds_fn(i), static_cast<uint32_t>(i)}));
    leaf.bounds = bounds_of(leaf.nodes.begin(), leaf.nodes.end(),
                            [](const std::shared_ptr<const NodeBounds<uint32_t>>& nb) { return nb->bounds; });
    return std::move(leaf).partitioned(0, max_depth, conf);
}

using KDIndexTree = KDTreeNode<uint32_t>;

// The tree behind a KDMesh: KDTreeNode<Triangle>, kdmesh.rs:19-24 (triangles
// kept in MeshData.triangles order, the tree refers to them by index).
struct KDMeshTree {
    std::vector<Triangle> tris;
    std::unique_ptr<KDIndexTree> root;
    bool has_normals = false, has_uvs = false;
};

struct KDTreeScene {  // Scene<KDTreeNode<FlatSceneNode>>, kdscene.rs:16
    std::vector<FlatSceneNode> nodes;  // flat instances, BFS order (flat_scene.rs:27-38)
    std::unique_ptr<KDIndexTree> root; // leaves hold indices into `nodes`
    std::vector<Light> lights;
    Rgb ambient{};
    static KDTreeScene from(FlatScene flat, size_t max_tree_depth) {
        KDTreeScene out;
        out.lights = std::move(flat.lights);
        out.ambient = flat.ambient;
        out.nodes = std::move(flat.root);
        const auto& nodes = out.nodes;
        out.root = build_kdtree(nodes.size(), [&nodes](size_t i) { return nodes[i].bounds(); }, max_tree_depth);
        return out;
    }
    static KDTreeScene from(FlatScene flat) { return from(std::move(flat), env_depth("KD_DEPTH", MAX_TREE_DEPTH)); }
};

}  // namespace portrayer
