// The "new Rust glue" of SURVEY §8b restated in C++: walk the prepared
// KDTreeScene (KDTreeNode src/kdtree/node.rs:13-25, KDLeaf leaf.rs:70-78,
// FlatSceneNode flat_scene.rs:50-61, KDMesh kdmesh.rs:19-24, MeshData
// mesh.rs:22-34) and emit the SoA buffers of include/portrayer_gpu.h, packed
// into one pointer-free blob.
#pragma once
#include <cstdint>
#include <vector>

#include "kdtree.hpp"
#include "portrayer_gpu.h"

namespace portrayer {

// Camera::new, src/camera.rs:34-45
PtCamera make_camera(const CameraSettings& cam, double width, double height);

// Serialise a prepared scene. `linear_tlas` puts every instance into one root
// leaf in flat order (the `flat_scene` cargo feature: FlatScene ray_cast,
// flat_scene.rs:71-99 + ray.rs:87-99).
std::vector<uint8_t> pack_scene(const KDTreeScene& scene);
std::vector<uint8_t> pack_scene(const std::vector<FlatSceneNode>& nodes, const KDIndexTree& root,
                                const std::vector<Light>& lights, Rgb ambient);

// The scene GRAPH as the boundary's pt_flatten takes it (include/portrayer_gpu.h): one PtHierNode per distinct
// SceneNode (shared nodes once: instancing), concatenated child lists, one PtGeometryRec per node that carries
// geometry.  mesh / material ids are the ones pack_scene assigns (first appearance in flat-instance order), so the
// device-flattened records can be compared with, and used next to, a packed blob of the same scene.
struct HierarchyExport {
    std::vector<PtHierNode> nodes;
    std::vector<uint32_t> children;
    std::vector<PtGeometryRec> geometries;
    uint32_t root = 0;
};
HierarchyExport export_hierarchy(const HierScene& scene);

// Breadth-first serialisation of a tree into the PtKdNode / leaf-item records of the boundary (front child, then back
// child; leaf lists in visiting order).  Returns the depth of the deepest node (root = 0).
uint32_t serialise_kd_tree(const KDIndexTree& root, std::vector<PtKdNode>& nodes_out, std::vector<uint32_t>& items_out);

}  // namespace portrayer
