// Host-callable launch wrappers of kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include "device_scene.cuh"
#include "kd_build.h"

namespace ptd {
void kernels_init();
// copy a frame's constants into __constant__ slot `slot` (stream-ordered)
cudaError_t upload_state(int slot, const FrameState& state, cudaStream_t st);

double launch_fp64_rate(double* sink, int iters, cudaStream_t st);
void launch_gamma_lut(double* lut256, cudaStream_t st);
// upload time: padded FP32 world boxes of all instances (mesh_bounds_scratch: n_meshes * 6 doubles)
void launch_instance_bounds(const DScene& sc, uint32_t n_meshes, double* mesh_bounds_scratch, float4* out, cudaStream_t st);

// upload time: the leaf cull structure of a forest of k-d trees (leaf_cull.cu).  One LcTree per tree: where its nodes
// and items lie in the forest's arrays (child / item indices inside the records are relative to these) and what to add
// to an item value to index `item_boxes` (0 for the scene tree, tri_first for a KDMesh tree).
struct LcTree {
    uint32_t node_first, node_count, item_first, box_first;
};
struct LeafCullSizes {
    uint32_t n_leaf_cap, n_grp_cap;
    size_t set_float4;     // float4 per box set; the storage holds two sets
    size_t scratch_bytes;
};
LeafCullSizes leaf_cull_sizes(uint32_t n_nodes, uint32_t n_items);
// nodes: the forest's node array in device memory — the leaves' `split` fields are overwritten with rank | gbase << 32
// max_depth: the deepest tree of the forest (sweeps of the bottom-up node-box union)
cudaError_t launch_leaf_cull(PtKdNode* nodes, uint32_t n_nodes, const uint32_t* items, uint32_t n_items, const LcTree* d_trees, uint32_t n_trees,
                             uint32_t max_depth, const float4* item_boxes, float4* storage, void* scratch, const KdAllocator& al, LeafCull* out,
                             cudaStream_t st);
void launch_root_box(const float4* boxes, uint32_t n, float4* out2, cudaStream_t st);
// upload time: padded FP32 object-space boxes of all triangles
void launch_triangle_bounds(const PtTriPos* tri_pos, uint32_t n, float4* tri_aabb, cudaStream_t st);
// fold structure of linear meshes (fold_order.cu, traverse.cuh mesh_fold)
size_t fold_sort_temp_bytes(uint32_t n);
size_t fold_scratch_bytes(uint32_t n, size_t temp_bytes);
cudaError_t launch_fold_order(const PtMesh* meshes, uint32_t n_meshes, const PtTriPos* tri_pos, const double* mesh_bounds, uint32_t n, bool sort,
                              uint32_t* order, void* scratch, size_t temp_bytes, int end_bit, cudaStream_t st);
void launch_gather_fold_boxes(const float4* tri_aabb, const uint32_t* order, uint32_t n, float4* out, cudaStream_t st);
void launch_group_bounds(const float4* in, uint32_t n, uint32_t run, float4* out, cudaStream_t st);

// stream path: one launch per call; batch / level bookkeeping lives in the device control block
void launch_camera(int slot, uint32_t first_slot, uint32_t n_slots, uint32_t samples, cudaStream_t st);
void launch_load_rays(int slot, uint32_t first, uint32_t n_paths, cudaStream_t st);
// which variant of the traversal kernels walks the trees: the reference's walk as it is (PT_RENDER_EXACT_WALK), the same
// with the work counters (PT_RENDER_COUNTERS), or the walk that skips subtrees the ray cannot hit anything in (default)
enum { kWalkExact = 0, kWalkCount = 1, kWalkPrune = 2 };
// linear: the PT_RENDER_LINEAR_TLAS cross-check kernels (always counting)
void launch_extend(int slot, uint64_t max_items, int mode, bool linear, cudaStream_t st);
void launch_shadow(int slot, uint64_t max_items, uint32_t n_lights, int mode, bool linear, cudaStream_t st);
// big: allocate the children of 512 parents together instead of 128 (frames whose batches hold >= 2 Mi paths: shade_big_group)
bool shade_big_group(uint64_t batch_paths);
void launch_shade(int slot, uint64_t max_items, bool big, cudaGraphConditionalHandle loop, cudaStream_t st);
void launch_tree_eval(int slot, uint32_t n_paths, cudaStream_t st);
void launch_resolve(int slot, uint32_t n_slots, cudaStream_t st);
void launch_export_rays(int slot, uint32_t n_paths, cudaStream_t st);

// image_io.cu: the byte-moving steps either side of the render loop
cudaError_t launch_to_rgb(const uint8_t* src, uint64_t n_pixels, uint32_t channels, uint32_t r_at, uint32_t g_at, uint32_t b_at, uint8_t* dst,
                          cudaStream_t st);
uint64_t png_file_bytes(uint32_t width, uint32_t height);
uint64_t png_scratch_bytes(uint32_t width, uint32_t height);
cudaError_t launch_png_encode(const uint8_t* d_rgb, uint32_t width, uint32_t height, uint8_t* d_png, void* d_scratch, cudaStream_t st);

// graph path: camera -> WHILE(level has rays){extend, shadow, shade} -> tree_eval -> resolve
cudaError_t build_frame_graph(int slot, uint32_t n_slots, uint32_t samples, uint32_t n_lights_max, uint64_t capacity,
                              int mode, cudaGraph_t* graph_out, cudaGraphExec_t* exec_out, cudaGraphNode_t* camera_node);
cudaError_t set_graph_batch(cudaGraphExec_t exec, cudaGraphNode_t camera_node, int slot, uint32_t first_slot, uint32_t n_slots,
                            uint32_t grid_slots, uint32_t samples);
}  // namespace ptd
