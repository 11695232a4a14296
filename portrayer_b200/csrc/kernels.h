// Host-callable launch wrappers of kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include "device_scene.cuh"

namespace ptd {
void kernels_init();
// copy a frame's constants into __constant__ slot `slot` (stream-ordered)
cudaError_t upload_state(int slot, const FrameState& state, cudaStream_t st);

double launch_fp64_rate(double* sink, int iters, cudaStream_t st);
void launch_gamma_lut(double* lut256, cudaStream_t st);
// upload time: padded FP32 world boxes of all instances (mesh_bounds_scratch: n_meshes * 6 doubles)
void launch_instance_bounds(const DScene& sc, uint32_t n_meshes, double* mesh_bounds_scratch, float4* out, cudaStream_t st);

// upload time: the instance boxes gathered into scene-tree leaf order + the union box of every aligned run of 8 of
// them (out: 2 * (n_items + ceil(n_items / 8)) float4)
void launch_gather_leaf_boxes(const float4* inst_aabb, const uint32_t* items, uint32_t n_items, float4* out, cudaStream_t st);
// upload time: padded FP32 object-space boxes of all triangles + of every aligned run of 32 / 1024 of them
void launch_gather_blas_leaf_boxes(const PtMesh* meshes, uint32_t n_meshes, const uint32_t* blas_items, const float4* tri_aabb, float4* out,
                                   uint32_t n_blas_items, cudaStream_t st);
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
void launch_extend(int slot, uint64_t max_items, bool count, cudaStream_t st);
void launch_shadow(int slot, uint64_t max_items, uint32_t n_lights, bool count, cudaStream_t st);
void launch_shade(int slot, uint64_t max_items, cudaGraphConditionalHandle loop, cudaStream_t st);
void launch_tree_eval(int slot, uint32_t n_paths, cudaStream_t st);
void launch_resolve(int slot, uint32_t n_slots, cudaStream_t st);
void launch_export_rays(int slot, uint32_t n_paths, cudaStream_t st);

// graph path: camera -> WHILE(level has rays){extend, shadow, shade} -> tree_eval -> resolve
cudaError_t build_frame_graph(int slot, uint32_t n_slots, uint32_t samples, uint32_t n_lights_max, uint64_t capacity,
                              bool count, cudaGraph_t* graph_out, cudaGraphExec_t* exec_out, cudaGraphNode_t* camera_node);
cudaError_t set_graph_batch(cudaGraphExec_t exec, cudaGraphNode_t camera_node, int slot, uint32_t first_slot, uint32_t n_slots,
                            uint32_t grid_slots, uint32_t samples);
}  // namespace ptd
