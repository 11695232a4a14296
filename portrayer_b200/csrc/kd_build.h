// Device k-d tree build (kd_build.cu); C ABI wrappers live in api.cu (pt_kd_build*).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "portrayer_gpu.h"

namespace ptd {

struct KdTreeDev;

// the caller's device allocator (stream-ordered reuse is fine: everything runs on the one stream given to the build)
struct KdAllocator {
    void* (*alloc)(size_t bytes, cudaError_t* err) = nullptr;
    void (*release)(void* p) = nullptr;
};

// d_bounds_aos: n x {min x, y, z, max x, y, z} in device memory.  Blocking (one small read-back per tree level).
cudaError_t kd_build_device(const double* d_bounds_aos, uint32_t n, const PtKdBuildConfig& cfg, const KdAllocator& al, cudaStream_t st,
                            KdTreeDev** out);
void kd_tree_release(KdTreeDev* t);
uint32_t kd_tree_node_count(const KdTreeDev* t);
uint32_t kd_tree_item_count(const KdTreeDev* t);
uint32_t kd_tree_depth(const KdTreeDev* t);
uint32_t kd_tree_launches(const KdTreeDev* t);
float kd_tree_device_ms(const KdTreeDev* t);
uint64_t kd_tree_algorithmic_bytes(const KdTreeDev* t);
uint32_t kd_tree_tries(const KdTreeDev* t);
const double* kd_tree_root_bounds(const KdTreeDev* t);  // min xyz, max xyz
const PtKdNode* kd_tree_nodes_device(const KdTreeDev* t);
const uint32_t* kd_tree_items_device(const KdTreeDev* t);

}  // namespace ptd
