// Host-callable launch wrappers of kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include "device_scene.cuh"

namespace ptd {
void kernels_init();
void launch_begin_batch(BatchCtl* ctl, uint32_t n_paths, cudaStream_t st);
void launch_camera(const FrameParams& fp, const NodePool& pool, uint32_t first_slot, uint32_t n_paths, cudaStream_t st);
void launch_load_rays(const double* origins, const double* dirs, const NodePool& pool, uint32_t first, uint32_t n_paths,
                      cudaStream_t st);
void launch_extend(const DScene& sc, const NodePool& pool, BatchCtl* ctl, int level, uint32_t max_items, bool count,
                   cudaStream_t st);
void launch_shadow(const DScene& sc, const FrameParams& fp, const NodePool& pool, BatchCtl* ctl, int level,
                   uint32_t first_slot, uint32_t max_items, bool count, cudaStream_t st);
void launch_shade(const DScene& sc, const FrameParams& fp, const NodePool& pool, BatchCtl* ctl, int level,
                  uint32_t first_slot, uint32_t max_items, cudaStream_t st);
void launch_tree_eval(const FrameParams& fp, const NodePool& pool, uint32_t first_slot, uint32_t n_paths, cudaStream_t st);
void launch_resolve(const FrameParams& fp, const NodePool& pool, uint32_t first_slot, uint32_t n_slots, uint8_t* rgb,
                    uint32_t* hit_id, double* hit_t, cudaStream_t st);
void launch_export_rays(const NodePool& pool, uint32_t first, uint32_t n_paths, double* color, uint32_t* hit_id,
                        double* hit_t, cudaStream_t st);
}  // namespace ptd
