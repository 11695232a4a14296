// Device scene flattening (flatten.cu); C ABI wrappers live in api.cu (pt_flatten*).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kd_build.h"
#include "portrayer_gpu.h"

namespace ptd {

struct FlatSceneDev;

// All pointers are device memory.  Blocking (one small read-back per hierarchy level).  max_levels bounds the walk
// (a hierarchy deeper than that is taken for a cycle -> cudaErrorInvalidValue).
cudaError_t flatten_device(const PtHierNode* d_nodes, uint32_t n_nodes, const uint32_t* d_children, uint32_t n_children, uint32_t root,
                           const PtGeometryRec* d_geoms, uint32_t max_levels, const KdAllocator& al, cudaStream_t st,
                           FlatSceneDev** out);
void flat_release(FlatSceneDev* f);
uint32_t flat_instance_count(const FlatSceneDev* f);
uint32_t flat_depth(const FlatSceneDev* f);
uint32_t flat_launches(const FlatSceneDev* f);
float flat_device_ms(const FlatSceneDev* f);
const PtInstance* flat_instances_device(const FlatSceneDev* f);
const PtInstanceTrans* flat_trans_device(const FlatSceneDev* f);
const double* flat_bounds_device(const FlatSceneDev* f);  // n x {min xyz, max xyz}

}  // namespace ptd
