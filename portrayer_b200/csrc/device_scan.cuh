// Pieces shared by the level-parallel builders (kd_build.cu, flatten.cu): a 3-pass exclusive scan of 64-bit values
// (two 32-bit counters packed as hi << 32 | lo), and growable device buffers backed by the caller's caching allocator.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "kd_build.h"

namespace ptd {
namespace {

constexpr int kB = 256;
constexpr unsigned kFull = 0xFFFFFFFFu;

// ------------------------------------------------------------------ exclusive scan of 64-bit values (3 passes)
// Two 32-bit counters are scanned at once, packed as (hi << 32 | lo); neither half can overflow (< 2^32 members).
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kB * kScanItems;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* total) {
    __shared__ unsigned long long warp_sums[kB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long o = __shfl_up_sync(kFull, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < kB / 32 ? warp_sums[lane] : 0ull;
#pragma unroll
        for (int off = 1; off < kB / 32; off <<= 1) {
            const unsigned long long o = __shfl_up_sync(kFull, w, off);
            if (lane >= off) w += o;
        }
        if (lane < kB / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    const unsigned long long before = warp ? warp_sums[warp - 1] : 0ull;
    if (total) *total = warp_sums[kB / 32 - 1];
    __syncthreads();
    return before + inc - v;
}

template <class Load>
__global__ void __launch_bounds__(kB) scan_reduce_kernel(Load load, uint32_t n, unsigned long long* __restrict__ tile_sums) {
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
        if (base + j < n) sum += load(base + j);
    unsigned long long total;
    block_exclusive_scan(sum, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place, grand total to *total_out
__global__ void __launch_bounds__(kB) scan_tiles_kernel(unsigned long long* __restrict__ tile_sums, uint32_t n_tiles,
                                                        unsigned long long* __restrict__ total_out) {
    unsigned long long carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += kB) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_sums[i] : 0ull;
        unsigned long long total;
        const unsigned long long ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// out[i] = exclusive prefix, out[n] = total
template <class Load>
__global__ void __launch_bounds__(kB) scan_apply_kernel(Load load, uint32_t n, const unsigned long long* __restrict__ tile_sums,
                                                        const unsigned long long* __restrict__ total, unsigned long long* __restrict__ out) {
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned long long v[kScanItems];
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        v[j] = base + j < n ? load(base + j) : 0ull;
        sum += v[j];
    }
    unsigned long long run = tile_sums[blockIdx.x] + block_exclusive_scan(sum, nullptr);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

// ------------------------------------------------------------------ host side
// Device scratch and result buffers come from the caller's caching allocator (api.cu's arena: no cudaMalloc /
// cudaFree — which synchronise the device — once a process has built a tree of a similar size).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    const KdAllocator* al = nullptr;
    cudaError_t reserve(size_t bytes, bool keep, cudaStream_t st) {
        if (bytes <= cap) return cudaSuccess;
        const size_t want = std::max(bytes, cap * 2);
        cudaError_t e = cudaSuccess;
        void* q = al->alloc(want, &e);
        if (!q) return e == cudaSuccess ? cudaErrorMemoryAllocation : e;
        if (keep && p && cap) {
            e = cudaMemcpyAsync(q, p, cap, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        if (p) al->release(p);
        p = q;
        cap = want;
        return e;
    }
    void release() {
        if (p) al->release(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

inline uint32_t blocks(uint64_t n, uint32_t per = kB) { return (uint32_t)((n + per - 1) / per); }

template <class Load>
cudaError_t exclusive_scan(Load load, uint32_t n, DevBuf& tiles, unsigned long long* d_total, unsigned long long* out, cudaStream_t st) {
    const uint32_t n_tiles = std::max<uint32_t>(1, blocks(n, kScanTile));
    cudaError_t e = tiles.reserve((size_t)n_tiles * 8, false, st);
    if (e != cudaSuccess) return e;
    scan_reduce_kernel<<<n_tiles, kB, 0, st>>>(load, n, tiles.as<unsigned long long>());
    scan_tiles_kernel<<<1, kB, 0, st>>>(tiles.as<unsigned long long>(), n_tiles, d_total);
    scan_apply_kernel<<<n_tiles, kB, 0, st>>>(load, n, tiles.as<unsigned long long>(), d_total, out);
    return cudaGetLastError();
}

}  // namespace
}  // namespace ptd
