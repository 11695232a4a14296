// Spatial order of the triangles of linear `Mesh`es, for the fold cull of traverse.cuh (mesh_fold).
//
// Mesh::ray_hit folds over EVERY triangle of the mesh in index order (src/primitive/mesh.rs:157-167).  The result of
// that fold is the lexicographic minimum of (t, index) over the triangles whose own test passes, so it can be
// evaluated in ANY order; this file picks an order in which consecutive triangles are neighbours in space (Morton
// order of the centroids inside the mesh's box), so that the boxes of aligned runs of 4, 16, 64 ... positions
// (kernels.cu group_bounds_kernel) form a tight implicit 4-ary hierarchy whatever order the OBJ file listed them in.
//
//   keys:  (tri_first of the mesh) << 32 | morton30(centroid)      one 64-bit radix sort for the whole scene
//   order: position -> triangle index; a mesh's positions stay inside its own [tri_first, tri_first + tri_count)
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include "kernels.h"

namespace ptd {
namespace {
constexpr int kBlock = 128;

__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// every triangle keeps its place unless a linear mesh claims it below
__global__ void __launch_bounds__(kBlock) fold_keys_identity_kernel(uint32_t n, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    keys[g] = (unsigned long long)g << 32;
    vals[g] = g;
}

// blockIdx.y walks the meshes, x their triangles
__global__ void __launch_bounds__(kBlock) fold_keys_kernel(const PtMesh* __restrict__ meshes, uint32_t n_meshes, const PtTriPos* __restrict__ tri_pos,
                                                          const double* __restrict__ mesh_bounds /* [n_meshes][6] */, uint32_t n_triangles,
                                                          unsigned long long* __restrict__ keys) {
    for (uint32_t m = blockIdx.y; m < n_meshes; m += gridDim.y) {
        const uint32_t first = meshes[m].tri_first, count = meshes[m].tri_count;
        const bool linear = meshes[m].kind == PT_MESH_LINEAR;
        double lo[3], inv[3];
        for (int c = 0; c < 3; ++c) {
            lo[c] = mesh_bounds[m * 6 + c];
            const double ext = mesh_bounds[m * 6 + 3 + c] - lo[c];
            inv[c] = ext > 0.0 ? 1023.0 / ext : 0.0;
        }
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
            if (first + k >= n_triangles) break;
            uint32_t low = k;  // KDMesh ranges keep their index order (the fold structure is not used for them)
            if (linear) {
                const double* v = reinterpret_cast<const double*>(tri_pos + first + k);
                uint32_t q[3];
                for (int c = 0; c < 3; ++c) {
                    const double x = ((v[c] + v[3 + c] + v[6 + c]) / 3.0 - lo[c]) * inv[c];
                    q[c] = x >= 0.0 ? (x < 1023.0 ? (uint32_t)x : 1023u) : 0u;  // NaN -> 0
                }
                low = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
            }
            keys[first + k] = ((unsigned long long)first << 32) | low;
        }
    }
}

__global__ void __launch_bounds__(kBlock) gather_fold_boxes_kernel(const float4* __restrict__ tri_aabb, const uint32_t* __restrict__ order, uint32_t n,
                                                                  float4* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t k = order[j];
    out[2 * (size_t)j] = tri_aabb[2 * (size_t)k];
    out[2 * (size_t)j + 1] = tri_aabb[2 * (size_t)k + 1];
}
}  // namespace

size_t fold_sort_temp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 64);
    return bytes;
}

// order[n]: position -> triangle.  `sort` false: identity (scenes without a linear mesh worth sorting).
// scratch: 2 * n u64 keys + n u32 values + temp_bytes of cub storage, laid out by the caller as below.
cudaError_t launch_fold_order(const PtMesh* meshes, uint32_t n_meshes, const PtTriPos* tri_pos, const double* mesh_bounds, uint32_t n, bool sort,
                              uint32_t* order, void* scratch, size_t temp_bytes, int end_bit, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const uint32_t blocks = (n + kBlock - 1) / kBlock;
    if (!sort || !n_meshes) {
        // keys are not needed: order = identity
        unsigned long long* keys = static_cast<unsigned long long*>(scratch);
        fold_keys_identity_kernel<<<blocks, kBlock, 0, st>>>(n, keys, order);
        return cudaGetLastError();
    }
    unsigned long long* keys_in = static_cast<unsigned long long*>(scratch);
    unsigned long long* keys_out = keys_in + n;
    uint32_t* vals_in = reinterpret_cast<uint32_t*>(keys_out + n);
    void* temp = reinterpret_cast<unsigned char*>(vals_in + n + (n & 1u));  // 8-byte aligned
    fold_keys_identity_kernel<<<blocks, kBlock, 0, st>>>(n, keys_in, vals_in);
    const dim3 grid(blocks < 1024u ? blocks : 1024u, n_meshes < 1024u ? n_meshes : 1024u);
    fold_keys_kernel<<<grid, kBlock, 0, st>>>(meshes, n_meshes, tri_pos, mesh_bounds, n, keys_in);
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, order, (int)n, 0, end_bit, st);
}

size_t fold_scratch_bytes(uint32_t n, size_t temp_bytes) {
    return (size_t)n * 16 + ((size_t)n + (n & 1u)) * 4 + temp_bytes + 16;
}

void launch_gather_fold_boxes(const float4* tri_aabb, const uint32_t* order, uint32_t n, float4* out, cudaStream_t st) {
    if (!n) return;
    gather_fold_boxes_kernel<<<(n + kBlock - 1) / kBlock, kBlock, 0, st>>>(tri_aabb, order, n, out);
}

}  // namespace ptd
