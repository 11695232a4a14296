// Leaf cull structure of the k-d trees (scene tree and KDMesh trees), built at upload time.
//
// The reference's trees are poor spatial indices: KDLeaf::partitioned (src/kdtree/leaf.rs:89-231) puts an item into
// BOTH children whenever its bounds straddle the plane and stops at depth 10, so on examples/graphics-castle every
// instance sits in 9 leaves on average, a leaf lists 62 candidates, and a primary ray walks through 10 leaves = 650
// candidates (tools/leafstats).  The walk itself (src/kdtree/node.rs:66-203) has to be reproduced as it is — the
// order of the leaves and the [s, e) ranges they are visited with decide which hit wins — but WHICH candidates of a
// leaf can possibly return a hit inside the leaf's range is a geometric question with a much better answer than the
// leaf list: a hit at parameter t in [s, e) lies (up to the EPSILON windows at the ends of the ancestors' ranges, see
// traverse.cuh `kClipPadT`) inside the leaf's own CELL — the intersection of the half spaces its ancestors' planes
// selected.  So every candidate's padded FP32 box is CLIPPED to the cell of the leaf that lists it, and on top of the
// clipped boxes sit the union box of every run of 8 list positions and one "occupied" box per leaf:
//
//   occ[rank]                  union of the leaf's clipped boxes        1 slab test dismisses 89 % of the castle's leaf visits
//   grp[gbase + g]             union of positions 8g .. 8g+7 of the leaf
//   item[(gbase + g) * 8 + j]  clipped box of position 8g + j           (empty box in the padding slots of the last run)
//
// graphics-castle, FP32 slab tests per primary ray: 105 + 49 (scene tree + KDMesh trees) -> 28 + 10.
// The same three arrays exist a second time with UNCLIPPED boxes ("full" set, `set_stride` float4 further on): a ray
// whose probe segment node.rs:119-122 may end INSIDE the tree's bounds (extent is the SQUARED diagonal, so this
// happens for scenes smaller than 1 unit, or for a KDMesh instance scaled up so far that the object-space extent is
// short in world-ray parameters) can find hits outside the cell it was sent to, and uses the full set instead.
//
// A leaf's record in the device copy of the node array is patched to carry where its boxes are: the `split` field
// (unused for leaves, portrayer_gpu.h PtKdNode) becomes rank | gbase << 32.
//
//   node[i]                    union of the occupied boxes of every leaf below node i (a leaf's own occupied box for a leaf)
//
// With the node boxes the walk can PRUNE: a subtree whose box the ray misses inside the range it would be visited with
// holds no leaf that can return a hit, so the walk need not descend into it (traverse.cuh kd_walk, PRUNE).  The reference
// descends anyway and finds nothing; on graphics-castle that is most of what a ray does (52 split steps per ray, a
// third of them below the last subtree that holds anything near the ray).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "device_scan.cuh"
#include "kernels.h"

namespace ptd {
namespace {

constexpr uint32_t kNoParent = 0xFFFFFFFFu;

// parent links (child -> parent << 1 | is_front_child) and the scan input (leaf: 1 << 32 | runs of 8) of every node
__global__ void __launch_bounds__(kB) lc_parent_kernel(const PtKdNode* __restrict__ nodes, const LcTree* __restrict__ trees, uint32_t n_trees,
                                                       uint32_t* __restrict__ parent, unsigned long long* __restrict__ scan_in) {
    for (uint32_t t = blockIdx.y; t < n_trees; t += gridDim.y) {
        const LcTree tr = trees[t];
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < tr.node_count; j += gridDim.x * blockDim.x) {
            const uint32_t i = tr.node_first + j;
            const PtKdNode n = nodes[i];
            if ((n.a & 3u) != 3u) {
                const uint32_t front = n.a >> 2, back = n.b;
                if (front < tr.node_count) parent[tr.node_first + front] = i << 1 | 1u;
                if (back < tr.node_count) parent[tr.node_first + back] = i << 1;
                scan_in[i] = 0ull;
            } else {
                scan_in[i] = 1ull << 32 | (unsigned long long)((n.b + 7u) >> 3);
            }
        }
    }
}

struct ScanLoad {
    const unsigned long long* in;
    __device__ unsigned long long operator()(uint32_t i) const { return in[i]; }
};

struct Box3 {
    float lo[3], hi[3];
};
__device__ __forceinline__ Box3 empty_box() { return Box3{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; }
__device__ __forceinline__ void grow(Box3& a, const Box3& b) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { a.lo[c] = fminf(a.lo[c], b.lo[c]); a.hi[c] = fmaxf(a.hi[c], b.hi[c]); }
}
__device__ __forceinline__ Box3 shfl_xor_box(const Box3& b, int mask) {
    Box3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c) { r.lo[c] = __shfl_xor_sync(kFull, b.lo[c], mask); r.hi[c] = __shfl_xor_sync(kFull, b.hi[c], mask); }
    return r;
}
__device__ __forceinline__ void store_box(float4* __restrict__ dst, size_t idx, const Box3& b) {
    dst[2 * idx] = make_float4(b.lo[0], b.lo[1], b.lo[2], 0.f);
    dst[2 * idx + 1] = make_float4(b.hi[0], b.hi[1], b.hi[2], 0.f);
}

// one warp per node; leaves build their boxes
__global__ void __launch_bounds__(kB) lc_build_kernel(PtKdNode* __restrict__ nodes, const uint32_t* __restrict__ items, const LcTree* __restrict__ trees,
                                                      uint32_t n_trees, const uint32_t* __restrict__ parent, const unsigned long long* __restrict__ scan,
                                                      const float4* __restrict__ item_boxes, LeafCull out, uint32_t n_leaf_cap, uint32_t n_grp_cap) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps_per_grid = gridDim.x * (kB / 32);
    for (uint32_t t = blockIdx.y; t < n_trees; t += gridDim.y) {
        const LcTree tr = trees[t];
        for (uint32_t j = blockIdx.x * (kB / 32) + (threadIdx.x >> 5); j < tr.node_count; j += warps_per_grid) {
            const uint32_t i = tr.node_first + j;
            const PtKdNode n = nodes[i];
            if ((n.a & 3u) != 3u) continue;
            const uint32_t first = n.a >> 2, count = n.b;
            const unsigned long long sc = scan[i];
            const uint32_t rank = (uint32_t)(sc >> 32), gbase = (uint32_t)sc;
            const uint32_t n_groups = (count + 7u) >> 3;
            if (rank >= n_leaf_cap || gbase + n_groups > n_grp_cap) continue;  // cannot happen for a well-formed forest (capacities are upper bounds)
            // the leaf's cell: every ancestor's plane keeps one half space (front child: coordinate >= split, infinite_plane.rs:27-35)
            double clo[3] = {-INFINITY, -INFINITY, -INFINITY}, chi[3] = {INFINITY, INFINITY, INFINITY};
            uint32_t p = parent[i];
            for (int guard = 0; p != kNoParent && guard < 4096; ++guard) {
                const uint32_t a = p >> 1;
                const PtKdNode an = nodes[a];
                const uint32_t axis = an.a & 3u;
                if (axis < 3u) {
                    if (p & 1u) clo[axis] = fmax(clo[axis], an.split);
                    else chi[axis] = fmin(chi[axis], an.split);
                }
                p = parent[a];
            }
            Box3 occ_clip = empty_box(), occ_full = empty_box();
            for (uint32_t base = 0; base < n_groups * 8u; base += 32u) {
                const uint32_t k = base + lane;
                Box3 full = empty_box(), clip = empty_box();
                if (k < count) {
                    const size_t idx = (size_t)tr.box_first + items[tr.item_first + first + k];
                    const float4 lo = item_boxes[2 * idx], hi = item_boxes[2 * idx + 1];
                    full = Box3{{lo.x, lo.y, lo.z}, {hi.x, hi.y, hi.z}};
                    clip = full;
                    const float diag = fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
                    bool empty = false;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        // cell planes rounded outward and padded like the boxes themselves (kernels.cu instance_bounds_kernel)
                        const double pl = clo[c] - (1e-5 * fabs(clo[c]) + 1e-5 * (double)diag + 1e-30);
                        const double ph = chi[c] + (1e-5 * fabs(chi[c]) + 1e-5 * (double)diag + 1e-30);
                        const float fl = isfinite(clo[c]) ? __double2float_rd(pl) : -INFINITY;
                        const float fh = isfinite(chi[c]) ? __double2float_ru(ph) : INFINITY;
                        clip.lo[c] = fmaxf(full.lo[c], fl);
                        clip.hi[c] = fminf(full.hi[c], fh);
                        if (!(clip.lo[c] <= clip.hi[c])) empty = true;
                    }
                    // a listed item whose box misses the cell (the reference's bounds() and these padded FP32 boxes differ
                    // in the last digits) or is not finite keeps its full box: slab tests have no encoding for "empty"
                    if (empty) clip = full;
                }
                // padding slots of the leaf's last run of 8 repeat the leaf's last box (the walk masks them out by count)
                if (base + 32u > count) {
                    const int src = (int)((count - 1u) & 31u);
                    const Box3 lc = Box3{{__shfl_sync(kFull, clip.lo[0], src), __shfl_sync(kFull, clip.lo[1], src), __shfl_sync(kFull, clip.lo[2], src)},
                                         {__shfl_sync(kFull, clip.hi[0], src), __shfl_sync(kFull, clip.hi[1], src), __shfl_sync(kFull, clip.hi[2], src)}};
                    const Box3 lf = Box3{{__shfl_sync(kFull, full.lo[0], src), __shfl_sync(kFull, full.lo[1], src), __shfl_sync(kFull, full.lo[2], src)},
                                         {__shfl_sync(kFull, full.hi[0], src), __shfl_sync(kFull, full.hi[1], src), __shfl_sync(kFull, full.hi[2], src)}};
                    if (k >= count) { clip = lc; full = lf; }
                }
                if (k < n_groups * 8u) {
                    store_box(out.item, (size_t)gbase * 8 + k, clip);
                    store_box(out.item + out.set_stride, (size_t)gbase * 8 + k, full);
                }
                grow(occ_clip, clip);
                grow(occ_full, full);
                Box3 gc = clip, gf = full;
#pragma unroll
                for (int m = 1; m < 8; m <<= 1) {
                    const Box3 oc = shfl_xor_box(gc, m), of = shfl_xor_box(gf, m);
                    grow(gc, oc);
                    grow(gf, of);
                }
                if ((lane & 7) == 0 && k < n_groups * 8u) {
                    store_box(out.grp, (size_t)gbase + (k >> 3), gc);
                    store_box(out.grp + out.set_stride, (size_t)gbase + (k >> 3), gf);
                }
            }
#pragma unroll
            for (int m = 1; m < 32; m <<= 1) {
                const Box3 oc = shfl_xor_box(occ_clip, m), of = shfl_xor_box(occ_full, m);
                grow(occ_clip, oc);
                grow(occ_full, of);
            }
            if (lane == 0) {
                store_box(out.occ, rank, occ_clip);
                store_box(out.occ + out.set_stride, rank, occ_full);
                store_box(out.node, i, occ_clip);
                store_box(out.node + out.set_stride, i, occ_full);
                nodes[i].split = __longlong_as_double((long long)((unsigned long long)rank | (unsigned long long)gbase << 32));
            }
        }
    }
}

// node boxes of the inner nodes: empty first, then `depth` sweeps of box[i] = box[front] U box[back] (children lie after
// their parents in the arrays, so sweep k settles every node whose subtree is at most k levels high)
__global__ void __launch_bounds__(kB) lc_node_init_kernel(const PtKdNode* __restrict__ nodes, uint32_t n_nodes, LeafCull out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        if ((nodes[i].a & 3u) == 3u && nodes[i].b != 0u) continue;  // a leaf with candidates: written by lc_build_kernel
        store_box(out.node, i, empty_box());
        store_box(out.node + out.set_stride, i, empty_box());
    }
}
__global__ void __launch_bounds__(kB) lc_node_union_kernel(const PtKdNode* __restrict__ nodes, const LcTree* __restrict__ trees, uint32_t n_trees,
                                                            LeafCull out) {
    for (uint32_t t = blockIdx.y; t < n_trees; t += gridDim.y) {
        const LcTree tr = trees[t];
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < tr.node_count; j += gridDim.x * blockDim.x) {
            const uint32_t i = tr.node_first + j;
            const PtKdNode n = nodes[i];
            if ((n.a & 3u) == 3u) continue;
            const uint32_t front = n.a >> 2, back = n.b;
            if (front >= tr.node_count || back >= tr.node_count) continue;
#pragma unroll
            for (int set = 0; set < 2; ++set) {
                float4* __restrict__ nb = out.node + (size_t)set * out.set_stride;
                const float4 fl = nb[2 * (size_t)(tr.node_first + front)], fh = nb[2 * (size_t)(tr.node_first + front) + 1];
                const float4 bl = nb[2 * (size_t)(tr.node_first + back)], bh = nb[2 * (size_t)(tr.node_first + back) + 1];
                nb[2 * (size_t)i] = make_float4(fminf(fl.x, bl.x), fminf(fl.y, bl.y), fminf(fl.z, bl.z), 0.f);
                nb[2 * (size_t)i + 1] = make_float4(fmaxf(fh.x, bh.x), fmaxf(fh.y, bh.y), fmaxf(fh.z, bh.z), 0.f);
            }
        }
    }
}

// union of n boxes -> out[0..1]; one block
__global__ void __launch_bounds__(kB) lc_root_kernel(const float4* __restrict__ boxes, uint32_t n, float4* __restrict__ out) {
    Box3 b = empty_box();
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        const float4 lo = boxes[2 * (size_t)k], hi = boxes[2 * (size_t)k + 1];
        const Box3 o{{lo.x, lo.y, lo.z}, {hi.x, hi.y, hi.z}};
        grow(b, o);
    }
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const Box3 o = shfl_xor_box(b, m);
        grow(b, o);
    }
    __shared__ Box3 s[kB / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kB / 32; ++w) grow(b, s[w]);
        out[0] = make_float4(b.lo[0], b.lo[1], b.lo[2], 0.f);
        out[1] = make_float4(b.hi[0], b.hi[1], b.hi[2], 0.f);
    }
}

}  // namespace

LeafCullSizes leaf_cull_sizes(uint32_t n_nodes, uint32_t n_items) {
    LeafCullSizes s;
    s.n_leaf_cap = std::max<uint32_t>(n_nodes, 1);
    s.n_grp_cap = (uint32_t)std::min<uint64_t>((uint64_t)n_items / 8 + n_nodes + 1, 0x0FFFFFFFull);
    s.set_float4 = 2 * ((size_t)s.n_leaf_cap + (size_t)s.n_grp_cap * 9 + (size_t)std::max<uint32_t>(n_nodes, 1));
    s.scratch_bytes = (size_t)std::max<uint32_t>(n_nodes, 1) * 20 + 128;
    return s;
}

// storage: 2 * set_float4 float4 (clipped set, then full set); scratch: leaf_cull_sizes().scratch_bytes + the scan's tile buffer
cudaError_t launch_leaf_cull(PtKdNode* nodes, uint32_t n_nodes, const uint32_t* items, uint32_t n_items, const LcTree* d_trees, uint32_t n_trees,
                             uint32_t max_depth, const float4* item_boxes, float4* storage, void* scratch, const KdAllocator& al, LeafCull* out,
                             cudaStream_t st) {
    const LeafCullSizes sz = leaf_cull_sizes(n_nodes, n_items);
    out->occ = storage;
    out->grp = storage + 2 * (size_t)sz.n_leaf_cap;
    out->item = out->grp + 2 * (size_t)sz.n_grp_cap;
    out->node = out->item + 16 * (size_t)sz.n_grp_cap;
    out->set_stride = (uint32_t)sz.set_float4;
    if (!n_nodes || !n_trees) return cudaSuccess;
    uint32_t* parent = static_cast<uint32_t*>(scratch);
    unsigned long long* scan_in = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(scratch) + (((size_t)n_nodes * 4 + 15) & ~size_t(15)));
    unsigned long long* scan_out = scan_in + n_nodes;  // n_nodes + 1 entries
    cudaError_t e = cudaMemsetAsync(parent, 0xFF, (size_t)n_nodes * 4, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(scan_in, 0, (size_t)n_nodes * 8, st);
    if (e != cudaSuccess) return e;
    const dim3 grid(std::min<uint32_t>(blocks(n_nodes), 1024u), std::min<uint32_t>(n_trees, 1024u));
    lc_parent_kernel<<<grid, kB, 0, st>>>(nodes, d_trees, n_trees, parent, scan_in);
    if (getenv("PT_DEBUG_SYNC")) fprintf(stderr, "[pt] lc_parent: %s\n", cudaGetErrorString(cudaStreamSynchronize(st)));
    DevBuf tiles;
    tiles.al = &al;
    unsigned long long* d_total = scan_out + n_nodes;  // exclusive_scan writes out[n] = total itself; the tile kernel needs a separate word
    e = exclusive_scan(ScanLoad{scan_in}, n_nodes, tiles, d_total + 1, scan_out, st);
    if (e != cudaSuccess) { tiles.release(); return e; }
    if (getenv("PT_DEBUG_SYNC")) fprintf(stderr, "[pt] lc scan: %s\n", cudaGetErrorString(cudaStreamSynchronize(st)));
    const dim3 grid_b(std::min<uint32_t>(blocks((uint64_t)n_nodes * 32), 4096u), std::min<uint32_t>(n_trees, 1024u));
    lc_node_init_kernel<<<std::min<uint32_t>(blocks(n_nodes), 4096u), kB, 0, st>>>(nodes, n_nodes, *out);
    lc_build_kernel<<<grid_b, kB, 0, st>>>(nodes, items, d_trees, n_trees, parent, scan_out, item_boxes, *out, sz.n_leaf_cap, sz.n_grp_cap);
    tiles.release();  // stream-ordered reuse
    for (uint32_t sweep = 0; sweep < max_depth; ++sweep) lc_node_union_kernel<<<grid, kB, 0, st>>>(nodes, d_trees, n_trees, *out);
    return cudaGetLastError();
}

void launch_root_box(const float4* boxes, uint32_t n, float4* out2, cudaStream_t st) { lc_root_kernel<<<1, kB, 0, st>>>(boxes, n, out2); }

}  // namespace ptd
