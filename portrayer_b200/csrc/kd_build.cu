// k-d tree BUILD on the device — SURVEY §8(f) rank 1, the step immediately before the render loop.
//
//   KDLeaf::partitioned          src/kdtree/leaf.rs:89-231   (the algorithm: bisection of the split plane, <= max_tries)
//   partition_node / which_side  src/kdtree/leaf.rs:114-131, src/primitive/infinite_plane.rs:27-35
//   Bounds for Vec<T>            src/bounding_box.rs:23-37   (child bounds = min / max over the members' bounds)
//   KDTreeScene::from            src/kdtree/kdscene.rs:19-44 (TLAS: items = flat instances, depth KD_DEPTH)
//   KDMesh::new                  src/kdtree/kdmesh.rs:37-58  (BLAS: items = triangles, depth KD_MESH_DEPTH)
//
// The reference recurses depth-first and re-partitions an Arc list per node.  Every decision it takes depends
// only on (a) the node's member list in its stable order and (b) min / max of the members' bounds along the split
// axis, so the same tree is built here breadth-first, one LEVEL per round, all nodes of the level at once:
//
//   bounds   min / max of every node's members along the level's axis (segmented warp reduction + ordered-key atomics)
//   tries    <= max_tries rounds of { classify every member against its node's plane, count front / back / shared
//            (warp-aggregated integer atomics) ; per node: merit = |front - back| + shared, stop or bisect }
//   split    classify against the final plane, exclusive scan of the (goes-front, goes-back) flags over the level,
//            stable scatter into the next level's member list (shared members go to both children), emit the
//            level's PtKdNode records and the member lists of its leaves
//
// Counts are integers and min / max are exact, so the result does not depend on the order of the atomics: nodes
// and leaf lists come out in exactly the breadth-first order of the host packer (front child, then back child, per
// parent in order), and the arrays are bit-identical to the host build's (tests/test_kd_build.py), up to the sign of
// a zero bound (min / max of -0.0 and +0.0 is order-dependent in the reference's fold; no comparison or quotient
// downstream can tell them apart).  All of it is integer / compare work bound by HBM: per level and try one read of
// 2 x 8 B of bounds + 8 B of (member, owner) per member.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "device_scan.cuh"
#include "kd_build.h"

namespace ptd {
namespace {

// ------------------------------------------------------------------ ordered keys for atomic min / max of doubles
__device__ __forceinline__ unsigned long long key_of(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// per-node state of the level being partitioned
struct LevelState {
    const uint32_t* seg;       // [K + 1] first member of every node in `items`
    unsigned long long* klo;   // [K] ordered key of min over members of bounds.min[axis]
    unsigned long long* khi;   // [K] ordered key of max over members of bounds.max[axis]
    double* plane;             // [K] sep_plane.point[axis]
    double* pmin;              // [K] plane_range
    double* pmax;
    uint32_t* counts;          // [K][4] front, back, shared, (unused)
    uint8_t* status;           // [K] 0 = leaf, 1 = still searching, 2 = plane fixed
    uint32_t* n_searching;     // [1] nodes with status 1
    uint32_t* tries_used;      // [1] tries of this level in which at least one node was still searching
};

__global__ void transpose_bounds_kernel(const double* __restrict__ aos, uint32_t n, double* __restrict__ soa) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 6; ++c) soa[(size_t)c * n + i] = aos[(size_t)i * 6 + c];
}

__global__ void iota_kernel(uint32_t* __restrict__ items, uint32_t* __restrict__ owner, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    items[i] = i;
    owner[i] = 0u;
}

__global__ void init_keys_kernel(unsigned long long* __restrict__ klo, unsigned long long* __restrict__ khi, uint32_t k_nodes) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_nodes) return;
    klo[k] = ~0ull;
    khi[k] = 0ull;
}

// min / max over every node's members along one axis.  Members of a node are contiguous, so a warp first reduces
// runs of equal owner with shuffles; the head of every run issues the two atomics.
__global__ void __launch_bounds__(kB) bounds_kernel(const uint32_t* __restrict__ items, const uint32_t* __restrict__ owner, uint32_t m,
                                                    const double* __restrict__ bmin, const double* __restrict__ bmax,
                                                    unsigned long long* __restrict__ klo, unsigned long long* __restrict__ khi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < m;
    uint32_t k = 0xFFFFFFFFu;
    unsigned long long lo = ~0ull, hi = 0ull;
    if (valid) {
        const uint32_t it = items[i];
        k = owner[i];
        lo = key_of(__ldg(bmin + it));
        hi = key_of(__ldg(bmax + it));
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long olo = __shfl_down_sync(kFull, lo, off), ohi = __shfl_down_sync(kFull, hi, off);
        const uint32_t ok = __shfl_down_sync(kFull, k, off);
        if (lane + off < 32 && ok == k) {
            lo = olo < lo ? olo : lo;
            hi = ohi > hi ? ohi : hi;
        }
    }
    const uint32_t prev = __shfl_up_sync(kFull, k, 1);
    if (valid && (lane == 0 || prev != k)) {
        atomicMin(klo + k, lo);
        atomicMax(khi + k, hi);
    }
}

// leaf.rs:91-93 (leaf test) and :135-148 (first plane = centre of the node's bounds along the axis)
__global__ void node_init_kernel(LevelState st, uint32_t k_nodes, uint32_t depth_left, uint32_t target_max_nodes) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_nodes) return;
    const uint32_t count = st.seg[k + 1] - st.seg[k];
    st.counts[4 * k + 0] = st.counts[4 * k + 1] = st.counts[4 * k + 2] = st.counts[4 * k + 3] = 0u;
    if (depth_left == 0 || count <= target_max_nodes) {
        st.status[k] = 0;
        return;
    }
    const double min_axis = value_of(st.klo[k]), max_axis = value_of(st.khi[k]);
    st.plane[k] = min_axis + (max_axis - min_axis) / 2.0;
    st.pmin[k] = min_axis;
    st.pmax[k] = max_axis;
    st.status[k] = 1;
    atomicAdd(st.n_searching, 1u);
}

// partition_node, leaf.rs:114-131: 0 front, 1 back, 2 shared
__device__ __forceinline__ uint32_t classify(double lo, double hi, double plane) {
    const bool a = (lo - plane) >= 0.0;  // which_side(node_min), infinite_plane.rs:27-35 with an axis-unit normal
    const bool b = (hi - plane) >= 0.0;  // which_side(node_max)
    return a == b ? (a ? 0u : 1u) : 2u;
}

// One try of leaf.rs:156-175 for every node still searching (ALL = false), or the final partition count
// against the chosen planes for every split node (ALL = true).
template <bool ALL>
__global__ void __launch_bounds__(kB) count_kernel(const uint32_t* __restrict__ items, const uint32_t* __restrict__ owner, uint32_t m,
                                                   const double* __restrict__ bmin, const double* __restrict__ bmax, LevelState st) {
    if (!ALL && *st.n_searching == 0u) return;
    if (!ALL && blockIdx.x == 0 && threadIdx.x == 0) ++*st.tries_used;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t k = 0xFFFFFFFFu, cls = 3u;
    if (i < m) {
        k = owner[i];
        const uint8_t s = st.status[k];
        if (ALL ? s != 0 : s == 1) {
            const uint32_t it = items[i];
            cls = classify(__ldg(bmin + it), __ldg(bmax + it), st.plane[k]);
        }
    }
    // warp-aggregated counting: one atomic per (run of equal owner, class)
    const unsigned peers = __match_any_sync(kFull, k);
    const unsigned front = __ballot_sync(kFull, cls == 0u) & peers;
    const unsigned back = __ballot_sync(kFull, cls == 1u) & peers;
    const unsigned shared = __ballot_sync(kFull, cls == 2u) & peers;
    if (cls != 3u && lane == __ffs(peers & (front | back | shared)) - 1) {
        if (front) atomicAdd(st.counts + 4 * k + 0, (uint32_t)__popc(front));
        if (back) atomicAdd(st.counts + 4 * k + 1, (uint32_t)__popc(back));
        if (shared) atomicAdd(st.counts + 4 * k + 2, (uint32_t)__popc(shared));
    }
}

// leaf.rs:176-200: merit test, then move the plane to the middle of the half of its range that holds more nodes
__global__ void try_update_kernel(LevelState st, uint32_t k_nodes, int target_max_merit) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_nodes || st.status[k] != 1) return;
    const long long front = st.counts[4 * k + 0], back = st.counts[4 * k + 1], shared = st.counts[4 * k + 2];
    st.counts[4 * k + 0] = st.counts[4 * k + 1] = st.counts[4 * k + 2] = 0u;
    const long long diff = front - back;
    const long long merit = (diff < 0 ? -diff : diff) + shared;
    if (merit <= (long long)target_max_merit) {
        st.status[k] = 2;
        atomicSub(st.n_searching, 1u);
        return;
    }
    const double plane = st.plane[k];
    if (front > back) {
        const double plane_max = st.pmax[k];
        st.pmin[k] = plane;
        st.plane[k] = plane + (plane_max - plane) / 2.0;
    } else {
        const double plane_min = st.pmin[k];
        st.pmax[k] = plane;
        st.plane[k] = plane_min + (plane - plane_min) / 2.0;
    }
}

// the final classification of every member (3 = member of a leaf node), kept for the scan and the scatter
__global__ void __launch_bounds__(kB) final_class_kernel(const uint32_t* __restrict__ items, const uint32_t* __restrict__ owner, uint32_t m,
                                                         const double* __restrict__ bmin, const double* __restrict__ bmax,
                                                         const double* __restrict__ plane, const uint8_t* __restrict__ status,
                                                         uint8_t* __restrict__ cls_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t k = owner[i];
    uint8_t cls = 3;
    if (status[k] != 0) {
        const uint32_t it = items[i];
        cls = (uint8_t)classify(__ldg(bmin + it), __ldg(bmax + it), plane[k]);
    }
    cls_out[i] = cls;
}

struct ClassLoad {  // member -> (goes to the back child, goes to the front child)
    const uint8_t* cls;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const {
        const uint8_t c = cls[i];
        const unsigned long long f = (c == 0 || c == 2) ? 1ull : 0ull, b = (c == 1 || c == 2) ? 1ull : 0ull;
        return (b << 32) | f;
    }
};
struct NodeLoad {  // node -> (is split, members of both children) ; leaf nodes contribute nothing
    const uint8_t* status;
    const uint32_t* counts;
    __device__ __forceinline__ unsigned long long operator()(uint32_t k) const {
        if (status[k] == 0) return 0ull;
        const unsigned long long front = counts[4 * k + 0], back = counts[4 * k + 1], shared = counts[4 * k + 2];
        return (1ull << 32) | (front + back + 2ull * shared);
    }
};
struct LeafLoad {  // node -> members of the node if it is a leaf
    const uint8_t* status;
    const uint32_t* seg;
    __device__ __forceinline__ unsigned long long operator()(uint32_t k) const {
        return status[k] == 0 ? (unsigned long long)(seg[k + 1] - seg[k]) : 0ull;
    }
};

// ------------------------------------------------------------------ emit the level
struct EmitArgs {
    const uint32_t* seg;                 // [K + 1]
    const uint8_t* status;               // [K]
    const uint32_t* counts;              // [K][4] of the final plane
    const double* plane;                 // [K]
    const unsigned long long* node_pfx;  // [K + 1] (splits before << 32 | child members before)
    const unsigned long long* leaf_pfx;  // [K + 1] leaf members before
    uint32_t axis;
    uint32_t level_base;       // index of the level's first node in the output
    uint32_t next_level_base;  // index of the next level's first node
    uint32_t leaf_items_base;  // leaf members emitted by earlier levels
};

// PtKdNode records of the level + the member ranges of the next level's nodes (front child, then back child)
__global__ void emit_nodes_kernel(EmitArgs a, uint32_t k_nodes, PtKdNode* __restrict__ out_nodes, uint32_t* __restrict__ seg_next) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_nodes) return;
    PtKdNode rec;
    if (a.status[k] == 0) {
        const uint32_t first = a.leaf_items_base + (uint32_t)a.leaf_pfx[k];
        rec.split = 0.0;
        rec.a = 3u | (first << 2);
        rec.b = a.seg[k + 1] - a.seg[k];
    } else {
        const uint32_t r = (uint32_t)(a.node_pfx[k] >> 32);
        const uint32_t child_base = (uint32_t)(a.node_pfx[k] & 0xFFFFFFFFull);
        const uint32_t n_front = a.counts[4 * k + 0] + a.counts[4 * k + 2];
        const uint32_t front = a.next_level_base + 2u * r;
        rec.split = a.plane[k];
        rec.a = a.axis | (front << 2);
        rec.b = front + 1u;
        seg_next[2u * r] = child_base;
        seg_next[2u * r + 1u] = child_base + n_front;
    }
    out_nodes[a.level_base + k] = rec;
    if (k == k_nodes - 1) seg_next[2u * (uint32_t)(a.node_pfx[k_nodes] >> 32)] = (uint32_t)(a.node_pfx[k_nodes] & 0xFFFFFFFFull);
}

// stable scatter of every member: leaf members to the output list, the others to the front / back child (shared: both)
__global__ void __launch_bounds__(kB) scatter_kernel(EmitArgs a, const uint32_t* __restrict__ items, const uint32_t* __restrict__ owner,
                                                     const uint8_t* __restrict__ cls, const unsigned long long* __restrict__ item_pfx,
                                                     uint32_t m, uint32_t* __restrict__ out_items, uint32_t* __restrict__ items_next,
                                                     uint32_t* __restrict__ owner_next) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t k = owner[i], it = items[i], c = cls[i];
    const uint32_t s = a.seg[k];
    if (c == 3u) {
        out_items[a.leaf_items_base + (uint32_t)a.leaf_pfx[k] + (i - s)] = it;
        return;
    }
    const uint32_t r = (uint32_t)(a.node_pfx[k] >> 32);
    const uint32_t child_base = (uint32_t)(a.node_pfx[k] & 0xFFFFFFFFull);
    const unsigned long long p = item_pfx[i], p0 = item_pfx[s];
    if (c == 0u || c == 2u) {
        const uint32_t dest = child_base + (uint32_t)((p & 0xFFFFFFFFull) - (p0 & 0xFFFFFFFFull));
        items_next[dest] = it;
        owner_next[dest] = 2u * r;
    }
    if (c == 1u || c == 2u) {
        const uint32_t n_front = a.counts[4 * k + 0] + a.counts[4 * k + 2];
        const uint32_t dest = child_base + n_front + (uint32_t)((p >> 32) - (p0 >> 32));
        items_next[dest] = it;
        owner_next[dest] = 2u * r + 1u;
    }
}

}  // namespace

struct KdTreeDev {
    KdAllocator al{};
    DevBuf nodes, items;
    uint32_t n_nodes = 0, n_items = 0, depth = 0;
    double root_bounds[6] = {0, 0, 0, 0, 0, 0};
    uint32_t launches = 0;
    float device_ms = 0.f;
    uint64_t bytes = 0;   // algorithmic bytes of the build (see the level loop)
    uint32_t tries = 0;   // plane-search rounds that did work, summed over levels
};

#define KD_TRY(expr)                                  \
    do {                                              \
        cudaError_t e_ = (expr);                      \
        if (e_ != cudaSuccess) { err = e_; goto done; } \
    } while (0)

cudaError_t kd_build_device(const double* d_bounds_aos, uint32_t n, const PtKdBuildConfig& cfg, const KdAllocator& al, cudaStream_t st,
                            KdTreeDev** out) {
    cudaError_t err = cudaSuccess;
    KdTreeDev* tree = new KdTreeDev();
    tree->al = al;
    tree->nodes.al = tree->items.al = &tree->al;
    DevBuf soa, items[2], owner[2], seg[2], cls, item_pfx, node_pfx, leaf_pfx, tiles, klo, khi, plane, pmin, pmax, counts, status, scalars;
    DevBuf* const all_scratch[] = {&soa, &items[0], &items[1], &owner[0], &owner[1], &seg[0], &seg[1], &cls, &item_pfx, &node_pfx,
                                   &leaf_pfx, &tiles, &klo, &khi, &plane, &pmin, &pmax, &counts, &status, &scalars};
    for (DevBuf* b : all_scratch) b->al = &tree->al;
    // first guesses that spare the level loop most re-allocations: a level rarely holds more than ~2n members (shared
    // members are duplicated) or more than n nodes; buffers still grow geometrically when a scene needs more
    const size_t m_guess = std::max<size_t>(2 * (size_t)n, 1024), k_guess = std::max<size_t>(n, 1024);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint32_t launches = 0;
    // pinned scalars read back once per level: [0] node scan total, [1] leaf scan total, [2..7] root bound keys
    unsigned long long* h_scalars = nullptr;
    uint32_t K = 1, M = n, level_base = 0, leaf_items_base = 0, level = 0;
    int cur = 0;

    KD_TRY(cudaEventCreate(&ev0));
    KD_TRY(cudaEventCreate(&ev1));
    KD_TRY(cudaMallocHost(&h_scalars, 8 * sizeof(unsigned long long)));
    KD_TRY(scalars.reserve(16 * 8, false, st));
    KD_TRY(cudaEventRecord(ev0, st));

    if (n == 0) {
        // Bounds for an empty Vec: BoundingBox::new(zero, zero) (bounding_box.rs:25-27); a single empty leaf
        PtKdNode rec{0.0, 3u, 0u};
        KD_TRY(tree->nodes.reserve(sizeof(PtKdNode), false, st));
        KD_TRY(cudaMemcpyAsync(tree->nodes.p, &rec, sizeof rec, cudaMemcpyHostToDevice, st));
        KD_TRY(cudaStreamSynchronize(st));
        tree->n_nodes = 1;
        goto done;
    }

    KD_TRY(soa.reserve((size_t)n * 6 * 8, false, st));
    transpose_bounds_kernel<<<blocks(n), kB, 0, st>>>(d_bounds_aos, n, soa.as<double>());
    for (int b = 0; b < 2; ++b) {
        KD_TRY(items[b].reserve(m_guess * 4, false, st));
        KD_TRY(owner[b].reserve(m_guess * 4, false, st));
        KD_TRY(seg[b].reserve((k_guess + 1) * 4, false, st));
    }
    KD_TRY(cls.reserve(m_guess, false, st));
    KD_TRY(item_pfx.reserve((m_guess + 1) * 8, false, st));
    for (DevBuf* b : {&klo, &khi, &plane, &pmin, &pmax, &node_pfx, &leaf_pfx}) KD_TRY(b->reserve((k_guess + 1) * 8, false, st));
    KD_TRY(counts.reserve(k_guess * 16, false, st));
    KD_TRY(status.reserve(k_guess, false, st));
    KD_TRY(tree->nodes.reserve(2 * k_guess * sizeof(PtKdNode), false, st));
    KD_TRY(tree->items.reserve(m_guess * 4, false, st));
    iota_kernel<<<blocks(n), kB, 0, st>>>(items[0].as<uint32_t>(), owner[0].as<uint32_t>(), n);
    {
        const uint32_t seg0[2] = {0u, n};
        KD_TRY(cudaMemcpyAsync(seg[0].p, seg0, sizeof seg0, cudaMemcpyHostToDevice, st));
    }
    launches += 2;

    // root bounds on all three axes (KDTreeNode::bounds of the root -> extent, node.rs:29,54-64)
    KD_TRY(klo.reserve(3 * 8, false, st));
    KD_TRY(khi.reserve(3 * 8, false, st));
    init_keys_kernel<<<1, kB, 0, st>>>(klo.as<unsigned long long>(), khi.as<unsigned long long>(), 3);
    for (int c = 0; c < 3; ++c) {
        bounds_kernel<<<blocks(n), kB, 0, st>>>(items[0].as<uint32_t>(), owner[0].as<uint32_t>(), n, soa.as<double>() + (size_t)c * n,
                                                soa.as<double>() + (size_t)(3 + c) * n, klo.as<unsigned long long>() + c,
                                                khi.as<unsigned long long>() + c);
    }
    launches += 4;
    KD_TRY(cudaMemcpyAsync(h_scalars + 2, klo.p, 3 * 8, cudaMemcpyDeviceToHost, st));
    KD_TRY(cudaMemcpyAsync(h_scalars + 5, khi.p, 3 * 8, cudaMemcpyDeviceToHost, st));
    KD_TRY(cudaStreamSynchronize(st));
    for (int c = 0; c < 3; ++c) {
        auto val = [](unsigned long long k) {
            const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
            double v;
            memcpy(&v, &b, 8);
            return v;
        };
        tree->root_bounds[c] = val(h_scalars[2 + c]);
        tree->root_bounds[3 + c] = val(h_scalars[5 + c]);
    }

    for (;; ++level) {
        const uint32_t axis = level % 3;  // (1,0,0) -> (0,1,0) -> (0,0,1), leaf.rs:97-103
        const uint32_t depth_left = cfg.max_depth > level ? cfg.max_depth - level : 0;
        const double* bmin = soa.as<double>() + (size_t)axis * n;
        const double* bmax = soa.as<double>() + (size_t)(3 + axis) * n;
        const uint32_t* d_items = items[cur].as<uint32_t>();
        const uint32_t* d_owner = owner[cur].as<uint32_t>();

        KD_TRY(klo.reserve((size_t)K * 8, false, st));
        KD_TRY(khi.reserve((size_t)K * 8, false, st));
        KD_TRY(plane.reserve((size_t)K * 8, false, st));
        KD_TRY(pmin.reserve((size_t)K * 8, false, st));
        KD_TRY(pmax.reserve((size_t)K * 8, false, st));
        KD_TRY(counts.reserve((size_t)K * 16, false, st));
        KD_TRY(status.reserve((size_t)K, false, st));
        KD_TRY(node_pfx.reserve((size_t)(K + 1) * 8, false, st));
        KD_TRY(leaf_pfx.reserve((size_t)(K + 1) * 8, false, st));
        KD_TRY(cls.reserve(std::max<size_t>(M, 1), false, st));
        KD_TRY(item_pfx.reserve((size_t)(M + 1) * 8, false, st));
        KD_TRY(tree->nodes.reserve((size_t)(level_base + K) * sizeof(PtKdNode), true, st));

        LevelState ls{seg[cur].as<uint32_t>(), klo.as<unsigned long long>(), khi.as<unsigned long long>(), plane.as<double>(),
                      pmin.as<double>(), pmax.as<double>(), counts.as<uint32_t>(), status.as<uint8_t>(),
                      reinterpret_cast<uint32_t*>(scalars.as<unsigned long long>() + 8),
                      reinterpret_cast<uint32_t*>(scalars.as<unsigned long long>() + 3)};
        KD_TRY(cudaMemsetAsync(ls.n_searching, 0, 4, st));
        KD_TRY(cudaMemsetAsync(ls.tries_used, 0, 8, st));
        if (depth_left > 0 && M > 0) {
            init_keys_kernel<<<blocks(K), kB, 0, st>>>(ls.klo, ls.khi, K);
            bounds_kernel<<<blocks(M), kB, 0, st>>>(d_items, d_owner, M, bmin, bmax, ls.klo, ls.khi);
            launches += 2;
        }
        node_init_kernel<<<blocks(K), kB, 0, st>>>(ls, K, depth_left, cfg.target_max_nodes);
        ++launches;
        if (depth_left > 0 && M > 0) {
            for (uint32_t attempt = 0; attempt < cfg.max_tries; ++attempt) {  // leaf.rs:156-201
                count_kernel<false><<<blocks(M), kB, 0, st>>>(d_items, d_owner, M, bmin, bmax, ls);
                try_update_kernel<<<blocks(K), kB, 0, st>>>(ls, K, cfg.target_max_merit);
                launches += 2;
            }
            // leaf.rs:204-215: the partition against the plane the search ended on
            count_kernel<true><<<blocks(M), kB, 0, st>>>(d_items, d_owner, M, bmin, bmax, ls);
            ++launches;
        }
        if (M > 0) {
            final_class_kernel<<<blocks(M), kB, 0, st>>>(d_items, d_owner, M, bmin, bmax, ls.plane, ls.status, cls.as<uint8_t>());
            ++launches;
        }
        unsigned long long* d_tot = scalars.as<unsigned long long>();
        KD_TRY(exclusive_scan(NodeLoad{ls.status, ls.counts}, K, tiles, d_tot + 0, node_pfx.as<unsigned long long>(), st));
        KD_TRY(exclusive_scan(LeafLoad{ls.status, ls.seg}, K, tiles, d_tot + 1, leaf_pfx.as<unsigned long long>(), st));
        KD_TRY(exclusive_scan(ClassLoad{cls.as<uint8_t>()}, M, tiles, d_tot + 2, item_pfx.as<unsigned long long>(), st));
        launches += 9;
        KD_TRY(cudaMemcpyAsync(h_scalars, d_tot, 4 * 8, cudaMemcpyDeviceToHost, st));
        KD_TRY(cudaStreamSynchronize(st));
        {
            // bytes this level has to move at the very least (the build's roofline numerator): per member 4 B id + 4 B
            // owner + 2 x 8 B bounds for the bounds pass, for every try that still had a searching node and for the final
            // classification (+ 1 B class out); class in + 8 B prefix out for the scan; id, owner, class, prefix in and
            // one or two (id, owner) pairs out for the scatter; per node 16 B record out + ~60 B of state
            const uint64_t tries = (uint32_t)h_scalars[3];
            const uint64_t m = M, mn = h_scalars[0] & 0xFFFFFFFFull;
            tree->bytes += (depth_left > 0 ? m * 24 * (1 + tries) : 0) + m * 25 + m * 9 + m * 17 + (mn + (uint32_t)h_scalars[1]) * 8 +
                           (uint64_t)K * 76;
            tree->tries += tries;
        }
        const uint32_t n_split = (uint32_t)(h_scalars[0] >> 32);
        const uint64_t m_next = h_scalars[0] & 0xFFFFFFFFull;
        const uint32_t leaf_members = (uint32_t)h_scalars[1];
        const uint32_t k_next = 2 * n_split;
        if ((uint64_t)level_base + K + k_next >= (1ull << 30) || (uint64_t)leaf_items_base + leaf_members >= (1ull << 30) ||
            m_next >= (1ull << 31)) {
            err = cudaErrorInvalidValue;  // does not fit the 30-bit fields of PtKdNode
            goto done;
        }
        KD_TRY(tree->items.reserve((size_t)(leaf_items_base + leaf_members) * 4 + 4, true, st));
        KD_TRY(items[cur ^ 1].reserve(std::max<size_t>(m_next, 1) * 4, false, st));
        KD_TRY(owner[cur ^ 1].reserve(std::max<size_t>(m_next, 1) * 4, false, st));
        KD_TRY(seg[cur ^ 1].reserve((size_t)(k_next + 1) * 4, false, st));

        EmitArgs ea{ls.seg, ls.status, ls.counts, ls.plane, node_pfx.as<unsigned long long>(), leaf_pfx.as<unsigned long long>(),
                    axis, level_base, level_base + K, leaf_items_base};
        emit_nodes_kernel<<<blocks(K), kB, 0, st>>>(ea, K, tree->nodes.as<PtKdNode>(), seg[cur ^ 1].as<uint32_t>());
        ++launches;
        if (M > 0) {
            scatter_kernel<<<blocks(M), kB, 0, st>>>(ea, d_items, d_owner, cls.as<uint8_t>(), item_pfx.as<unsigned long long>(), M,
                                                     tree->items.as<uint32_t>(), items[cur ^ 1].as<uint32_t>(),
                                                     owner[cur ^ 1].as<uint32_t>());
            ++launches;
        }
        KD_TRY(cudaGetLastError());
        level_base += K;
        leaf_items_base += leaf_members;
        if (k_next == 0) break;
        K = k_next;
        M = (uint32_t)m_next;
        cur ^= 1;
    }
    tree->n_nodes = level_base;
    tree->n_items = leaf_items_base;
    tree->depth = level;

done:
    if (err == cudaSuccess && ev0 && ev1) {
        cudaEventRecord(ev1, st);
        err = cudaStreamSynchronize(st);
        if (err == cudaSuccess) cudaEventElapsedTime(&tree->device_ms, ev0, ev1);
    }
    tree->launches = launches;
    for (DevBuf* b : all_scratch) b->release();
    if (h_scalars) cudaFreeHost(h_scalars);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (err != cudaSuccess) {
        tree->nodes.release();
        tree->items.release();
        delete tree;
        return err;
    }
    *out = tree;
    return cudaSuccess;
}

void kd_tree_release(KdTreeDev* t) {
    if (!t) return;
    t->nodes.release();
    t->items.release();
    delete t;
}
uint32_t kd_tree_node_count(const KdTreeDev* t) { return t->n_nodes; }
uint32_t kd_tree_item_count(const KdTreeDev* t) { return t->n_items; }
uint32_t kd_tree_depth(const KdTreeDev* t) { return t->depth; }
uint32_t kd_tree_launches(const KdTreeDev* t) { return t->launches; }
float kd_tree_device_ms(const KdTreeDev* t) { return t->device_ms; }
uint64_t kd_tree_algorithmic_bytes(const KdTreeDev* t) { return t->bytes; }
uint32_t kd_tree_tries(const KdTreeDev* t) { return t->tries; }
const double* kd_tree_root_bounds(const KdTreeDev* t) { return t->root_bounds; }
const PtKdNode* kd_tree_nodes_device(const KdTreeDev* t) { return t->nodes.as<PtKdNode>(); }
const uint32_t* kd_tree_items_device(const KdTreeDev* t) { return t->items.as<uint32_t>(); }

}  // namespace ptd
