// Scene flattening on the device — SURVEY §8(f) rank 2, the step that feeds the k-d tree build.
//
//   FlatScene::from            src/flat_scene.rs:18-46    breadth-first walk, total_trans = parent_trans * node.trans()
//   FlatSceneNode::new         src/flat_scene.rs:101-108  invtrans = trans.inverted(), normal_trans = invtrans^T
//   FlatSceneNode::bounds      src/flat_scene.rs:63-69    trans * primitive.bounds()
//   Mat4 * BoundingBox         src/bounding_box.rs:123-148 (the 8 corners, min / max)
//
// The reference pops a queue; the order in which it emits instances is the breadth-first order of the (instanced,
// i.e. expanded) hierarchy.  Here one LEVEL of that walk is one round: every entry of the level multiplies its
// parent's matrix by its node's, two exclusive scans give every entry its instance slot and the place of its
// children in the next level, and a second kernel writes the instance records — the 3x4 rows of trans and of its
// general 4x4 inverse (Laplace expansion, the same expression order as the host mirror, no fused multiply-add) and
// the world box of the primitive's 8 transformed corners — and the next level's entries.  Instances come out in the
// reference's order with bit-identical matrices and bounds (tests/test_flatten.py); the bounds feed pt_kd_build_device
// without leaving HBM.
#include "flatten.h"

#include <cstring>

#include "device_scan.cuh"

namespace ptd {
namespace {

struct Mat4d {
    double m[16];  // row-major
};

__device__ __forceinline__ Mat4d mat_mul(const Mat4d& a, const double* __restrict__ b) {  // a * b, Mat4::operator* order
    Mat4d r;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            r.m[i * 4 + j] = a.m[i * 4 + 0] * b[0 * 4 + j] + a.m[i * 4 + 1] * b[1 * 4 + j] + a.m[i * 4 + 2] * b[2 * 4 + j] +
                             a.m[i * 4 + 3] * b[3 * 4 + j];
    return r;
}

// general 4x4 inverse by Laplace expansion: expression for expression the host mirror's Mat4::inverted (host/math.hpp)
__device__ __forceinline__ Mat4d mat_inverse(const Mat4d& a) {
    const double m00 = a.m[0], m01 = a.m[1], m02 = a.m[2], m03 = a.m[3];
    const double m10 = a.m[4], m11 = a.m[5], m12 = a.m[6], m13 = a.m[7];
    const double m20 = a.m[8], m21 = a.m[9], m22 = a.m[10], m23 = a.m[11];
    const double m30 = a.m[12], m31 = a.m[13], m32 = a.m[14], m33 = a.m[15];
    const double s0 = m00 * m11 - m10 * m01, s1 = m00 * m12 - m10 * m02, s2 = m00 * m13 - m10 * m03;
    const double s3 = m01 * m12 - m11 * m02, s4 = m01 * m13 - m11 * m03, s5 = m02 * m13 - m12 * m03;
    const double c5 = m22 * m33 - m32 * m23, c4 = m21 * m33 - m31 * m23, c3 = m21 * m32 - m31 * m22;
    const double c2 = m20 * m33 - m30 * m23, c1 = m20 * m32 - m30 * m22, c0 = m20 * m31 - m30 * m21;
    const double invdet = 1.0 / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
    Mat4d r;
    r.m[0] = (m11 * c5 - m12 * c4 + m13 * c3) * invdet;
    r.m[1] = (-m01 * c5 + m02 * c4 - m03 * c3) * invdet;
    r.m[2] = (m31 * s5 - m32 * s4 + m33 * s3) * invdet;
    r.m[3] = (-m21 * s5 + m22 * s4 - m23 * s3) * invdet;
    r.m[4] = (-m10 * c5 + m12 * c2 - m13 * c1) * invdet;
    r.m[5] = (m00 * c5 - m02 * c2 + m03 * c1) * invdet;
    r.m[6] = (-m30 * s5 + m32 * s2 - m33 * s1) * invdet;
    r.m[7] = (m20 * s5 - m22 * s2 + m23 * s1) * invdet;
    r.m[8] = (m10 * c4 - m11 * c2 + m13 * c0) * invdet;
    r.m[9] = (-m00 * c4 + m01 * c2 - m03 * c0) * invdet;
    r.m[10] = (m30 * s4 - m31 * s2 + m33 * s0) * invdet;
    r.m[11] = (-m20 * s4 + m21 * s2 - m23 * s0) * invdet;
    r.m[12] = (-m10 * c3 + m11 * c1 - m12 * c0) * invdet;
    r.m[13] = (m00 * c3 - m01 * c1 + m02 * c0) * invdet;
    r.m[14] = (-m30 * s3 + m31 * s1 - m32 * s0) * invdet;
    r.m[15] = (m20 * s3 - m21 * s1 + m22 * s0) * invdet;
    return r;
}

// phase 1 of a level: total_trans of every entry (flat_scene.rs:30)
__global__ void __launch_bounds__(kB) level_transform_kernel(const PtHierNode* __restrict__ nodes, const uint32_t* __restrict__ entry_node,
                                                             const uint32_t* __restrict__ entry_parent, const Mat4d* __restrict__ parent_total,
                                                             uint32_t n_entries, Mat4d* __restrict__ total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    const PtHierNode& node = nodes[entry_node[i]];
    Mat4d parent;
    if (parent_total) {
        parent = parent_total[entry_parent[i]];
    } else {  // the root's parent is the identity (flat_scene.rs:24)
#pragma unroll
        for (int k = 0; k < 16; ++k) parent.m[k] = (k % 5 == 0) ? 1.0 : 0.0;
    }
    total[i] = mat_mul(parent, node.trans);
}

struct EntryLoad {  // entry -> (has geometry, number of children)
    const PtHierNode* nodes;
    const uint32_t* entry_node;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const {
        const PtHierNode& n = nodes[entry_node[i]];
        return ((unsigned long long)(n.geometry != 0xFFFFFFFFu ? 1u : 0u) << 32) | n.child_count;
    }
};

// the entries of the next level, one thread per CHILD (a root with 10^6 children is one entry): entry j belongs to
// the parent whose child range [pfx[i], pfx[i + 1]) holds j — remaining.push_back((total_trans, child)), flat_scene.rs:40-42
__global__ void __launch_bounds__(kB) level_children_kernel(const PtHierNode* __restrict__ nodes, const uint32_t* __restrict__ children,
                                                            const uint32_t* __restrict__ entry_node, const unsigned long long* __restrict__ pfx,
                                                            uint32_t n_entries, uint32_t n_next, uint32_t* __restrict__ next_node,
                                                            uint32_t* __restrict__ next_parent) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_next) return;
    uint32_t lo = 0, hi = n_entries;  // last i with child_base(i) <= j
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)(pfx[mid] & 0xFFFFFFFFull) <= j) lo = mid;
        else hi = mid;
    }
    const PtHierNode& node = nodes[entry_node[lo]];
    next_node[j] = children[node.first_child + (j - (uint32_t)(pfx[lo] & 0xFFFFFFFFull))];
    next_parent[j] = lo;
}

// phase 2 of a level: instance records of the entries that carry geometry
__global__ void __launch_bounds__(kB) level_emit_kernel(const PtHierNode* __restrict__ nodes,
                                                        const PtGeometryRec* __restrict__ geoms, const uint32_t* __restrict__ entry_node,
                                                        const Mat4d* __restrict__ total, const unsigned long long* __restrict__ pfx,
                                                        uint32_t n_entries, uint32_t instance_base, PtInstance* __restrict__ instances,
                                                        PtInstanceTrans* __restrict__ inst_trans, double* __restrict__ bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    const PtHierNode& node = nodes[entry_node[i]];
    const unsigned long long p = pfx[i];
    if (node.geometry == 0xFFFFFFFFu) return;
    const uint32_t slot = instance_base + (uint32_t)(p >> 32);
    const PtGeometryRec& g = geoms[node.geometry];
    const Mat4d t = total[i];
    const Mat4d inv = mat_inverse(t);  // FlatSceneNode::new, flat_scene.rs:103-104
    PtInstance rec;
    PtInstanceTrans tr;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        rec.invtrans[k] = inv.m[k];
        tr.trans[k] = t.m[k];
    }
    rec.prim = g.prim;
    rec.mesh = g.mesh;
    rec.material = g.material;
    rec.reserved = 0;
    rec.pad[0] = rec.pad[1] = 0.0;
    instances[slot] = rec;
    inst_trans[slot] = tr;
    // Mat4 * BoundingBox, bounding_box.rs:123-148: all 8 corners, x outermost, through transformed_point
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
        const double x = g.bounds[(corner & 4) ? 3 : 0], y = g.bounds[(corner & 2) ? 4 : 1], z = g.bounds[(corner & 1) ? 5 : 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double w = t.m[r * 4 + 0] * x + t.m[r * 4 + 1] * y + t.m[r * 4 + 2] * z + t.m[r * 4 + 3] * 1.0;
            lo[r] = w < lo[r] ? w : lo[r];   // Vec3::partial_min / partial_max
            hi[r] = w > hi[r] ? w : hi[r];
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        bounds[(size_t)slot * 6 + r] = lo[r];
        bounds[(size_t)slot * 6 + 3 + r] = hi[r];
    }
}

}  // namespace

struct FlatSceneDev {
    KdAllocator al{};
    DevBuf instances, inst_trans, bounds;
    uint32_t n_instances = 0, depth = 0, launches = 0;
    float device_ms = 0.f;
};

#define FL_TRY(expr)                                    \
    do {                                                \
        cudaError_t e_ = (expr);                        \
        if (e_ != cudaSuccess) { err = e_; goto done; } \
    } while (0)

cudaError_t flatten_device(const PtHierNode* d_nodes, uint32_t n_nodes, const uint32_t* d_children, uint32_t n_children, uint32_t root,
                           const PtGeometryRec* d_geoms, uint32_t max_levels, const KdAllocator& al, cudaStream_t st,
                           FlatSceneDev** out) {
    (void)n_children;
    cudaError_t err = cudaSuccess;
    FlatSceneDev* flat = new FlatSceneDev();
    flat->al = al;
    DevBuf entry_node[2], entry_parent[2], total[2], pfx, tiles, scalars;
    DevBuf* const scratch[] = {&entry_node[0], &entry_node[1], &entry_parent[0], &entry_parent[1], &total[0], &total[1], &pfx, &tiles, &scalars};
    for (DevBuf* b : scratch) b->al = &flat->al;
    flat->instances.al = flat->inst_trans.al = flat->bounds.al = &flat->al;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long* h_total = nullptr;
    uint32_t n_entries = 1, level = 0, launches = 0, n_inst = 0;
    int cur = 0;
    const size_t guess = std::max<size_t>(n_nodes, 1024);

    FL_TRY(cudaEventCreate(&ev0));
    FL_TRY(cudaEventCreate(&ev1));
    FL_TRY(cudaMallocHost(&h_total, sizeof(unsigned long long)));
    FL_TRY(scalars.reserve(64, false, st));
    for (int b = 0; b < 2; ++b) {
        FL_TRY(entry_node[b].reserve(guess * 4, false, st));
        FL_TRY(entry_parent[b].reserve(guess * 4, false, st));
        FL_TRY(total[b].reserve(guess * sizeof(Mat4d), false, st));
    }
    FL_TRY(pfx.reserve((guess + 1) * 8, false, st));
    FL_TRY(flat->instances.reserve(guess * sizeof(PtInstance), false, st));
    FL_TRY(flat->inst_trans.reserve(guess * sizeof(PtInstanceTrans), false, st));
    FL_TRY(flat->bounds.reserve(guess * 6 * sizeof(double), false, st));
    FL_TRY(cudaEventRecord(ev0, st));
    FL_TRY(cudaMemcpyAsync(entry_node[0].p, &root, 4, cudaMemcpyHostToDevice, st));

    for (;; ++level) {
        if (level > max_levels) { err = cudaErrorInvalidValue; goto done; }  // a cycle in the hierarchy
        FL_TRY(total[cur].reserve((size_t)n_entries * sizeof(Mat4d), false, st));
        FL_TRY(pfx.reserve((size_t)(n_entries + 1) * 8, false, st));
        level_transform_kernel<<<blocks(n_entries), kB, 0, st>>>(d_nodes, entry_node[cur].as<uint32_t>(), entry_parent[cur].as<uint32_t>(),
                                                                 level ? total[cur ^ 1].as<Mat4d>() : nullptr, n_entries,
                                                                 total[cur].as<Mat4d>());
        FL_TRY(exclusive_scan(EntryLoad{d_nodes, entry_node[cur].as<uint32_t>()}, n_entries, tiles, scalars.as<unsigned long long>(),
                              pfx.as<unsigned long long>(), st));
        launches += 4;
        FL_TRY(cudaMemcpyAsync(h_total, scalars.p, 8, cudaMemcpyDeviceToHost, st));
        FL_TRY(cudaStreamSynchronize(st));
        const uint32_t level_instances = (uint32_t)(*h_total >> 32);
        const uint64_t n_next = *h_total & 0xFFFFFFFFull;
        if ((uint64_t)n_inst + level_instances >= (1ull << 31) || n_next >= (1ull << 31)) { err = cudaErrorInvalidValue; goto done; }
        FL_TRY(flat->instances.reserve((size_t)(n_inst + level_instances) * sizeof(PtInstance), true, st));
        FL_TRY(flat->inst_trans.reserve((size_t)(n_inst + level_instances) * sizeof(PtInstanceTrans), true, st));
        FL_TRY(flat->bounds.reserve((size_t)(n_inst + level_instances) * 6 * sizeof(double), true, st));
        FL_TRY(entry_node[cur ^ 1].reserve(std::max<size_t>(n_next, 1) * 4, false, st));
        FL_TRY(entry_parent[cur ^ 1].reserve(std::max<size_t>(n_next, 1) * 4, false, st));
        level_emit_kernel<<<blocks(n_entries), kB, 0, st>>>(d_nodes, d_geoms, entry_node[cur].as<uint32_t>(), total[cur].as<Mat4d>(),
                                                            pfx.as<unsigned long long>(), n_entries, n_inst,
                                                            flat->instances.as<PtInstance>(), flat->inst_trans.as<PtInstanceTrans>(),
                                                            flat->bounds.as<double>());
        ++launches;
        if (n_next) {
            level_children_kernel<<<blocks(n_next), kB, 0, st>>>(d_nodes, d_children, entry_node[cur].as<uint32_t>(),
                                                                 pfx.as<unsigned long long>(), n_entries, (uint32_t)n_next,
                                                                 entry_node[cur ^ 1].as<uint32_t>(), entry_parent[cur ^ 1].as<uint32_t>());
            ++launches;
        }
        FL_TRY(cudaGetLastError());
        n_inst += level_instances;
        if (n_next == 0) break;
        n_entries = (uint32_t)n_next;
        cur ^= 1;
    }
    flat->n_instances = n_inst;
    flat->depth = level;

done:
    if (err == cudaSuccess && ev0 && ev1) {
        cudaEventRecord(ev1, st);
        err = cudaStreamSynchronize(st);
        if (err == cudaSuccess) cudaEventElapsedTime(&flat->device_ms, ev0, ev1);
    }
    flat->launches = launches;
    for (DevBuf* b : scratch) b->release();
    if (h_total) cudaFreeHost(h_total);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (err != cudaSuccess) {
        flat_release(flat);
        return err;
    }
    *out = flat;
    return cudaSuccess;
}

void flat_release(FlatSceneDev* f) {
    if (!f) return;
    f->instances.release();
    f->inst_trans.release();
    f->bounds.release();
    delete f;
}
uint32_t flat_instance_count(const FlatSceneDev* f) { return f->n_instances; }
uint32_t flat_depth(const FlatSceneDev* f) { return f->depth; }
uint32_t flat_launches(const FlatSceneDev* f) { return f->launches; }
float flat_device_ms(const FlatSceneDev* f) { return f->device_ms; }
const PtInstance* flat_instances_device(const FlatSceneDev* f) { return f->instances.as<PtInstance>(); }
const PtInstanceTrans* flat_trans_device(const FlatSceneDev* f) { return f->inst_trans.as<PtInstanceTrans>(); }
const double* flat_bounds_device(const FlatSceneDev* f) { return f->bounds.as<double>(); }

}  // namespace ptd
