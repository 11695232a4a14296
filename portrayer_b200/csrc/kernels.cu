// The wavefront pipeline that replaces the rayon pixel loop of
// src/render.rs:127-150.  One batch of paths is a forest of ray trees stored
// level by level in a node pool (device_scene.cuh):
//
//   camera_kernel        render.rs:36-40 + camera.rs:48-84     primary rays -> level 0
//   per level d = 0..10:
//     extend_kernel      ray.rs:140-141                        closest hit of every ray of the level
//     shadow_kernel      material.rs:150-179                   one any-hit walk per (hit, light)
//     shade_kernel       material.rs:91-320                    local colour; children appended to level d+1
//   tree_eval_kernel     material.rs:280,307-309,315           bottom-up colour of each path, reference order
//   resolve_kernel       render.rs:43-50,143-147               sum over samples, gamma, clamp, truncate to u8
//
// No atomics touch colours: every node's colour is computed by one thread in
// the reference's own expression order, so the result is deterministic and
// independent of batch size, tile ownership and GPU count.
#include <cuda_runtime.h>

#include "kernels.h"
#include "shade.cuh"
#include "traverse.cuh"

namespace ptd {

// Per-frame constants (scene pointers, camera, node pool, outputs).  Kernels take only the slot number, so a
// frame's CUDA graph can be replayed for any scene / camera after one cudaMemcpyToSymbolAsync.
__constant__ FrameState c_state[kStateSlots];

namespace {

#ifndef PT_BLOCK
#define PT_BLOCK 128
#endif
constexpr int kBlock = PT_BLOCK;
// Threads per block of the shade kernel = parents whose children are allocated together (block_alloc).  The bigger the
// group, the longer the runs of same-kind children from neighbouring parents in the next level, and the better that
// level's warps stay together: graphics-castle 4K x 16 renders in 324 / 304 / 291 ms with groups of 128 / 256 / 512
// (the traversal kernels' own block size does not matter: 325 ms at 256; profiles/r02_ab_shade_group.txt).  A 512-thread
// block leaves a small level with too few blocks to fill the machine (configs[1], 0.5 M paths per frame: -4 %), so a
// frame picks by the size of its batches.
#ifndef PT_SHADE_BLOCK_BIG
#define PT_SHADE_BLOCK_BIG 512
#endif
// rounds of the big block per tile: 4 x 512 parents = 165 KB of shared memory, one block per SM
#ifndef PT_SHADE_TILE_ROUNDS
#define PT_SHADE_TILE_ROUNDS 4
#endif
constexpr int kShadeBlockSmall = kBlock, kShadeBlockBig = PT_SHADE_BLOCK_BIG, kShadeRoundsBig = PT_SHADE_TILE_ROUNDS;
constexpr uint64_t kShadeBigPaths = 1ull << 21;  // batches of at least this many paths use the big group
// minimum resident blocks per SM asked of ptxas (register budget = 65536 / (kBlock * min blocks)); tuned on B200,
// see DESIGN.md "occupancy"
#ifndef PT_EXTEND_MIN_BLOCKS
#define PT_EXTEND_MIN_BLOCKS 4
#endif
#ifndef PT_SHADOW_MIN_BLOCKS
#define PT_SHADOW_MIN_BLOCKS 4
#endif
#ifndef PT_SHADE_MIN_BLOCKS
#define PT_SHADE_MIN_BLOCKS 5
#endif

PT_D void background_of(const FrameParams& fp, uint32_t pixel, double* bg) {
    const double* src = fp.bg_mode == PT_BG_PER_PIXEL ? fp.background + (size_t)pixel * 3
                        : fp.bg_mode == PT_BG_PER_ROW ? fp.background + (size_t)(pixel / fp.width) * 3
                                                      : fp.background;
    bg[0] = __ldg(src);
    bg[1] = __ldg(src + 1);
    bg[2] = __ldg(src + 2);
}

// warp-aggregated add of per-thread work counters into the batch control block
PT_D void flush_counters(BatchCtl* ctl, int kind, const WorkCounters& wc) {
    unsigned long long v[10] = {wc.kd_splits, wc.instance_tests, wc.triangle_tests, wc.bbox_gates, wc.prim_flops,
                                wc.x_box,     wc.x_inst,         wc.x_tri,          wc.x_gate,     wc.x_prim_flops};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xFFFFFFFFu, v[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k)
            if (v[k]) atomicAdd(&ctl->work[kind][k], v[k]);
    }
}

// first device-detected error of the batch: where the reference would have panicked (cold path)
PT_D void record_error(const FrameParams& fp, const NodePool& pool, BatchCtl* ctl, uint32_t bits, uint32_t node, uint32_t where) {
    if (atomicCAS(&ctl->err_info[0], 0u, bits) != 0u) return;
    const uint32_t root = pool.root[node];
    ctl->err_info[1] = fp.pixel_index ? fp.pixel_index[ctl->first_slot + root / fp.samples] : ctl->first_slot + root;
    ctl->err_info[2] = root % fp.samples;
    ctl->err_info[3] = pool.pathid[node];
    ctl->err_info[4] = where;
}

// the rays of the level the control block points at: n of them, ray j at pool index at(j)
struct LevelRays {
    uint32_t begin, n_lo, hi_base, n;
    PT_D uint32_t at(uint32_t j) const { return j < n_lo ? begin + j : hi_base + (j - n_lo); }
};
PT_D LevelRays level_rays(const BatchCtl* ctl, const NodePool& pool) {
    const uint32_t level = ctl->level;
    LevelRays r;
    r.begin = ctl->level_start[level];
    r.n_lo = ctl->level_start[level + 1] - r.begin;
    const uint32_t hi_now = ctl->level_hi[level], hi_prev = level ? ctl->level_hi[level - 1] : 0u;
    r.hi_base = pool.capacity - hi_now;
    r.n = r.n_lo + (hi_now - hi_prev);
    return r;
}

// ------------------------------------------------------------------ camera
// path p of the batch = (owned pixel slot, sample); Camera::ray_at, camera.rs:48-84
// Block 0 also resets the batch control block: nothing in this kernel reads it, and every later kernel of the
// batch starts after this one has finished.
PT_D void begin_batch(BatchCtl* ctl, uint32_t first_slot, uint32_t n_slots, uint32_t n_paths) {
    if (blockIdx.x != 0) return;
    uint32_t* words = reinterpret_cast<uint32_t*>(ctl);
    for (uint32_t i = threadIdx.x; i < sizeof(BatchCtl) / 4; i += blockDim.x) words[i] = 0u;
    __syncthreads();
    if (threadIdx.x < 16) ctl->level_start[threadIdx.x] = threadIdx.x == 0 ? 0u : n_paths;
    if (threadIdx.x == 0) {
        ctl->pool_count = n_paths;
        ctl->first_slot = first_slot;
        ctl->n_slots = n_slots;
        ctl->n_paths = n_paths;
    }
}

__global__ void __launch_bounds__(kBlock) camera_kernel(int slot_id, uint32_t first_slot, uint32_t n_slots) {
    const FrameState& fs = c_state[slot_id];
    const FrameParams& fp = fs.fp;
    const NodePool& pool = fs.pool;
    const uint32_t n_paths = n_slots * fp.samples;
    begin_batch(fs.ctl, first_slot, n_slots, n_paths);
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const uint32_t slot = first_slot + p / fp.samples;
    const uint32_t sample = p % fp.samples;
    const uint32_t pixel = __ldg(fp.pixel_index + slot);
    const uint32_t px = pixel % fp.width, py = pixel / fp.width;
    // Choose a random point in the pixel square, render.rs:38-39
    const double x = (double)px + draw(fp.rng_mode, fp.seed, pixel, sample, 1, 0);
    const double y = (double)py + draw(fp.rng_mode, fp.seed, pixel, sample, 1, 1);

    const PtCamera& cam = fp.cam;
    const double pixel_ndc_y = y / cam.height;
    const double pixel_view_y = (1.0 - 2.0 * pixel_ndc_y) * cam.fov_factor;
    const double pixel_ndc_x = x / cam.width;
    const double pixel_view_x = (2.0 * pixel_ndc_x - 1.0) * cam.aspect_ratio * cam.fov_factor;
    const V3 pixel_view = v3(pixel_view_x, pixel_view_y, -1.0);
    const V3 pixel_world = xf_point(cam.view_to_world, pixel_view);
    const V3 eye = v3(cam.eye[0], cam.eye[1], cam.eye[2]);
    const V3 dir = normalized(pixel_world - eye);

    pool.ox[p] = eye.x; pool.oy[p] = eye.y; pool.oz[p] = eye.z;
    pool.dx[p] = dir.x; pool.dy[p] = dir.y; pool.dz[p] = dir.z;
    pool.root[p] = p;
    pool.pathid[p] = 1u;
}

// explicit rays instead of camera rays (pt_trace_rays): ray i is "pixel" i, sample 0
__global__ void __launch_bounds__(kBlock) load_rays_kernel(int slot_id, uint32_t first, uint32_t n_paths) {
    const FrameState& fs = c_state[slot_id];
    const NodePool& pool = fs.pool;
    const double* __restrict__ origins = fs.ray_origins;
    const double* __restrict__ dirs = fs.ray_dirs;
    begin_batch(fs.ctl, first, n_paths, n_paths);
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const size_t i = (size_t)(first + p) * 3;
    pool.ox[p] = origins[i]; pool.oy[p] = origins[i + 1]; pool.oz[p] = origins[i + 2];
    pool.dx[p] = dirs[i]; pool.dy[p] = dirs[i + 1]; pool.dz[p] = dirs[i + 2];
    pool.root[p] = p;
    pool.pathid[p] = 1u;
}

// ------------------------------------------------------------------ extend (closest hit)
// LINEAR: the PT_RENDER_LINEAR_TLAS cross-check (scene_cast_linear: no k-d tree, every instance in list order)
// PRUNE: the walk skips subtrees whose node box the ray misses (traverse.cuh kd_walk); off for PT_RENDER_EXACT_WALK and
// for the counting kernels, which report the reference's work
template <bool COUNT, bool LINEAR = false, bool PRUNE = false>
__global__ void __launch_bounds__(kBlock, PT_EXTEND_MIN_BLOCKS) extend_kernel(int slot_id) {
    const FrameState& fs = c_state[slot_id];
    const DScene& sc = fs.sc;
    const NodePool& pool = fs.pool;
    BatchCtl* ctl = fs.ctl;
    const uint32_t level = ctl->level;
    const LevelRays rays = level_rays(ctl, pool);
    KdStack tlas_stack, blas_stack;
    WorkCounters wc;
    uint32_t err = 0;
    // Dynamic distribution: a warp claims the next 32 rays when it has finished its last 32 (one atomic per warp
    // and chunk).  Ray cost varies by orders of magnitude (sky vs. deep kd walks); a static grid-stride split leaves
    // most of the machine idle while the slowest warps finish.
    const uint32_t n = rays.n;
    const int lane = threadIdx.x & 31;
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(&ctl->cursor[0], 32u);
    for (;;) {
        const uint32_t base = __shfl_sync(0xFFFFFFFFu, next, 0);
        if (base >= n) break;
        // claim the following chunk now: the atomic's round trip overlaps the walk of this one
        if (lane == 0) next = atomicAdd(&ctl->cursor[0], 32u);
        if (base + lane >= n) continue;
        const uint32_t i = rays.at(base + lane);
        const V3 o = v3(pool.ox[i], pool.oy[i], pool.oz[i]);
        const V3 d = v3(pool.dx[i], pool.dy[i], pool.dz[i]);
        Hit hit{(double)INFINITY, kNone, 0};
        const uint32_t err_before = err;
        const bool found = LINEAR ? scene_cast_linear<false, COUNT>(sc, o, d, hit, blas_stack, err, wc)
                                  : scene_cast<false, COUNT, PRUNE>(sc, o, d, hit, tlas_stack, blas_stack, err, wc);
        if (err != err_before) record_error(fs.fp, pool, ctl, err & ~err_before, i, 0u | level << 8);
        pool.t[i] = found ? hit.t : (double)INFINITY;
        pool.inst[i] = found ? hit.inst : kNone;
        pool.sub[i] = found ? hit.sub : 0u;
    }
    if (err) atomicOr(&ctl->error_bits, err);
    if (COUNT) flush_counters(ctl, 0, wc);
}

// ------------------------------------------------------------------ shadow (any hit), light-major
template <bool COUNT, bool LINEAR = false, bool PRUNE = false>
__global__ void __launch_bounds__(kBlock, PT_SHADOW_MIN_BLOCKS) shadow_kernel(int slot_id) {
    const FrameState& fs = c_state[slot_id];
    const DScene& sc = fs.sc;
    const FrameParams& fp = fs.fp;
    const NodePool& pool = fs.pool;
    BatchCtl* ctl = fs.ctl;
    const uint32_t level = ctl->level;
    const uint32_t first_slot = ctl->first_slot;
    const LevelRays rays = level_rays(ctl, pool);
    const uint32_t n = rays.n;
    const unsigned long long total = (unsigned long long)n * sc.n_lights;
    KdStack tlas_stack, blas_stack;
    WorkCounters wc;
    uint32_t err = 0;
    unsigned long long cast = 0;
    const int lane = threadIdx.x & 31;
    uint32_t next = 0;
    if (lane == 0) next = atomicAdd(&ctl->cursor[1], 32u);
    for (;;) {  // dynamic distribution, as in extend_kernel (total < 2^32 is guaranteed by the host when it sizes the pool)
        const uint32_t base = __shfl_sync(0xFFFFFFFFu, next, 0);
        if ((unsigned long long)base >= total) break;
        if (lane == 0) next = atomicAdd(&ctl->cursor[1], 32u);
        const unsigned long long j = (unsigned long long)base + lane;
        if (j >= total) continue;
        const uint32_t l = (uint32_t)(j / n);
        const uint32_t i = rays.at((uint32_t)(j % n));
        const uint32_t inst = pool.inst[i];
        if (inst == kNone) continue;
        const V3 o = v3(pool.ox[i], pool.oy[i], pool.oz[i]);
        const V3 d = v3(pool.dx[i], pool.dy[i], pool.dz[i]);
        const V3 hit_point = world_hit_point(sc, inst, o, d, pool.t[i], nullptr, nullptr, nullptr);
        const uint32_t root = pool.root[i];
        const uint32_t pixel = fp.pixel_index ? __ldg(fp.pixel_index + first_slot + root / fp.samples) : first_slot + root;
        const uint32_t sample = root % fp.samples;
        const V3 light_pos = light_sample_position(sc.lights + l, l, fp.rng_mode, fp.seed, pixel, sample, pool.pathid[i]);
        // material.rs:156-179
        const V3 hit_to_light = light_pos - hit_point;
        const double light_dist = magnitude(hit_to_light);
        const V3 light_dir = hit_to_light / light_dist;
        Hit hit{(double)INFINITY, kNone, 0};
        const uint32_t err_before = err;
        const bool occluded = LINEAR ? scene_cast_linear<true, COUNT>(sc, hit_point, light_dir, hit, blas_stack, err, wc)
                                     : scene_cast<true, COUNT, PRUNE>(sc, hit_point, light_dir, hit, tlas_stack, blas_stack, err, wc);
        if (err != err_before) record_error(fp, pool, ctl, err & ~err_before, i, 1u | level << 8 | l << 16);
        pool.occl[(size_t)l * pool.capacity + i] = occluded ? 1 : 0;
        ++cast;
    }
    if (err) atomicOr(&ctl->error_bits, err);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cast += __shfl_down_sync(0xFFFFFFFFu, cast, off);
    if ((threadIdx.x & 31) == 0 && cast) atomicAdd(&ctl->rays_shadow, cast);
    if (COUNT) flush_counters(ctl, 1, wc);
}

// ------------------------------------------------------------------ shade
// `loop`: the conditional handle of the frame graph's WHILE node (0 on the stream path); the last block to
// finish decides whether another recursion level has rays to trace.
// A block shades a TILE of K * BLOCK consecutive rays of the level, K rounds of BLOCK; the children the rounds ask for are
// kept in shared memory (ShadeTile) and allocated ONCE per tile, in parent order: reflected children to the bottom end
// of the pool, refracted children to its top end (device_scene.cuh BatchCtl), so a level is all of its reflected rays
// followed by all of its refracted rays.  Allocating per warp (32 parents -> up to 32 + 32 slots) made every 32-ray chunk
// of the next level a mixture of reflected and refracted rays, and the mixture compounds level by level: on
// graphics-castle ncu counted 12 of 32 lanes active in the extend launches of levels >= 1.  And where few parents spawn
// children at all — the deep levels of a scene with a few dielectrics — a small group yields a run of a few dozen
// children, so every chunk of the next level straddles runs from unrelated corners of the picture: the bigger the tile,
// the longer the runs (graphics-castle 4K x 16: 324 / 304 / 291 ms with groups of 128 / 256 / 512 parents,
// profiles/r02_ab_shade_group.txt).  (Sorting every level by (direction octant, origin cell) was measured too and lost:
// DESIGN.md section 10.)
template <int BLOCK, int K>
struct ShadeTile {
    static constexpr int T = BLOCK * K;  // parents per tile
    static constexpr int W = T / 32;     // warp-rounds per tile
    static constexpr size_t kBytes = (size_t)T * (9 * sizeof(double) + 2 * sizeof(uint32_t)) + (size_t)W * 4 * sizeof(uint32_t);
    double* d;       // [9][T]: hit point, reflected direction, refracted direction
    uint32_t* u;     // [2][T]: root, path id
    uint32_t* mask;  // [2][W]: which lanes of each warp-round want a reflected / refracted child
    uint32_t* pre;   // [2][W]: children wanted by the warp-rounds before this one
    PT_D explicit ShadeTile(unsigned char* smem) {
        d = reinterpret_cast<double*>(smem);
        u = reinterpret_cast<uint32_t*>(d + 9 * T);
        mask = u + 2 * T;
        pre = mask + 2 * W;
    }
};

template <int BLOCK, int K>
__global__ void __launch_bounds__(BLOCK, (BLOCK >= 512 ? 1 : PT_SHADE_MIN_BLOCKS)) shade_kernel(int slot_id, cudaGraphConditionalHandle loop) {
    extern __shared__ __align__(16) unsigned char shade_smem[];
    using Tile = ShadeTile<BLOCK, K>;
    const Tile tile_mem(shade_smem);
    __shared__ uint32_t s_base[2], s_limit[2];
    const FrameState& fs = c_state[slot_id];
    const DScene& sc = fs.sc;
    const FrameParams& fp = fs.fp;
    const NodePool& pool = fs.pool;
    BatchCtl* ctl = fs.ctl;
    const uint32_t level = ctl->level;
    const uint32_t first_slot = ctl->first_slot;
    const LevelRays rays = level_rays(ctl, pool);
    const uint32_t n = rays.n;
    const uint32_t n_tiles = (n + Tile::T - 1) / Tile::T;  // (tiles are per block: every thread of a block runs the same loops)
    const int lane = threadIdx.x & 31;
    uint32_t err = 0, err_seen = 0;
    unsigned long long n_shaded = 0, n_reflect = 0, n_refract = 0, n_cut = 0, n_texel = 0;

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int k = 0; k < K; ++k) {
        const int slot = k * BLOCK + (int)threadIdx.x;
        const uint32_t j = tile * (uint32_t)Tile::T + (uint32_t)slot;
        const bool active = j < n;
        const uint32_t i = active ? rays.at(j) : 0u;
        bool want0 = false, want1 = false;  // children to trace
        V3 hit_point = v3(0, 0, 0), dir0 = v3(0, 0, 0), dir1 = v3(0, 0, 0);
        uint32_t root = 0, pathid = 0;

        if (active) {
            root = pool.root[i];
            pathid = pool.pathid[i];
            const uint32_t pixel = fp.pixel_index ? __ldg(fp.pixel_index + first_slot + root / fp.samples) : first_slot + root;
            const uint32_t sample = root % fp.samples;
            const uint32_t inst = pool.inst[i];
            double color[3];
            uint8_t mode = kModeLeaf;
            uint32_t c0 = kChildNone, c1 = kChildNone;
            double refl = 0.0, fres = 0.0;

            if (inst == kNone) {
                background_of(fp, pixel, color);  // ray.rs:146
            } else {
                ++n_shaded;
                const V3 o = v3(pool.ox[i], pool.oy[i], pool.oz[i]);
                const V3 ray_dir = v3(pool.dx[i], pool.dy[i], pool.dz[i]);
                const PtMaterial* mat = sc.materials + __ldg(&sc.instances[inst].material);
                const int tex_id = __ldg(&mat->texture), nrm_id = __ldg(&mat->normals);
                SurfaceHit sh;
                reconstruct_hit(sc, inst, pool.sub[i], o, ray_dir, pool.t[i], tex_id >= 0 || nrm_id >= 0, sh);
                hit_point = sh.hit_point;
                const V3 view = -ray_dir;

                // uv' = uv_trans * (u, v, 1), material.rs:114-117
                double tu = 0.0, tv = 0.0;
                if (sh.has_uv) {
                    const double* m = mat->uv_trans;
                    tu = __ldg(m + 0) * sh.u + __ldg(m + 1) * sh.v + __ldg(m + 2) * 1.0;
                    tv = __ldg(m + 3) * sh.u + __ldg(m + 4) * sh.v + __ldg(m + 5) * 1.0;
                }
                V3 normal;
                bool ok = true;
                if (nrm_id < 0) {
                    normal = normalized(sh.normal);
                } else if (sh.has_uv && sh.has_nmt) {
                    double c[3];
                    texture_at(sc, nrm_id, tu, tv, c);
                    ++n_texel;
                    // NormalMap::normal_at: (nx, ny, nz) -> (nx, -nz, -ny), texture.rs:203-220
                    const V3 norm = v3(2.0 * c[0] - 1.0, 2.0 * c[1] - 1.0, -(2.0 * c[2] - 1.0));
                    const V3 tn = normalized(v3(norm.x, -norm.z, -norm.y));
                    const double* m = sh.nmt;
                    normal = v3(m[0] * tn.x + m[1] * tn.y + m[2] * tn.z, m[3] * tn.x + m[4] * tn.y + m[5] * tn.z,
                                m[6] * tn.x + m[7] * tn.y + m[8] * tn.z);
                } else {
                    err |= PT_DEVERR_NORMALMAP;  // material.rs:133
                    normal = v3(0, 0, 0);
                    ok = false;
                }
                double kd[3] = {0, 0, 0};
                if (tex_id < 0) {
                    kd[0] = __ldg(&mat->diffuse[0]); kd[1] = __ldg(&mat->diffuse[1]); kd[2] = __ldg(&mat->diffuse[2]);
                } else if (sh.has_uv) {
                    texture_at_gamma(sc, tex_id, tu, tv, kd);  // texel / 255 then powf(2.2), texture.rs:166
                    ++n_texel;
                } else {
                    err |= PT_DEVERR_TEXTURE;  // material.rs:141
                    ok = false;
                }

                color[0] = sc.ambient[0] * kd[0]; color[1] = sc.ambient[1] * kd[1]; color[2] = sc.ambient[2] * kd[2];
                if (ok) {
                    const double spec[3] = {__ldg(&mat->specular[0]), __ldg(&mat->specular[1]), __ldg(&mat->specular[2])};
                    const bool has_spec = spec[0] > kEps || spec[1] > kEps || spec[2] > kEps;
                    const double shininess = __ldg(&mat->shininess);
                    for (uint32_t l = 0; l < sc.n_lights; ++l) {  // material.rs:149-212
                        if (pool.occl[(size_t)l * pool.capacity + i]) continue;
                        const PtLight* light = sc.lights + l;
                        const V3 light_pos = light_sample_position(light, l, fp.rng_mode, fp.seed, pixel, sample, pathid);
                        const V3 hit_to_light = light_pos - hit_point;
                        const double light_dist = magnitude(hit_to_light);
                        const V3 light_dir = hit_to_light / light_dist;
                        const double f0 = __ldg(&light->falloff[0]), f1 = __ldg(&light->falloff[1]), f2 = __ldg(&light->falloff[2]);
                        const double attenuation = f0 + f1 * light_dist + f2 * light_dist * light_dist;
                        const double lc[3] = {__ldg(&light->color[0]), __ldg(&light->color[1]), __ldg(&light->color[2])};
                        const double normal_light = fmax(dot(normal, light_dir), 0.0);
                        double specular[3] = {0.0, 0.0, 0.0};
                        if (has_spec) {
                            const V3 half = normalized(view + light_dir);
                            const double normal_half_shiny = pow(fmax(dot(normal, half), 0.0), 4.0 * shininess);
                            specular[0] = spec[0] * lc[0] * normal_half_shiny;
                            specular[1] = spec[1] * lc[1] * normal_half_shiny;
                            specular[2] = spec[2] * lc[2] * normal_half_shiny;
                        }
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const double diffuse = kd[ch] * lc[ch] * normal_light;
                            color[ch] += (diffuse + specular[ch]) / attenuation;
                        }
                    }

                    refl = __ldg(&mat->reflectivity);
                    if (refl > 0.0) {  // material.rs:216-317
                        V3 reflect_dir = ray_dir - (normal * 2.0) * dot(ray_dir, normal);
                        const double g = __ldg(&mat->glossy_side_length);
                        if (g > 0.0) {  // material.rs:220-239; the result is NOT renormalised
                            const V3 offset_vector = (fabs(reflect_dir.x) < kEps && fabs(reflect_dir.y) < kEps)
                                                         ? reflect_dir + v3(0.0, 0.1, 0.0)
                                                         : reflect_dir + v3(0.0, 0.0, 0.1);
                            const V3 u_basis = cross(reflect_dir, offset_vector);
                            const V3 v_basis = cross(reflect_dir, u_basis);
                            const double u_coord = -g / 2.0 + draw(fp.rng_mode, fp.seed, pixel, sample, pathid, 2 + 2 * sc.n_lights) * g;
                            const double v_coord = -g / 2.0 + draw(fp.rng_mode, fp.seed, pixel, sample, pathid, 3 + 2 * sc.n_lights) * g;
                            reflect_dir = reflect_dir + (u_basis * u_coord + v_basis * v_coord);
                        }
                        dir0 = reflect_dir;
                        want0 = true;
                        mode = kModeReflect;
                        const double ior = __ldg(&mat->refraction_index);
                        if (ior > 0.0) {
                            V3 refract_dir;
                            double cos_incident = 0.0;
                            bool have = false;
                            if (dot(ray_dir, normal) < 0.0) {
                                if (refracted_direction(ray_dir, normal, ior, refract_dir)) {
                                    cos_incident = dot(-ray_dir, normal);
                                    have = true;
                                } else {
                                    err |= PT_DEVERR_TIR;  // material.rs:258
                                }
                            } else if (refracted_direction(ray_dir, -normal, 1.0 / ior, refract_dir)) {
                                cos_incident = dot(refract_dir, normal);
                                have = true;
                            }  // else: total internal reflection -> reflect-only combine (material.rs:277-284)
                            if (have) {
                                double r0 = (ior - 1.0) * (ior - 1.0);
                                r0 = r0 / ((ior + 1.0) * (ior + 1.0));
                                const double base = 1.0 - cos_incident;
                                double p4 = base * base;  // powi(5) = base * (base^2)^2
                                p4 = p4 * p4;
                                fres = r0 + (1.0 - r0) * (base * p4);
                                dir1 = refract_dir;
                                want1 = true;
                                mode = kModeDielectric;
                            }
                        }
                    }
                }
            }

            if (err & ~err_seen & ~PT_DEVERR_OVERFLOW) {
                record_error(fp, pool, ctl, err & ~err_seen & ~PT_DEVERR_OVERFLOW, i, 2u | level << 8);
                err_seen = err;
            }
            // depth cut-off: a child at depth > max_depth is bg whatever it hits (material.rs:102-104)
            if (level + 1 > fp.max_depth) {
                if (want0) { c0 = kChildBg; ++n_cut; want0 = false; }
                if (want1) { c1 = kChildBg; ++n_cut; want1 = false; }
            }
            pool.cr[i] = color[0]; pool.cg[i] = color[1]; pool.cb[i] = color[2];
            pool.mode[i] = mode;
            pool.refl[i] = refl;
            pool.fres[i] = fres;
            pool.child0[i] = c0;
            pool.child1[i] = c1;
        }

        // what this ray asks for goes into the tile (path ids: reflected = parent << 1, refracted = parent << 1 | 1,
        // material.rs:243 before :303 — they do not depend on where the children are stored)
        const unsigned m0 = __ballot_sync(0xFFFFFFFFu, want0), m1 = __ballot_sync(0xFFFFFFFFu, want1);
        if (lane == 0) { tile_mem.mask[slot >> 5] = m0; tile_mem.mask[Tile::W + (slot >> 5)] = m1; }
        if (want0 || want1) {
            double* d = tile_mem.d + slot;
            d[0 * Tile::T] = hit_point.x; d[1 * Tile::T] = hit_point.y; d[2 * Tile::T] = hit_point.z;
            d[3 * Tile::T] = dir0.x; d[4 * Tile::T] = dir0.y; d[5 * Tile::T] = dir0.z;
            d[6 * Tile::T] = dir1.x; d[7 * Tile::T] = dir1.y; d[8 * Tile::T] = dir1.z;
            tile_mem.u[slot] = root;
            tile_mem.u[Tile::T + slot] = pathid;
        }
      }
      __syncthreads();
      // one allocation for the tile: exclusive prefix of the warp-rounds' counts (warp 0), then the two atomics
      if (threadIdx.x < 32) {
          uint32_t run0 = 0, run1 = 0;
          for (int w0 = 0; w0 < Tile::W; w0 += 32) {
              const int w = w0 + lane;
              const uint32_t c0 = w < Tile::W ? (uint32_t)__popc(tile_mem.mask[w]) : 0u, c1 = w < Tile::W ? (uint32_t)__popc(tile_mem.mask[Tile::W + w]) : 0u;
              uint32_t i0 = c0, i1 = c1;  // inclusive scans over the 32 lanes
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                  const uint32_t a0 = __shfl_up_sync(0xFFFFFFFFu, i0, off), a1 = __shfl_up_sync(0xFFFFFFFFu, i1, off);
                  if (lane >= off) { i0 += a0; i1 += a1; }
              }
              if (w < Tile::W) { tile_mem.pre[w] = run0 + i0 - c0; tile_mem.pre[Tile::W + w] = run1 + i1 - c1; }
              run0 += __shfl_sync(0xFFFFFFFFu, i0, 31);
              run1 += __shfl_sync(0xFFFFFFFFu, i1, 31);
          }
          if (lane == 0) {
              // reflected children ascend from the bottom of the pool, refracted children descend from its top
              s_base[0] = run0 ? atomicAdd(&ctl->pool_count, run0) : 0u;
              const uint32_t hi = run1 ? atomicAdd(&ctl->hi_count, run1) : 0u;
              // Both counters keep counting past each other when the pool is full (their sum is what the retry sizes its
              // batches by).  A slot is usable only while the two ends have not met: each side reads the OTHER end after
              // its own allocation — whichever of two colliding allocations comes second sees the first.
              const uint32_t hi_now = atomicAdd(&ctl->hi_count, 0u), lo_now = atomicAdd(&ctl->pool_count, 0u);
              s_limit[0] = pool.capacity - min(hi_now, pool.capacity);  // reflected slots must lie below this index
              s_limit[1] = lo_now;                                      // refracted slots must lie at or above it
              // the tile's refracted children in parent order inside its slice [capacity - (hi + run1), capacity - hi)
              const uint64_t end_excl = (uint64_t)hi + run1;
              s_base[1] = end_excl <= pool.capacity ? pool.capacity - (uint32_t)end_excl : kNone;
          }
      }
      __syncthreads();
      for (int k = 0; k < K; ++k) {
        const int slot = k * BLOCK + (int)threadIdx.x;
        const uint32_t j = tile * (uint32_t)Tile::T + (uint32_t)slot;
        if (j >= n) continue;
        const unsigned m0 = tile_mem.mask[slot >> 5], m1 = tile_mem.mask[Tile::W + (slot >> 5)];
        const bool want0 = (m0 >> lane) & 1u, want1 = (m1 >> lane) & 1u;
        if (!want0 && !want1) continue;
        const uint32_t i = rays.at(j);
        const unsigned lt = (1u << lane) - 1u;
        const double* d = tile_mem.d + slot;
        const uint32_t root = tile_mem.u[slot], pathid = tile_mem.u[Tile::T + slot];
        if (want0) {
            const uint32_t n0 = s_base[0] + tile_mem.pre[slot >> 5] + (uint32_t)__popc(m0 & lt);
            if (n0 < s_limit[0]) {
                pool.ox[n0] = d[0 * Tile::T]; pool.oy[n0] = d[1 * Tile::T]; pool.oz[n0] = d[2 * Tile::T];
                pool.dx[n0] = d[3 * Tile::T]; pool.dy[n0] = d[4 * Tile::T]; pool.dz[n0] = d[5 * Tile::T];
                pool.root[n0] = root;
                pool.pathid[n0] = pathid << 1;
                pool.child0[i] = n0;
                ++n_reflect;
            } else {
                err |= PT_DEVERR_OVERFLOW;
                pool.child0[i] = kChildBg;
            }
        }
        if (want1) {
            const uint32_t n1 = s_base[1] == kNone ? kNone : s_base[1] + tile_mem.pre[Tile::W + (slot >> 5)] + (uint32_t)__popc(m1 & lt);
            if (n1 != kNone && n1 >= s_limit[1]) {
                pool.ox[n1] = d[0 * Tile::T]; pool.oy[n1] = d[1 * Tile::T]; pool.oz[n1] = d[2 * Tile::T];
                pool.dx[n1] = d[6 * Tile::T]; pool.dy[n1] = d[7 * Tile::T]; pool.dz[n1] = d[8 * Tile::T];
                pool.root[n1] = root;
                pool.pathid[n1] = (pathid << 1) | 1u;
                pool.child1[i] = n1;
                ++n_refract;
            } else {
                err |= PT_DEVERR_OVERFLOW;
                pool.child1[i] = kChildBg;
            }
        }
      }
      __syncthreads();  // the tile's shared memory is reused
    }

    if (err) atomicOr(&ctl->error_bits, err);
    unsigned long long v[5] = {n_shaded, n_reflect, n_refract, n_cut, n_texel};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xFFFFFFFFu, v[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (v[0]) atomicAdd(&ctl->shaded_hits, v[0]);
        if (v[1]) atomicAdd(&ctl->rays_reflect, v[1]);
        if (v[2]) atomicAdd(&ctl->rays_refract, v[2]);
        if (v[3]) atomicAdd(&ctl->rays_depth_cut, v[3]);
        if (v[4]) atomicAdd(&ctl->texel_lookups, v[4]);
    }

    // the last block to finish publishes where the next level ends
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t ticket = atomicAdd(&ctl->blocks_done[level], 1u);
        if (ticket == gridDim.x - 1) {
            const uint32_t count = atomicAdd(&ctl->pool_count, 0u), hi = atomicAdd(&ctl->hi_count, 0u);
            uint32_t next_end = count < pool.capacity ? count : pool.capacity;
            uint32_t next_hi = hi < pool.capacity ? hi : pool.capacity;
            const uint32_t end = ctl->level_start[level + 1];
            if ((uint64_t)next_end + next_hi > pool.capacity) {
                // the two ends have met: the batch is redone with a smaller size (api.cu), its next level gets no rays
                // (the counters above keep what was asked for)
                atomicOr(&ctl->error_bits, PT_DEVERR_OVERFLOW);
                next_end = end;
                next_hi = ctl->level_hi[level];
            }
            ctl->level_start[level + 2] = next_end;
            ctl->level_hi[level + 1] = next_hi;
            ctl->level = level + 1;
            ctl->levels_run = level + 1;
            ctl->cursor[0] = 0u;
            ctl->cursor[1] = 0u;
            const bool more = (next_end > end || next_hi > ctl->level_hi[level]) && level + 1 < fs.n_levels;
            if (loop) cudaGraphSetConditional(loop, more ? 1u : 0u);
        }
    }
}

// ------------------------------------------------------------------ tree evaluation
// Colour of one path = post-order walk of its ray tree with the reference's own
// expressions (material.rs:280,307-309,315); writes the result over the root's local colour.
__global__ void __launch_bounds__(kBlock) tree_eval_kernel(int slot_id) {
    const FrameState& fs = c_state[slot_id];
    const FrameParams& fp = fs.fp;
    const NodePool& pool = fs.pool;
    const uint32_t first_slot = fs.ctl->first_slot, n_paths = fs.ctl->n_paths;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    if (pool.mode[p] == kModeLeaf) return;  // colour already final

    const uint32_t pixel = fp.pixel_index ? __ldg(fp.pixel_index + first_slot + p / fp.samples) : first_slot + p;
    double bg[3];
    background_of(fp, pixel, bg);

    constexpr int kMaxFrames = PT_MAX_DEPTH_SUPPORTED + 2;  // a path's tree has max_depth + 1 levels; one spare
    uint32_t f_node[kMaxFrames];
    uint8_t f_stage[kMaxFrames];
    double f_refl[kMaxFrames][3];  // reflected_color once child0 is done
    int sp = 0;
    f_node[0] = p;
    f_stage[0] = 0;
    double ret[3] = {0, 0, 0};  // colour returned by the frame that just finished
    for (;;) {
        const uint32_t node = f_node[sp];
        const uint8_t mode = pool.mode[node];
        const uint8_t stage = f_stage[sp];
        if (stage == 0) {
            if (mode == kModeLeaf) {
                ret[0] = pool.cr[node]; ret[1] = pool.cg[node]; ret[2] = pool.cb[node];
            } else {
                const uint32_t c0 = pool.child0[node];
                f_stage[sp] = 1;
                if (c0 == kChildBg) {
                    ret[0] = bg[0]; ret[1] = bg[1]; ret[2] = bg[2];
                } else {
                    ++sp;
                    f_node[sp] = c0;
                    f_stage[sp] = 0;
                }
                continue;
            }
        } else if (stage == 1) {  // ret = reflected_color
            if (mode == kModeReflect) {
                const double refl = pool.refl[node];
                ret[0] = pool.cr[node] + refl * ret[0];
                ret[1] = pool.cg[node] + refl * ret[1];
                ret[2] = pool.cb[node] + refl * ret[2];
            } else {
                f_refl[sp][0] = ret[0]; f_refl[sp][1] = ret[1]; f_refl[sp][2] = ret[2];
                const uint32_t c1 = pool.child1[node];
                f_stage[sp] = 2;
                if (c1 == kChildBg) {
                    ret[0] = bg[0]; ret[1] = bg[1]; ret[2] = bg[2];
                } else {
                    ++sp;
                    f_node[sp] = c1;
                    f_stage[sp] = 0;
                }
                continue;
            }
        } else {  // stage 2: ret = refracted_color, material.rs:292-309
            const double refl = pool.refl[node];
            const double reflectivity = pool.fres[node];
            const double transmittance = 1.0 - reflectivity;
            const double local[3] = {pool.cr[node], pool.cg[node], pool.cb[node]};
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const double total_color = reflectivity * f_refl[sp][ch] + transmittance * ret[ch];
                ret[ch] = local[ch] + refl * total_color;
            }
        }
        // frame finished with `ret`
        if (sp == 0) break;
        --sp;
    }
    pool.cr[p] = ret[0]; pool.cg[p] = ret[1]; pool.cb[p] = ret[2];
}

// ------------------------------------------------------------------ resolve
PT_D uint8_t to_u8(double c) {  // Rust `as u8`: truncate, saturate, NaN -> 0. render.rs:143-147
    const double v = c * 255.0;
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}
PT_D double clamp01(double v) {
    const double lo = v >= 0.0 ? v : 0.0;
    return lo <= 1.0 ? lo : 1.0;
}

// one thread per owned pixel of the batch: render.rs:43-50,143-147
__global__ void __launch_bounds__(kBlock) resolve_kernel(int slot_id) {
    const FrameState& fs = c_state[slot_id];
    const FrameParams& fp = fs.fp;
    const NodePool& pool = fs.pool;
    const uint32_t first_slot = fs.ctl->first_slot, n_slots = fs.ctl->n_slots;
    uint8_t* __restrict__ rgb = fs.rgb;
    uint32_t* __restrict__ hit_id = fs.hit_id;
    double* __restrict__ hit_t = fs.hit_t;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < n_slots;  // (no early return: the warp's lanes cooperate on the image stores below)
    const uint32_t p0 = k * fp.samples;
    double total[3] = {0.0, 0.0, 0.0};
    if (active) {
        for (uint32_t s = 0; s < fp.samples; ++s) {
            total[0] = total[0] + pool.cr[p0 + s];
            total[1] = total[1] + pool.cg[p0 + s];
            total[2] = total[2] + pool.cb[p0 + s];
        }
    }
    // compact owned-pixel order (multi-GPU gather), or the pixel's own place in a full row-major image
    const size_t own_pixel = active ? (size_t)__ldg(fp.pixel_index + first_slot + k) : 0;
    const size_t slot = fs.row_major ? own_pixel : (size_t)first_slot + k;
    uint32_t packed = 0;  // r | g << 8 | b << 16
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        double c = total[ch] / (double)fp.samples;
        c = pow(c, 1.0 / PT_GAMMA);
        const uint8_t v = to_u8(clamp01(c));
        if (active) rgb[slot * 3 + ch] = v;
        packed |= (uint32_t)v << (8 * ch);
    }
    // The exchange step of the multi-GPU path, fused: the pixel also goes to its own place in the collecting device's
    // full image (peer memory over NVLink for the other members); tiles are disjoint, so no two devices write one byte.
    // Owned pixels are listed in 8 x 4 micro-tiles (tiles.c): 4 consecutive lanes hold 4 consecutive pixels of a row =
    // 12 contiguous bytes, written as THREE aligned 32-bit words by three of the lanes.  Byte stores (three per pixel)
    // cost 34 MB of NVLink traffic for the 3 MB of a 1080p half frame (ncu nvltx__bytes, profiles/r02_peer_stores.txt).
    uint8_t* __restrict__ image = fs.rgb_image;
    if (image) {  // uniform over the grid
        const unsigned full = 0xFFFFFFFFu;
        const int lane = threadIdx.x & 31, j = lane & 3, lead = lane & ~3;
        const unsigned long long px = active ? (unsigned long long)own_pixel : ~0ull;
        const unsigned long long px_lead = __shfl_sync(full, px, lead);
        const unsigned ok = __ballot_sync(full, active && px == px_lead + (unsigned)j);
        const uint32_t c0 = __shfl_sync(full, packed, lead), c1 = __shfl_sync(full, packed, lead + 1);
        const uint32_t c2 = __shfl_sync(full, packed, lead + 2), c3 = __shfl_sync(full, packed, lead + 3);
        const bool group = ((ok >> lead) & 0xFu) == 0xFu && (((size_t)px_lead * 3 + (size_t)image) & 3u) == 0;
        if (group) {
            if (j < 3) {
                const uint32_t word = j == 0 ? (c0 | (c1 & 0xFFu) << 24) : j == 1 ? ((c1 >> 8) | (c2 & 0xFFFFu) << 16) : ((c2 >> 16) | c3 << 8);
                *reinterpret_cast<uint32_t*>(image + (size_t)px_lead * 3 + 4 * j) = word;
            }
        } else if (active) {
            image[own_pixel * 3] = (uint8_t)packed;
            image[own_pixel * 3 + 1] = (uint8_t)(packed >> 8);
            image[own_pixel * 3 + 2] = (uint8_t)(packed >> 16);
        }
    }
    if (!active) return;
    if (hit_id) {
        hit_id[slot * 2] = pool.inst[p0];
        hit_id[slot * 2 + 1] = pool.sub[p0];
    }
    if (hit_t) hit_t[slot] = pool.t[p0];
}

// pt_trace_rays: linear colour of each explicit ray, no gamma
__global__ void __launch_bounds__(kBlock) export_rays_kernel(int slot_id) {
    const FrameState& fs = c_state[slot_id];
    const NodePool& pool = fs.pool;
    const uint32_t first = fs.ctl->first_slot, n_paths = fs.ctl->n_paths;
    double* __restrict__ color = fs.ray_color;
    uint32_t* __restrict__ hit_id = fs.hit_id;
    double* __restrict__ hit_t = fs.hit_t;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const size_t i = (size_t)first + p;
    color[i * 3] = pool.cr[p]; color[i * 3 + 1] = pool.cg[p]; color[i * 3 + 2] = pool.cb[p];
    hit_id[i * 2] = pool.inst[p];
    hit_id[i * 2 + 1] = pool.sub[p];
    hit_t[i] = pool.t[p];
}

// ------------------------------------------------------------------ instance bounds (upload time)
// Object-space bounds of every mesh: one block per mesh, min/max over its triangle vertices.
__global__ void __launch_bounds__(kBlock) mesh_bounds_kernel(const PtMesh* __restrict__ meshes, const PtTriPos* __restrict__ tri_pos,
                                                            uint32_t n_meshes, double* __restrict__ out /* [n_meshes][6] */) {
    const uint32_t m = blockIdx.x;
    if (m >= n_meshes) return;
    const uint32_t first = meshes[m].tri_first, count = meshes[m].tri_count;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t k = threadIdx.x; k < count; k += blockDim.x) {
        const double* v = reinterpret_cast<const double*>(tri_pos + first + k);
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const double x = v[c];
            lo[c % 3] = fmin(lo[c % 3], x);
            hi[c % 3] = fmax(hi[c % 3], x);
        }
    }
    __shared__ double s_lo[kBlock / 32][3], s_hi[kBlock / 32][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[c] = fmin(lo[c], __shfl_down_sync(0xFFFFFFFFu, lo[c], off));
            hi[c] = fmax(hi[c], __shfl_down_sync(0xFFFFFFFFu, hi[c], off));
        }
        if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5][c] = lo[c]; s_hi[threadIdx.x >> 5][c] = hi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        double l = s_lo[0][c], h = s_hi[0][c];
        for (int w = 1; w < kBlock / 32; ++w) { l = fmin(l, s_lo[w][c]); h = fmax(h, s_hi[w][c]); }
        out[m * 6 + c] = l;
        out[m * 6 + 3 + c] = h;
    }
}

// Padded world-space box of every flat instance, FP32, rounded outward (the FP32 cull of traverse.cuh).
// Object boxes: unit sphere [-1,1]^3 (sphere.rs), everything else analytic [-0.5,0.5]^3 (cube / cylinder / cone;
// the plane is the y = 0 slice of it), meshes their vertex bounds; all grown by 1e-4 of their size, which
// covers the 1e-5 slack of Cube::contains / Plane (cube.rs:22-27, plane.rs:43).
__global__ void __launch_bounds__(kBlock) instance_bounds_kernel(const PtInstance* __restrict__ instances,
                                                                const PtInstanceTrans* __restrict__ trans, uint32_t n,
                                                                const double* __restrict__ mesh_bounds, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t prim = instances[i].prim;
    double lo[3], hi[3];
    if (prim == PT_PRIM_SPHERE) {
        lo[0] = lo[1] = lo[2] = -1.0; hi[0] = hi[1] = hi[2] = 1.0;
    } else if (prim == PT_PRIM_TRIANGLE || prim == PT_PRIM_MESH || prim == PT_PRIM_KDMESH) {
        const double* b = mesh_bounds + (size_t)instances[i].mesh * 6;
        for (int c = 0; c < 3; ++c) { lo[c] = b[c]; hi[c] = b[3 + c]; }
    } else {
        lo[0] = lo[1] = lo[2] = -0.5; hi[0] = hi[1] = hi[2] = 0.5;
        if (prim == PT_PRIM_PLANE) { lo[1] = 0.0; hi[1] = 0.0; }
    }
    double diag = 0.0;
    for (int c = 0; c < 3; ++c) diag = fmax(diag, hi[c] - lo[c]);
    for (int c = 0; c < 3; ++c) {
        const double g = 1e-4 * diag + 1e-4 * fmax(fabs(lo[c]), fabs(hi[c]));
        lo[c] -= g;
        hi[c] += g;
    }
    const double* m = trans[i].trans;  // object -> world, rows 0..2
    double wlo[3] = {INFINITY, INFINITY, INFINITY}, whi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool finite = true;
    for (int corner = 0; corner < 8; ++corner) {
        const double x = (corner & 1) ? hi[0] : lo[0], y = (corner & 2) ? hi[1] : lo[1], z = (corner & 4) ? hi[2] : lo[2];
        for (int r = 0; r < 3; ++r) {
            const double w = m[r * 4 + 0] * x + m[r * 4 + 1] * y + m[r * 4 + 2] * z + m[r * 4 + 3];
            if (!isfinite(w)) finite = false;
            wlo[r] = fmin(wlo[r], w);
            whi[r] = fmax(whi[r], w);
        }
    }
    float flo[3], fhi[3];
    double wdiag = 0.0;
    for (int r = 0; r < 3; ++r) wdiag = fmax(wdiag, whi[r] - wlo[r]);
    for (int r = 0; r < 3; ++r) {
        const double g = 1e-5 * fmax(fabs(wlo[r]), fabs(whi[r])) + 1e-5 * wdiag + 1e-30;
        flo[r] = __double2float_rd(wlo[r] - g);
        fhi[r] = __double2float_ru(whi[r] + g);
        if (!finite || !(wlo[r] <= whi[r])) { flo[r] = -INFINITY; fhi[r] = INFINITY; }  // never cull what cannot be bounded
    }
    out[2 * (size_t)i] = make_float4(flo[0], flo[1], flo[2], 0.f);
    out[2 * (size_t)i + 1] = make_float4(fhi[0], fhi[1], fhi[2], 0.f);
}

// FP32 box of every triangle (object space), rounded outward and padded by 1e-5 of the triangle's largest
// coordinate magnitude and extent: the per-triangle cull of Mesh folds (traverse.cuh mesh_fold).
__global__ void __launch_bounds__(kBlock) triangle_bounds_kernel(const PtTriPos* __restrict__ tri_pos, uint32_t n, float4* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* v = reinterpret_cast<const double*>(tri_pos + k);
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    double mag = 0.0;
    bool finite = true;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        const double x = v[c];
        if (!isfinite(x)) finite = false;
        lo[c % 3] = fmin(lo[c % 3], x);
        hi[c % 3] = fmax(hi[c % 3], x);
        mag = fmax(mag, fabs(x));
    }
    double diag = 0.0;
    for (int c = 0; c < 3; ++c) diag = fmax(diag, hi[c] - lo[c]);
    const double g = 1e-5 * mag + 1e-5 * diag + 1e-30;
    float flo[3], fhi[3];
    for (int c = 0; c < 3; ++c) {
        flo[c] = __double2float_rd(lo[c] - g);
        fhi[c] = __double2float_ru(hi[c] + g);
        if (!finite) { flo[c] = -INFINITY; fhi[c] = INFINITY; }  // never cull what cannot be bounded
    }
    out[2 * (size_t)k] = make_float4(flo[0], flo[1], flo[2], 0.f);
    out[2 * (size_t)k + 1] = make_float4(fhi[0], fhi[1], fhi[2], 0.f);
}

// union of every aligned run of `run` boxes of `in` (n boxes) -> out[ceil(n / run)]
__global__ void __launch_bounds__(kBlock) group_bounds_kernel(const float4* __restrict__ in, uint32_t n, uint32_t run, float4* __restrict__ out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_groups = (n + run - 1u) / run;
    if (g >= n_groups) return;
    float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    const uint32_t end = min(n, (g + 1u) * run);
    for (uint32_t k = g * run; k < end; ++k) {
        const float4 a = in[2 * (size_t)k], b = in[2 * (size_t)k + 1];
        lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
        hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
    }
    out[2 * (size_t)g] = lo;
    out[2 * (size_t)g + 1] = hi;
}

// pow(i / 255, 2.2) for the 256 possible texel values, with the device's own pow(): bit-identical to evaluating it per hit
__global__ void gamma_lut_kernel(double* __restrict__ lut) {
    const int i = threadIdx.x;
    lut[i] = pow((double)i / 255.0, PT_GAMMA);
}

// FP64 issue-rate microbenchmark: 8 independent multiply-then-add chains per thread.  Built with -fmad=false like
// everything else here, so each step is one DMUL and one DADD — the instruction mix of the render kernels — and the
// result is the no-FMA f64 ceiling that their roofline is quoted against.
__global__ void __launch_bounds__(256) fp64_rate_kernel(double* __restrict__ sink, int iters, double m, double c) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) sink[0] = s;  // never true: keeps the chains alive
}

int g_grid_extend[3] = {0, 0, 0}, g_grid_shadow[3] = {0, 0, 0};  // [kWalkExact | kWalkCount | kWalkPrune]
int g_grid_shade[2] = {0, 0};                                     // [small group | big group]

template <class K>
int persistent_grid(K kernel, int block = kBlock, size_t dynamic_smem = 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dynamic_smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynamic_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dynamic_smem);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

}  // namespace

// ------------------------------------------------------------------ launch wrappers
void kernels_init() {
    g_grid_extend[0] = persistent_grid(extend_kernel<false>);
    g_grid_extend[1] = persistent_grid(extend_kernel<true>);
    g_grid_shadow[0] = persistent_grid(shadow_kernel<false>);
    g_grid_shadow[1] = persistent_grid(shadow_kernel<true>);
    g_grid_extend[2] = persistent_grid(extend_kernel<false, false, true>);
    g_grid_shadow[2] = persistent_grid(shadow_kernel<false, false, true>);
    g_grid_shade[0] = persistent_grid(shade_kernel<kShadeBlockSmall, 1>, kShadeBlockSmall, ShadeTile<kShadeBlockSmall, 1>::kBytes);
    g_grid_shade[1] = persistent_grid(shade_kernel<kShadeBlockBig, kShadeRoundsBig>, kShadeBlockBig, ShadeTile<kShadeBlockBig, kShadeRoundsBig>::kBytes);
}

static inline uint32_t blocks_for(uint64_t n) { return (uint32_t)((n + kBlock - 1) / kBlock); }
static inline int capped(int grid, uint64_t max_items, int block = kBlock) {
    const uint64_t need = ((max_items ? max_items : 1) + block - 1) / block;
    return (uint64_t)grid > need ? (int)need : grid;
}

// returns the flops issued (2 per chain step)
double launch_fp64_rate(double* sink, int iters, cudaStream_t st) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8;
    fp64_rate_kernel<<<blocks, 256, 0, st>>>(sink, iters, 0.999999, 1e-9);
    return (double)blocks * 256.0 * (double)iters * 8.0 * 2.0;
}

void launch_gamma_lut(double* lut, cudaStream_t st) { gamma_lut_kernel<<<1, 256, 0, st>>>(lut); }

void launch_instance_bounds(const DScene& sc, uint32_t n_meshes, double* mesh_bounds_scratch, float4* out, cudaStream_t st) {
    if (n_meshes) mesh_bounds_kernel<<<n_meshes, kBlock, 0, st>>>(sc.meshes, sc.tri_pos, n_meshes, mesh_bounds_scratch);
    if (sc.n_instances)
        instance_bounds_kernel<<<blocks_for(sc.n_instances), kBlock, 0, st>>>(sc.instances, sc.instance_trans, sc.n_instances,
                                                                            mesh_bounds_scratch, out);
}

void launch_triangle_bounds(const PtTriPos* tri_pos, uint32_t n, float4* tri_aabb, cudaStream_t st) {
    if (!n) return;
    triangle_bounds_kernel<<<blocks_for(n), kBlock, 0, st>>>(tri_pos, n, tri_aabb);
}

// out[ceil(n / run)] = union of every aligned run of `run` boxes of in[n]
void launch_group_bounds(const float4* in, uint32_t n, uint32_t run, float4* out, cudaStream_t st) {
    if (!n) return;
    group_bounds_kernel<<<blocks_for((n + run - 1u) / run), kBlock, 0, st>>>(in, n, run, out);
}

cudaError_t upload_state(int slot, const FrameState& state, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_state, &state, sizeof(FrameState), (size_t)slot * sizeof(FrameState),
                                   cudaMemcpyHostToDevice, st);
}

void launch_camera(int slot, uint32_t first_slot, uint32_t n_slots, uint32_t samples, cudaStream_t st) {
    camera_kernel<<<blocks_for((uint64_t)n_slots * samples), kBlock, 0, st>>>(slot, first_slot, n_slots);
}
void launch_load_rays(int slot, uint32_t first, uint32_t n_paths, cudaStream_t st) {
    load_rays_kernel<<<blocks_for(n_paths), kBlock, 0, st>>>(slot, first, n_paths);
}
// A level can never hold more rays than `max_items`; the persistent grid is capped to it.
void launch_extend(int slot, uint64_t max_items, int mode, bool linear, cudaStream_t st) {
    const int grid = capped(g_grid_extend[mode], max_items);
    if (linear) extend_kernel<true, true><<<grid, kBlock, 0, st>>>(slot);
    else if (mode == kWalkCount) extend_kernel<true><<<grid, kBlock, 0, st>>>(slot);
    else if (mode == kWalkPrune) extend_kernel<false, false, true><<<grid, kBlock, 0, st>>>(slot);
    else extend_kernel<false><<<grid, kBlock, 0, st>>>(slot);
}
void launch_shadow(int slot, uint64_t max_items, uint32_t n_lights, int mode, bool linear, cudaStream_t st) {
    const int grid = capped(g_grid_shadow[mode], max_items * (n_lights ? n_lights : 1));
    if (linear) shadow_kernel<true, true><<<grid, kBlock, 0, st>>>(slot);
    else if (mode == kWalkCount) shadow_kernel<true><<<grid, kBlock, 0, st>>>(slot);
    else if (mode == kWalkPrune) shadow_kernel<false, false, true><<<grid, kBlock, 0, st>>>(slot);
    else shadow_kernel<false><<<grid, kBlock, 0, st>>>(slot);
}
bool shade_big_group(uint64_t batch_paths) { return batch_paths >= kShadeBigPaths; }
void launch_shade(int slot, uint64_t max_items, bool big, cudaGraphConditionalHandle loop, cudaStream_t st) {
    if (big)
        shade_kernel<kShadeBlockBig, kShadeRoundsBig><<<capped(g_grid_shade[1], max_items, kShadeBlockBig * kShadeRoundsBig), kShadeBlockBig,
                                                        ShadeTile<kShadeBlockBig, kShadeRoundsBig>::kBytes, st>>>(slot, loop);
    else
        shade_kernel<kShadeBlockSmall, 1><<<capped(g_grid_shade[0], max_items, kShadeBlockSmall), kShadeBlockSmall, ShadeTile<kShadeBlockSmall, 1>::kBytes, st>>>(slot, loop);
}
void launch_tree_eval(int slot, uint32_t n_paths, cudaStream_t st) {
    tree_eval_kernel<<<blocks_for(n_paths), kBlock, 0, st>>>(slot);
}
void launch_resolve(int slot, uint32_t n_slots, cudaStream_t st) {
    resolve_kernel<<<blocks_for(n_slots), kBlock, 0, st>>>(slot);
}
void launch_export_rays(int slot, uint32_t n_paths, cudaStream_t st) {
    export_rays_kernel<<<blocks_for(n_paths), kBlock, 0, st>>>(slot);
}

// The frame graph: camera -> WHILE(level has rays){extend, shadow, shade} -> tree_eval -> resolve.  Grids are
// sized for a full batch of `n_slots` pixels; the camera node's (first_slot, n_slots) parameters are patched
// per batch (cudaGraphExecKernelNodeSetParams), everything else is read from c_state[slot] / the control block.
cudaError_t build_frame_graph(int slot, uint32_t n_slots, uint32_t samples, uint32_t n_lights_max, uint64_t capacity,
                              int mode, cudaGraph_t* graph_out, cudaGraphExec_t* exec_out, cudaGraphNode_t* camera_node) {
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaGraphCreate(&graph, 0);
    if (e != cudaSuccess) return e;
    const uint64_t n_paths = (uint64_t)n_slots * samples;

    static uint32_t zero_first = 0;
    uint32_t first_slot = zero_first, slots = n_slots;
    void* cam_args[3] = {&slot, &first_slot, &slots};
    cudaKernelNodeParams kp{};
    kp.func = (void*)camera_kernel;
    kp.gridDim = dim3(blocks_for(n_paths));
    kp.blockDim = dim3(kBlock);
    kp.kernelParams = cam_args;
    cudaGraphNode_t n_camera = nullptr, n_loop = nullptr, n_tree = nullptr, n_resolve = nullptr;
    e = cudaGraphAddKernelNode(&n_camera, graph, nullptr, 0, &kp);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }

    cudaGraphConditionalHandle handle;
    e = cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    e = cudaGraphAddNode(&n_loop, graph, &n_camera, 1, &cp);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    {
        // one recursion level; a level never holds more rays than the node pool
        void* args1[1] = {&slot};
        cudaGraphNode_t n_ext = nullptr, n_shd = nullptr, n_sha = nullptr;
        cudaKernelNodeParams k{};
        k.blockDim = dim3(kBlock);
        k.kernelParams = args1;
        k.func = mode == kWalkCount ? (void*)extend_kernel<true> : mode == kWalkPrune ? (void*)extend_kernel<false, false, true> : (void*)extend_kernel<false>;
        k.gridDim = dim3(capped(g_grid_extend[mode], capacity));
        e = cudaGraphAddKernelNode(&n_ext, body, nullptr, 0, &k);
        if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }
        k.func = mode == kWalkCount ? (void*)shadow_kernel<true> : mode == kWalkPrune ? (void*)shadow_kernel<false, false, true> : (void*)shadow_kernel<false>;
        k.gridDim = dim3(capped(g_grid_shadow[mode], capacity * (n_lights_max ? n_lights_max : 1)));
        e = cudaGraphAddKernelNode(&n_shd, body, &n_ext, 1, &k);
        if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }
        void* args2[2] = {&slot, &handle};
        const bool big = shade_big_group(n_paths);
        k.func = big ? (void*)shade_kernel<kShadeBlockBig, kShadeRoundsBig> : (void*)shade_kernel<kShadeBlockSmall, 1>;
        k.kernelParams = args2;
        k.blockDim = dim3(big ? kShadeBlockBig : kShadeBlockSmall);
        k.gridDim = dim3(big ? capped(g_grid_shade[1], capacity, kShadeBlockBig * kShadeRoundsBig) : capped(g_grid_shade[0], capacity, kShadeBlockSmall));
        k.sharedMemBytes = (unsigned)(big ? ShadeTile<kShadeBlockBig, kShadeRoundsBig>::kBytes : ShadeTile<kShadeBlockSmall, 1>::kBytes);
        e = cudaGraphAddKernelNode(&n_sha, body, &n_shd, 1, &k);
        if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }
    }
    void* args1[1] = {&slot};
    kp.func = (void*)tree_eval_kernel;
    kp.kernelParams = args1;
    kp.gridDim = dim3(blocks_for(n_paths));
    e = cudaGraphAddKernelNode(&n_tree, graph, &n_loop, 1, &kp);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }
    kp.func = (void*)resolve_kernel;
    kp.gridDim = dim3(blocks_for(n_slots));
    e = cudaGraphAddKernelNode(&n_resolve, graph, &n_tree, 1, &kp);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }

    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return e; }
    *graph_out = graph;
    *exec_out = exec;
    *camera_node = n_camera;
    return cudaSuccess;
}

// patch the camera node for one batch (a smaller last batch keeps the grids: surplus threads exit at once)
cudaError_t set_graph_batch(cudaGraphExec_t exec, cudaGraphNode_t camera_node, int slot, uint32_t first_slot, uint32_t n_slots,
                            uint32_t grid_slots, uint32_t samples) {
    void* cam_args[3] = {&slot, &first_slot, &n_slots};
    cudaKernelNodeParams kp{};
    kp.func = (void*)camera_kernel;
    kp.gridDim = dim3(blocks_for((uint64_t)grid_slots * samples));
    kp.blockDim = dim3(kBlock);
    kp.kernelParams = cam_args;
    return cudaGraphExecKernelNodeSetParams(exec, camera_node, &kp);
}

}  // namespace ptd
