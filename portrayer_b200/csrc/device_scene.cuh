// Device view of an uploaded scene blob and of one batch's ray-tree node pool.
#pragma once
#include <cstdint>

#include "device_math.cuh"
#include "portrayer_gpu.h"

namespace ptd {

constexpr int kFoldLevels = 6;  // group levels of the Mesh fold structure: runs of 4, 16, 64, 256, 1024, 4096 positions

// One texture as the kernels see it: the texels live either inside the uploaded blob or in the
// library's texture residency cache (PtTexture.key), so the record carries a device pointer.
struct TextureDev {
    uint32_t width, height;
    const uint8_t* texels;
};

// Leaf cull structure of one forest of k-d trees (leaf_cull.cu): per leaf an "occupied" box, per run of 8 list positions
// a union box, per position the candidate's box clipped to the leaf's cell; the same arrays with unclipped boxes lie
// `set_stride` float4 further on.  A leaf node's `split` field in the device copy of the tree carries rank | gbase << 32.
struct LeafCull {
    float4* occ;          // [2 * leaves]       occ[2 * rank], occ[2 * rank + 1] = lo, hi
    float4* grp;          // [2 * runs]         grp[2 * (gbase + g)]
    float4* item;         // [2 * 8 * runs]     item[2 * ((gbase + g) * 8 + j)]
    float4* node;         // [2 * nodes]        node[2 * i]: union of the occupied boxes of the leaves below node i (forest index)
    uint32_t set_stride;  // in float4
    uint32_t pad_;
};

// Pointers into the scene records as they lie in HBM (the record sections of the blob are
// uploaded verbatim, so an NCCL-broadcast copy is usable as is).
struct DScene {
    const PtKdNode* tlas_nodes;
    const uint32_t* tlas_items;
    const PtInstance* instances;
    const PtInstanceTrans* instance_trans;
    const PtMesh* meshes;
    const PtKdNode* blas_nodes;
    const uint32_t* blas_items;
    const PtTriPos* tri_pos;
    const PtTriNormals* tri_normals;
    const PtTriUvs* tri_uvs;
    const PtMaterial* materials;
    const PtLight* lights;
    const TextureDev* textures;
    const double* gamma_lut;  // [256] pow(i / 255, 2.2) computed once on the device (ImageTexture::at, texture.rs:162-168)
    const float4* inst_aabb;  // [2 * n_instances] padded world-space box of every instance (lo, hi), FP32, rounded outward
    LeafCull tl_cull;         // scene tree: boxes of the leaves' candidates (inst_aabb clipped to the leaf cells)
    LeafCull bl_cull;         // KDMesh trees: boxes of the leaves' triangles (tri_aabb clipped to the leaf cells)
    const float4* tl_root;    // [2] union of all instance boxes (the probe-segment check of traverse.cuh scene_cast)
    // padded object-space FP32 box of every triangle (index order)
    const float4* tri_aabb;
    // Mesh fold cull (traverse.cuh mesh_fold): the triangles of every linear Mesh in Morton order of their centroids
    // (fold_order: position -> triangle index, positions of a mesh stay inside its own [tri_first, tri_first + tri_count)),
    // their boxes in that order (fold_aabb[0]) and the union boxes of every aligned run of 4^l positions (fold_aabb[l])
    const uint32_t* fold_order;
    const float4* fold_aabb[kFoldLevels + 1];
    uint32_t fold_levels;  // group levels actually built (<= kFoldLevels)
    double ambient[3];
    double tlas_extent;
    uint32_t n_lights;
    uint32_t n_instances;
    uint32_t n_tlas_nodes;
    uint32_t n_tlas_items;
};

// child / node sentinels
constexpr uint32_t kChildNone = 0xFFFFFFFFu;  // no such child
constexpr uint32_t kChildBg = 0xFFFFFFFEu;    // child's colour is the pixel background (depth cut-off, material.rs:102-104)

// how a node combines its children (material.rs:280,307-309,315)
constexpr uint8_t kModeLeaf = 0;        // colour = local
constexpr uint8_t kModeReflect = 1;     // colour = local + refl * C(child0)            (mirror, and total internal reflection)
constexpr uint8_t kModeDielectric = 2;  // colour = local + refl * (F*C(child0) + (1-F)*C(child1))

// One batch's ray tree, structure-of-arrays, `capacity` nodes.  Level d of the
// recursion occupies the index range [level_start[d], level_start[d+1]) plus a run at the top of the pool (BatchCtl);
// level 0 is the batch's primary rays, node index == path index.
struct NodePool {
    double *ox, *oy, *oz, *dx, *dy, *dz;  // ray (world space; direction not necessarily unit, material.rs:238-242)
    double* t;                            // closest hit parameter
    uint32_t* inst;                       // flat instance index or kNone
    uint32_t* sub;                        // triangle / face / part id
    uint32_t* root;                       // path index (-> pixel, sample)
    uint32_t* pathid;                     // 1 for the primary ray; child = parent << 1 | (0 reflected, 1 refracted)
    double *cr, *cg, *cb;                 // local colour (ambient + unshadowed lights), later the subtree colour of roots
    double* refl;                         // material.reflectivity
    double* fres;                         // Schlick reflectivity of dielectrics
    uint32_t *child0, *child1;
    uint8_t* mode;
    uint8_t* occl;                        // [light][capacity]: 1 = the shadow ray of that light found an occluder
    uint32_t capacity;
};

// device-side control block of a batch
struct BatchCtl {
    uint32_t level_start[16];  // PT_MAX_RECURSION_DEPTH + 2 used
    uint32_t pool_count;       // bump allocator of the node pool
    uint32_t error_bits;       // PT_DEVERR_*
    uint32_t blocks_done[16];  // per level "last block" tickets of the shade kernel
    uint32_t level;            // the recursion level the next extend / shadow / shade launch works on
    uint32_t first_slot;       // first owned-pixel slot of the batch
    uint32_t n_slots;          // owned pixels in the batch
    uint32_t n_paths;          // n_slots * samples
    uint32_t levels_run;       // levels that held at least one ray
    uint32_t cursor[2];        // dynamic work distribution of the current level: next unclaimed item of [0] extend, [1] shadow
    uint32_t pad_[1];
    // where the first device-detected error of the batch happened (the reference's panic location):
    // [0] PT_DEVERR_* bit (0 = none), [1] global pixel, [2] sample, [3] path id, [4] kernel (0 extend, 1 shadow, 2 shade)
    // | level << 8 | light << 16, [5] reserved
    uint32_t err_info[6];
    // The pool is filled from BOTH ends: primary rays and reflected children ascend from 0 (pool_count), refracted
    // children descend from `capacity` (hi_count of them so far).  A level is then two runs — all of its reflected rays,
    // all of its refracted rays — instead of an interleaving in units of the shade kernel's blocks:
    //   lo run  [level_start[d], level_start[d + 1])
    //   hi run  [capacity - level_hi[d], capacity - level_hi[d - 1])      (level_hi[-1] = 0; level 0 has none)
    uint32_t hi_count;
    uint32_t level_hi[16];
    uint32_t pad2_[1];
    unsigned long long rays_shadow, rays_reflect, rays_refract, rays_depth_cut, shaded_hits, texel_lookups;
    // [0 extend | 1 shadow][kd_splits, instance_tests, triangle_tests, bbox_gates, prim_flops (the reference's work),
    //                       box tests, instance tests, triangle tests, bbox gates, prim_flops EXECUTED on the device]
    unsigned long long work[2][10];
};

// per-frame constants
struct FrameParams {
    PtCamera cam;
    const uint32_t* pixel_index;  // owned slot -> global pixel (y*W+x)
    const double* background;
    uint32_t bg_mode;
    uint32_t width, height;
    uint32_t samples;
    uint32_t rng_mode;
    uint64_t seed;
    uint32_t max_depth;
};

// Everything the kernels of one frame read, in __constant__ memory (kernels.cu: c_state[slot]):
// rebinding a frame to another scene / camera is one small async copy, and the frame's CUDA graph
// (whose kernel nodes only carry the slot number) is replayed unchanged.
struct FrameState {
    DScene sc;
    FrameParams fp;
    NodePool pool;
    BatchCtl* ctl;
    uint8_t* rgb;       // resolve outputs: compact owned-pixel order, or full-image row-major
    uint8_t* rgb_image; // nullable: full-image row-major RGB8 target, possibly another GPU's memory (pt_frame_set_image_target)
    uint32_t* hit_id;   // nullable
    double* hit_t;      // nullable
    uint32_t row_major;
    uint32_t n_levels;  // 1 when no material reflects, else max_depth + 1
    // pt_trace_rays: explicit rays in, linear colours out
    const double* ray_origins;
    const double* ray_dirs;
    double* ray_color;
};
constexpr int kStateSlots = 16;

}  // namespace ptd
