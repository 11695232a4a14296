/*
 * portrayer_gpu.h — C ABI of the B200 render-loop library (libportrayer_gpu.so).
 *
 * This is the drop-in boundary for portrayer's per-pixel render loop.  The
 * reference (sunjay/portrayer, pure Rust) has NO FFI today; the cut line is
 * between "scene prepared" and the rayon pixel loop:
 *
 *     src/render.rs:124-126   FlatScene::from + KDTreeScene::from   (host, stays)
 *     src/render.rs:127-150   par_chunks_mut pixel loop              (REPLACED)
 *
 * Everything the replaced loop reads is passed here as plain pointers + sizes:
 * the flattened scene (src/flat_scene.rs:50-61), the kd-tree over it
 * (src/kdtree/node.rs:13-25, src/kdtree/leaf.rs:70-78), per-mesh kd-trees
 * (src/kdtree/kdmesh.rs:19-24), triangle soups (src/primitive/mesh.rs:22-34,
 * src/primitive/triangle.rs:9-19), materials (src/material.rs:51-86), lights
 * (src/light.rs:74-85), textures (src/texture.rs:78-80), the prepared camera
 * (src/camera.rs:17-32) and the per-pixel background colour
 * (src/render.rs:31-34).  The output is the RGB8 buffer of
 * src/render.rs:143-147, written only inside the slice (src/render.rs:136-138).
 *
 * No torch / C++ types appear in any signature.  All functions return 0 on
 * success or a negative PtError; pt_last_error() gives the message (the
 * reference's panic text where the reference would have panicked).
 */
#ifndef PORTRAYER_GPU_H
#define PORTRAYER_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* Constants of the reference that are part of the contract                    */
/* ------------------------------------------------------------------------- */
#define PT_EPSILON 0.00001            /* src/math.rs:15 */
#define PT_GAMMA 2.2                  /* src/math.rs:20 */
#define PT_MAX_RECURSION_DEPTH 10u    /* src/material.rs:12 */
#define PT_MAX_DEPTH_SUPPORTED 13u    /* largest PtRenderParams.max_depth the library accepts (ray trees of 14 levels) */
#define PT_DEFAULT_SAMPLES 100u       /* src/render.rs:113 */

#define PT_MAX_KD_STACK 48            /* deepest kd-tree (KD_DEPTH / KD_MESH_DEPTH) the device walks */
#define PT_MAX_LIGHTS 32

/* ------------------------------------------------------------------------- */
/* Errors                                                                      */
/* ------------------------------------------------------------------------- */
typedef enum PtError {
    PT_OK = 0,
    PT_ERR_INVALID = -1,          /* bad argument / malformed blob */
    PT_ERR_CUDA = -2,             /* CUDA runtime failure (no device, OOM, launch error) */
    PT_ERR_NO_TEXCOORD_NORMALMAP = -3, /* "Normal/Texture mapping is not supported for this primitive!" material.rs:133 */
    PT_ERR_NO_TEXCOORD_TEXTURE = -4,   /* "Texture mapping is not supported for this primitive!" material.rs:141 */
    PT_ERR_KD_PLANE_MISS = -5,    /* "bug: ray should definitely hit infinite plane" kdtree/node.rs:147,178 */
    PT_ERR_TIR_INSIDE = -6,       /* "bug: should not have total internal reflection when casting inside surface" material.rs:258 */
    PT_ERR_KD_TOO_DEEP = -7,      /* kd depth > PT_MAX_KD_STACK */
    PT_ERR_OVERFLOW = -8          /* ray-tree node pool exhausted even at the minimum batch size */
} PtError;

/* bits of the device-side error word (PtStats.device_error_bits) */
#define PT_DEVERR_NORMALMAP 1u
#define PT_DEVERR_TEXTURE 2u
#define PT_DEVERR_KD_PLANE 4u
#define PT_DEVERR_TIR 8u
#define PT_DEVERR_OVERFLOW 16u

/* ------------------------------------------------------------------------- */
/* Scene records (all little-endian, f64 unless stated)                        */
/* ------------------------------------------------------------------------- */

/* Primitive enum, src/primitive.rs:67-81 */
typedef enum PtPrimType {
    PT_PRIM_SPHERE = 0,   /* src/primitive/sphere.rs */
    PT_PRIM_TRIANGLE = 1, /* src/primitive/triangle.rs as a top-level primitive (mesh record, no bbox gate) */
    PT_PRIM_MESH = 2,     /* src/primitive/mesh.rs: bbox gate + linear scan */
    PT_PRIM_KDMESH = 3,   /* src/kdtree/kdmesh.rs: bbox gate + kd walk */
    PT_PRIM_PLANE = 4,    /* src/primitive/plane.rs */
    PT_PRIM_CUBE = 5,     /* src/primitive/cube.rs */
    PT_PRIM_CYLINDER = 6, /* src/primitive/cylinder.rs */
    PT_PRIM_CONE = 7      /* src/primitive/cone.rs */
} PtPrimType;

/* One kd-tree node, 16 bytes.  Replaces KDTreeNode::{Split,Leaf}
 * (src/kdtree/node.rs:13-25); the per-node BoundingBox is dropped because the
 * traversal (node.rs:66-203) only ever reads the ROOT's extent.
 *   split node: a = axis(0|1|2) | front_child << 2 ; b = back_child ; split = plane coordinate
 *   leaf  node: a = 3 | first_item << 2            ; b = item_count ; split unused
 * child / item indices are relative to the tree's own node / item base. */
typedef struct PtKdNode {
    double split;
    uint32_t a;
    uint32_t b;
} PtKdNode;

/* PartitionConfig + max depth of KDLeaf::partitioned (src/kdtree/leaf.rs:43-67,89).  The reference's values:
 * target_max_nodes 3, target_max_merit 3, max_tries 10 (kdscene.rs:30-34, kdmesh.rs:45-49); max_depth = KD_DEPTH /
 * KD_MESH_DEPTH env or 10 (kdscene.rs:13,36-38; kdmesh.rs:15,51-53). */
typedef struct PtKdBuildConfig {
    uint32_t max_depth;
    uint32_t target_max_nodes;
    int32_t target_max_merit;
    uint32_t max_tries;
} PtKdBuildConfig;

/* What an intersection test needs for one flat instance
 * (FlatSceneNode, src/flat_scene.rs:50-61).  128 bytes, one cache line. */
typedef struct PtInstance {
    double invtrans[12]; /* rows 0..2 of the world->object Mat4, row-major 3x4 */
    uint32_t prim;       /* PtPrimType */
    uint32_t mesh;       /* index into meshes for TRIANGLE/MESH/KDMESH, else 0xFFFFFFFF */
    uint32_t material;   /* index into materials */
    uint32_t reserved;
    double pad[2];
} PtInstance;

/* object->world rows 0..2 (flat_scene.rs:87); normal_trans = invtrans^T is
 * derived from PtInstance.invtrans (flat_scene.rs:105). 96 bytes. */
typedef struct PtInstanceTrans {
    double trans[12];
} PtInstanceTrans;

#define PT_MESH_LINEAR 0u   /* Mesh */
#define PT_MESH_KD 1u       /* KDMesh */
#define PT_MESH_TRIANGLE 2u /* single Triangle primitive, no bbox gate */
#define PT_MESH_FLAG_NORMALS 1u /* Shading::Smooth: per-corner normals present */
#define PT_MESH_FLAG_UVS 2u     /* per-corner texture coordinates present */

/* Per-mesh record. bbox_invtrans is BoundingBox.invtrans
 * (src/bounding_box.rs:57-82), extent is BoundingBox::extent of the kd root
 * (src/bounding_box.rs:95-99, squared diagonal).  160 bytes. */
typedef struct PtMesh {
    double bbox_invtrans[12];
    double extent;
    uint32_t kind;
    uint32_t flags;
    uint32_t tri_first;  /* first triangle in tri_pos (and tri_normals/tri_uvs via *_first below) */
    uint32_t tri_count;
    uint32_t node_first; /* KD: first node in blas_nodes */
    uint32_t node_count;
    uint32_t item_first; /* KD: base of this tree's leaf items in blas_items (values are triangle indices relative to tri_first) */
    uint32_t item_count;
    uint32_t nrm_first;  /* first record in tri_normals or 0xFFFFFFFF */
    uint32_t uv_first;   /* first record in tri_uvs or 0xFFFFFFFF */
    uint32_t kd_depth;
    uint32_t reserved;
    double pad;
} PtMesh;

typedef struct PtTriPos { double a[3], b[3], c[3]; } PtTriPos;       /* 72 B, triangle.rs:10-12 */
typedef struct PtTriNormals { double na[3], nb[3], nc[3]; } PtTriNormals; /* 72 B, triangle.rs:15 */
typedef struct PtTriUvs { double uva[2], uvb[2], uvc[2]; } PtTriUvs; /* 48 B, triangle.rs:18 */

/* src/material.rs:51-86. texture / normals are indices into textures or -1. 160 bytes. */
typedef struct PtMaterial {
    double diffuse[3];
    double specular[3];
    double shininess;
    double reflectivity;
    double glossy_side_length;
    double refraction_index;
    double uv_trans[9]; /* row-major 3x3 */
    int32_t texture;
    int32_t normals;
} PtMaterial;

/* src/light.rs:74-85 (+ Falloff :11-15, Parallelogram :41-46). 128 bytes. */
typedef struct PtLight {
    double position[3];
    double color[3];
    double falloff[3]; /* c0, c1, c2 */
    double area_a[3];
    double area_b[3];
    double pad;
} PtLight;

/* RGB8 row-major image in the texel pool (RgbImageBuffer, src/texture.rs:78-80).  32 bytes.
 * key: 0, or a caller-chosen identity of the image CONTENT (the Rust glue passes the address of the
 * Arc<Texture>'s RgbImageBuffer; the host mirror a hash of the file path + size).  Textures with a
 * non-zero key stay resident in HBM between scenes: a later pt_scene_upload that carries the same
 * (key, width, height) does not copy those texels to the device again.  The caller guarantees that
 * equal keys mean equal texels (textures are immutable once loaded in the reference, texture.rs:96-102). */
typedef struct PtTexture {
    uint32_t width;
    uint32_t height;
    uint64_t offset; /* byte offset into the texel pool */
    uint64_t key;
    uint64_t reserved;
} PtTexture;

/* Pointer form of a prepared scene: what the Rust glue fills from KDTreeScene. */
typedef struct PtSceneDesc {
    double ambient[3];  /* Scene.ambient, src/scene.rs:17 */
    double tlas_extent; /* extent() of the scene kd root, node.rs:29 */
    uint32_t tlas_depth;
    uint32_t n_tlas_nodes;  const PtKdNode* tlas_nodes;
    uint32_t n_tlas_items;  const uint32_t* tlas_items; /* instance indices */
    uint32_t n_instances;   const PtInstance* instances; const PtInstanceTrans* instance_trans;
    uint32_t n_meshes;      const PtMesh* meshes;
    uint32_t n_blas_nodes;  const PtKdNode* blas_nodes;
    uint32_t n_blas_items;  const uint32_t* blas_items;
    uint32_t n_triangles;   const PtTriPos* tri_pos;
    uint32_t n_tri_normals; const PtTriNormals* tri_normals;
    uint32_t n_tri_uvs;     const PtTriUvs* tri_uvs;
    uint32_t n_materials;   const PtMaterial* materials;
    uint32_t n_lights;      const PtLight* lights;
    uint32_t n_textures;    const PtTexture* textures;
    uint64_t n_texel_bytes; const uint8_t* texels;
} PtSceneDesc;

/* Pointer-free form of the same thing: one contiguous blob, every section
 * 128-byte aligned, offsets from the start of the blob.  This is what crosses
 * NVLink verbatim when the scene is broadcast to the other ranks. */
#define PT_BLOB_MAGIC 0x43535450u /* "PTSC" */
#define PT_BLOB_VERSION 3u

typedef struct PtBlobHeader {
    uint32_t magic;
    uint32_t version;
    uint64_t total_bytes;
    double ambient[3];
    double tlas_extent;
    uint32_t tlas_depth;
    uint32_t blas_max_depth;
    uint32_t n_tlas_nodes, n_tlas_items, n_instances, n_meshes;
    uint32_t n_blas_nodes, n_blas_items, n_triangles, n_tri_normals;
    uint32_t n_tri_uvs, n_materials, n_lights, n_textures;
    uint64_t n_texel_bytes;
    uint64_t off_tlas_nodes, off_tlas_items, off_instances, off_instance_trans;
    uint64_t off_meshes, off_blas_nodes, off_blas_items, off_tri_pos;
    uint64_t off_tri_normals, off_tri_uvs, off_materials, off_lights;
    uint64_t off_textures, off_texels;
} PtBlobHeader;

/* Prepared camera, src/camera.rs:17-32 (Camera::new stays on the host). */
typedef struct PtCamera {
    double eye[3];
    double view_to_world[16]; /* row-major Mat4 */
    double fov_factor;
    double aspect_ratio;
    double width;
    double height;
} PtCamera;

/* Deterministic replacement for thread_rng() (render.rs:38, material.rs:106):
 *   FIXED : every draw is 0.5  -> pixel-centre rays, area lights sampled at
 *           their centre, glossy offset exactly 0.
 *   HASH  : draw(pixel, sample, path, dim) =
 *             (mix(mix(seed ^ (pixel * 0x9E3779B97F4A7C15 + sample)) + (path << 8 | dim)) >> 11) * 2^-53
 *           with mix = splitmix64 finaliser
 *             z ^= z >> 30; z *= 0xBF58476D1CE4E5B9; z ^= z >> 27; z *= 0x94D049BB133111EB; z ^= z >> 31.
 *   pixel = y * width + x (global, independent of slices, tiles and ranks),
 *   path  = 1 for the primary ray; child = parent << 1 | (0 reflected, 1 refracted),
 *   dim   = 0,1 pixel jitter (render.rs:39, root only); then per shaded node
 *           2+2l, 3+2l for light l (light.rs:66-67) and 2+2L, 3+2L for the
 *           glossy square (material.rs:235-236). */
#define PT_RNG_FIXED 0u
#define PT_RNG_HASH 1u

#define PT_BG_PER_PIXEL 0u /* background[W*H*3], row-major: background.at(x/w, y/h) at integer x,y (render.rs:31-34) */
#define PT_BG_PER_ROW 1u   /* background[H*3]: the closure depends on v only */
#define PT_BG_CONSTANT 2u  /* background[3] */

typedef struct PtRenderParams {
    uint32_t width, height;   /* full image */
    uint32_t x1, y1, x2, y2;  /* inclusive slice, render.rs:116-119 */
    uint32_t samples;         /* SAMPLES, render.rs:107-113 */
    uint32_t rng_mode;
    uint64_t seed;
    uint32_t bg_mode;
    uint32_t max_depth;       /* 0 -> PT_MAX_RECURSION_DEPTH */
    /* multi-GPU tile ownership: tile k (row-major over the image's tile grid)
     * is rendered iff k % world == rank.  world = 0 or 1 -> everything. */
    uint32_t tile_w, tile_h;  /* 0 -> 32 */
    uint32_t rank, world;
    /* tuning (0 = defaults) */
    uint64_t max_batch_paths;
    uint64_t node_pool_capacity;
    uint32_t flags;           /* PT_RENDER_* */
    uint32_t reserved;
} PtRenderParams;

#define PT_RENDER_COUNTERS 1u  /* run the counting variant of the traversal kernels (fills PtStats work counters) */
#define PT_RENDER_LINEAR_TLAS 2u /* ignore the scene kd-tree: FlatScene linear scan, flat_scene.rs:71-99 + ray.rs:87-99 */
#define PT_RENDER_KERNEL_TIMES 4u /* bracket every extend / shadow / shade launch with CUDA events (fills PtStats.ms_*); uses the stream path */
#define PT_RENDER_ROW_MAJOR 8u    /* pt_frame_*: device outputs are full-image row-major (W*H entries) instead of compact owned-pixel order; world <= 1 only */
#define PT_RENDER_NO_GRAPH 16u    /* launch kernel by kernel on the stream (one host check per recursion level) instead of replaying the frame's CUDA graph */
/* The reference panics ("bug: ray should definitely hit infinite plane", kdtree/node.rs:147,178) when rounding puts a
 * split-plane crossing outside the ray's range; over 10^9 rays that does happen (README.md:247-248 "occasionally buggy").
 * By default the call fails with PT_ERR_KD_PLANE_MISS like the panic would.  With this flag the call succeeds: the
 * offending ray treats that split as a miss of both children (what the device does anyway before it reports), the bit
 * stays in PtStats.device_error_bits and PtStats.err_* say where it happened. */
#define PT_RENDER_TOLERATE_KD_PLANE 32u
/* By default the k-d walks (scene tree and KDMesh trees) do not enter a subtree when the ray certainly misses the union
 * box of everything the subtree's leaves hold, inside the range the subtree would be walked with: nothing down there can
 * return a hit, so pictures, hit ids and hit parameters are exactly those of the reference's full walk.  What a skipped
 * subtree can hide is the reference's kd-plane panic (above), should the reference have tripped it DOWN THERE — on a ray
 * segment that cannot hit anything.  This flag makes the device walk every subtree the reference walks: every such panic
 * is reproduced, at about 1.3x the traversal time.  PT_RENDER_COUNTERS implies it (the counters are the reference's work). */
#define PT_RENDER_EXACT_WALK 64u

typedef struct PtStats {
    /* rays = every ray_cast issued against the scene root */
    uint64_t rays_primary, rays_shadow, rays_reflect, rays_refract;
    uint64_t rays_depth_cut;    /* depth-11 rays the reference casts but whose result is always bg (elided) */
    /* work counters (only with PT_RENDER_COUNTERS) */
    uint64_t kd_splits, instance_tests, triangle_tests, bbox_gates;
    uint64_t shaded_hits, texel_lookups;
    uint64_t nodes_total;       /* ray-tree nodes allocated */
    uint32_t batches, retries;
    uint32_t max_level;         /* deepest recursion level that held a ray */
    uint32_t device_error_bits;
    uint32_t kernel_launches;   /* launches of this library's kernels inside the call */
    uint32_t reserved;
    double device_ms;           /* CUDA-event time of the kernels of the call */
    double h2d_ms, d2h_ms;
    uint64_t h2d_bytes, d2h_bytes;
    /* work counters split by traversal kernel (only with PT_RENDER_COUNTERS): [0] extend, [1] shadow */
    uint64_t k_kd_splits[2], k_instance_tests[2], k_triangle_tests[2], k_bbox_gates[2];
    /* per-kernel device time and launch counts (only with PT_RENDER_KERNEL_TIMES): extend, shadow, shade */
    double ms_extend, ms_shadow, ms_shade;
    uint32_t n_extend, n_shadow, n_shade, reserved2;
    /* sum over instance tests of the primitive's own f64 op count (SURVEY 8d P_type: sphere 30, cube 150, plane 25,
     * cylinder 70, cone 90), per traversal kernel; with the counters above it gives the algorithmic f64 flops */
    uint64_t k_prim_flops[2];
    /* location of the first device-detected error (valid when device_error_bits != 0): the PT_DEVERR_* bit, global
     * pixel index y*W+x, sample, path id (1 = primary; child = parent << 1 | refracted), and
     * kernel (0 extend, 1 shadow, 2 shade) | recursion level << 8 | light << 16 */
    uint32_t err_bit, err_pixel, err_sample, err_pathid, err_where, reserved3;
    /* only with PT_RENDER_KERNEL_TIMES: the extend / shadow time split by recursion level (level 0 = primary rays) */
    double ms_extend_level[16], ms_shadow_level[16];
    /* only with PT_RENDER_COUNTERS: what the device EXECUTED, per traversal kernel ([0] extend, [1] shadow) — FP32 slab
     * tests of the conservative cull, and the exact f64 instance tests / triangle tests / bbox gates that survived it
     * (x_prim_flops: the primitives' own f64 op counts of those instance tests); kd splits are walked exactly as the
     * reference walks them (k_kd_splits).  The k_* counters above are the REFERENCE's work for the same rays. */
    uint64_t x_box_tests[2], x_instance_tests[2], x_triangle_tests[2], x_bbox_gates[2], x_prim_flops[2];
} PtStats;

typedef struct PtScene PtScene; /* opaque, library-owned */
typedef struct PtFrame PtFrame; /* opaque: device buffers of one render target */

/* report_finished_pixels(n), src/reporter.rs:12 — called on the calling thread between batches */
typedef void (*PtProgressFn)(void* user, uint64_t finished_pixels);

/* ---- library ---- */
int pt_init(int device);                 /* cudaSetDevice + stream; device < 0 -> current */
/* A device GROUP inside one process: what rayon's thread pool is to the reference's one `Image::render` call
 * (src/render.rs:127, :216-223 — the call keeps every core of the box busy), this is for the GPUs of a box.
 * ids[0] becomes the primary device.  Afterwards pt_scene_upload also replicates the scene onto the other members
 * (textures device to device) and pt_render fans its tiles over all members — member i renders rank r+w*i of world w*n,
 * (r, w) the caller's own PtRenderParams.rank / world — concurrently, every member's resolve kernel storing its RGB8
 * pixels straight into one image on the primary (peer stores over NVLink) that one D2H copy brings to the host.
 * The picture is bit-identical to a one-device render.  pt_frame_*, pt_kd_*, pt_flatten*, pt_trace_rays run on the
 * primary; pt_scene_set_tlas / pt_scene_set_instances refuse a replicated scene.  n = 1 is pt_init(ids[0]).
 * Calling it (or pt_init) with another configuration shuts the library down first: handles made before are invalid. */
int pt_init_devices(const int* ids, int n);
int pt_device_group_size(void);          /* members of the current group (1 after pt_init; 0 before any init) */
void pt_shutdown(void);
const char* pt_last_error(void);
const char* pt_error_string(int code);   /* the reference's panic text for PT_ERR_* */
int pt_device_count(void);
/* The library keeps device buffers (node pools, frames, resident textures) cached between calls;
 * this returns them to the driver. */
void pt_release_cached_memory(void);
/* bytes of texels currently resident in the texture cache */
uint64_t pt_resident_texture_bytes(void);
/* sizeof() of the boundary's structs as this library was compiled, so a binding can check its mirrors:
 * 0 PtCamera, 1 PtRenderParams, 2 PtStats, 3 PtBlobHeader, 4 PtSceneDesc; anything else -> 0 */
uint64_t pt_abi_sizeof(int which);
/* Measured f64 issue ceiling of this GPU for the library's own instruction mix (separate DMUL + DADD, no FMA):
 * runs a register-resident microbenchmark for about `milliseconds` and returns TFLOP/s in *tflops_out. */
int pt_measure_fp64_rate(double milliseconds, double* tflops_out);

/* ---- texture ingest: what follows the file decode in RgbImageBuffer::open, src/texture.rs:96-102 —
 * image::open(path)?.to_rgb() — on the device.  `pixels` is the decoder's output as it is (host memory, row-major, no
 * padding) in one of the DynamicImage layouts of image 0.21; the library copies it to the device, converts it to RGB8
 * there (to_rgb: gray replicated, alpha dropped, BGR swapped) and keeps the texels resident under `key` (non-zero; the
 * PtTexture.key of the scenes that use it).  A scene blob whose textures are all resident may then be uploaded
 * records-only (pt_scene_upload with bytes = header.off_texels).  In a device group every member gets a copy. */
#define PT_PIXELS_LUMA8 1u
#define PT_PIXELS_LUMAA8 2u
#define PT_PIXELS_RGB8 3u
#define PT_PIXELS_RGBA8 4u
#define PT_PIXELS_BGR8 5u
#define PT_PIXELS_BGRA8 6u
int pt_texture_ingest(const void* pixels, uint32_t width, uint32_t height, uint32_t layout, uint64_t key);
/* the resident texels of `key` back on the host (width * height * 3 bytes), e.g. to hand the same bytes to the oracle */
int pt_texture_read(uint64_t key, uint8_t* rgb_out, uint64_t capacity, uint32_t* width_out, uint32_t* height_out);

/* ---- PNG encode: Image::save, src/render.rs:200-208 (image::ImageBuffer::save -> 8-bit RGB PNG), on
 * the device: filter-0 scanlines in stored deflate blocks, Adler-32 and CRC-32 computed in parallel, IHDR / IDAT / IEND
 * assembled in HBM — the finished FILE crosses PCIe once.  Any PNG decoder returns exactly the rendered pixels; the
 * file's bytes are not the `image` crate's (that deflates with compression). */
uint64_t pt_png_size(uint32_t width, uint32_t height);   /* bytes of the file for an image of this size */
/* device image (row-major RGB8, width * height * 3 bytes) -> device file (pt_png_size bytes); enqueued on `stream`
 * (NULL: the library's stream, and the call waits for it) */
int pt_png_encode_device(const uint8_t* d_rgb, uint32_t width, uint32_t height, uint8_t* d_png_out, void* stream);
/* host image -> host file: H2D, encode, D2H (what Image::save calls when the pixels are in host memory) */
int pt_png_encode(const uint8_t* rgb, uint32_t width, uint32_t height, uint8_t* png_out, uint64_t capacity, uint64_t* png_bytes_out);
/* the picture the frame's last render left in HBM, as a PNG file in host memory (row-major frames: world <= 1) */
int pt_frame_encode_png(PtFrame* frame, uint8_t* png_out, uint64_t capacity, uint64_t* png_bytes_out);

/* ---- scene (replaces nothing in the reference: it is the glue's output) ---- */
/* bytes needed to pack desc; pack it. Pure host code, works without a GPU. */
uint64_t pt_scene_blob_size(const PtSceneDesc* desc);
int pt_scene_pack(const PtSceneDesc* desc, void* blob_out, uint64_t capacity);
/* validate a blob and view it as a desc (pointers into the blob). Host only. */
int pt_scene_unpack(const void* blob, uint64_t bytes, PtSceneDesc* desc_out);
/* the same for a blob cut off at its texel section (bytes >= off_texels; desc_out->texels = NULL): what a caller
 * sends when every texture carries a key and is already resident on the device. Host only. */
int pt_scene_unpack_records(const void* blob, uint64_t bytes, PtSceneDesc* desc_out);

/* host blob -> device.  The record sections are copied verbatim; textures with a key that is already resident are
 * not copied again.  `bytes` may stop at header.off_texels (records-only upload) when every texture is resident,
 * otherwise PT_ERR_INVALID names the missing texture. */
int pt_scene_upload(const void* blob, uint64_t bytes, PtScene** out);
int pt_scene_upload_device(const void* d_blob, uint64_t bytes, PtScene** out); /* blob already in HBM (after the NCCL broadcast); same rules */
void pt_scene_free(PtScene* scene);
/* bytes pt_scene_upload copied host -> device for this scene (records + the textures that were not resident) */
uint64_t pt_scene_uploaded_bytes(const PtScene* scene);

/* ---- render: replaces ImageSliceMut::render's pixel loop, render.rs:127-150 ---- */
/* One blocking call with HOST buffers (the call the Rust shim makes).
 * rgb_inout: W*H*3, only slice pixels owned by (rank, world) are written.
 * hit_id_out (nullable): W*H*2 u32 = (instance, sub id) of the sample-0 primary ray; instance 0xFFFFFFFF = miss.
 * hit_t_out (nullable): W*H f64 ray parameter of that hit (inf on miss). */
int pt_render(PtScene* scene, const PtCamera* camera, const PtRenderParams* params,
              const double* background, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out,
              PtProgressFn progress, void* user, PtStats* stats);

/* Ray::color(scene, background, 0) (src/ray.rs:139-148) for n explicit world-space rays — the
 * per-ray evaluation of the reference's own mesh_equivalence test (src/kdtree/kdmesh.rs:155-163).
 * origins/dirs: n*3; background3: the constant background colour; ray i draws its random numbers
 * as pixel i, sample 0.  Outputs (nullable): color_out n*3 linear f64 (no gamma), hit_id_out n*2,
 * hit_t_out n.  flags: PT_RENDER_COUNTERS. */
int pt_trace_rays(PtScene* scene, uint64_t n, const double* origins, const double* dirs, const double* background3,
                  uint32_t rng_mode, uint64_t seed, uint32_t max_depth, uint32_t flags, double* color_out,
                  uint32_t* hit_id_out, double* hit_t_out, PtStats* stats);

/* The same path with device-resident inputs/outputs, for callers that keep the
 * frame in HBM (multi-GPU gather, benchmarks). */
/* The global pixel indices (y * W + x) a (rank, world) pair renders for these params, in output order.
 * Pure host function (no GPU needed): returns the count; writes at most `capacity` entries when index_out != NULL. */
uint64_t pt_owned_pixels(const PtRenderParams* params, uint32_t* index_out, uint64_t capacity);

int pt_frame_create(PtScene* scene, const PtCamera* camera, const PtRenderParams* params, PtFrame** out);
void pt_frame_free(PtFrame* frame);
/* Point an existing frame (buffers, CUDA graph) at another scene and / or camera of the same image geometry:
 * how a program that renders many scenes at one resolution avoids re-creating frames (pt_render does this
 * internally).  The scene must have the same light count and reflectivity class as the one the frame was sized for
 * (else PT_ERR_INVALID: create a new frame).  seed / rng_mode may change too; pass NULL to keep a value. */
int pt_frame_rebind(PtFrame* frame, PtScene* scene, const PtCamera* camera, const uint64_t* seed, const uint32_t* rng_mode);
uint64_t pt_frame_owned_pixels(const PtFrame* frame);   /* pixels this rank renders */
uint64_t pt_frame_background_doubles(const PtFrame* frame); /* doubles expected in the background buffer */
int pt_frame_set_background(PtFrame* frame, const double* background);          /* host -> device */
int pt_frame_set_background_device(PtFrame* frame, const double* d_background); /* device -> device */
/* render every owned pixel; outputs stay in HBM. stream: cudaStream_t or NULL for the library stream. */
int pt_frame_render(PtFrame* frame, void* stream, PtProgressFn progress, void* user, PtStats* stats);
/* The same render split in two, so that several frames can be in flight on the device at once: enqueue puts the
 * whole frame (state copy, graph replays, control-block read-back) on the stream and returns without waiting;
 * finish waits for it, fills stats and reports device-detected errors.  One render in flight per frame. */
int pt_frame_enqueue(PtFrame* frame, void* stream);
int pt_frame_finish(PtFrame* frame, PtStats* stats);
/* device pointers to the compact outputs, owned-pixel order (see pt_frame_pixel_index) */
const uint8_t* pt_frame_rgb_device(const PtFrame* frame);     /* owned_pixels * 3 */
const uint32_t* pt_frame_hit_id_device(const PtFrame* frame); /* owned_pixels * 2 */
const double* pt_frame_hit_t_device(const PtFrame* frame);    /* owned_pixels */
/* global pixel index (y * W + x) of every owned pixel, in output order (host array of owned_pixels u32) */
int pt_frame_pixel_index(const PtFrame* frame, uint32_t* index_out);
/* copy the compact outputs back and scatter them into full-size host images (any may be NULL) */
int pt_frame_read(PtFrame* frame, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out, PtStats* stats);

/* ---- multi-GPU: the image exchange fused into the resolve kernel ----
 * One process per GPU; every rank owns interleaved tiles of the image (PtRenderParams.rank / world).  The rank that
 * collects the picture allocates a full-image RGB8 buffer with pt_peer_alloc and hands the 64-byte handle to the
 * other ranks (any transport: bench.py broadcasts it with torch.distributed); they map it with pt_peer_open and
 * point their frames at it with pt_frame_set_image_target.  The resolve kernel of every rank then stores its owned
 * pixels at their own place (y * W + x) * 3 of that image — over NVLink for the remote ranks — so no gather and no
 * un-tiling pass follow the render: the only thing left of the exchange is a completion signal (a barrier).
 * The compact owned-pixel outputs (pt_frame_rgb_device) are still written.  Replaces nothing in the reference (it is
 * single-process, src/render.rs:127-150); it is the multi-GPU form of "write RGB8 into the caller's image". */
typedef struct PtPeerHandle { unsigned char bytes[64]; } PtPeerHandle;
int pt_peer_alloc(uint64_t bytes, void** d_ptr, PtPeerHandle* handle_out); /* zero-filled device memory other processes can map */
int pt_peer_free(void* d_ptr);                                               /* pointer from pt_peer_alloc */
int pt_peer_open(const PtPeerHandle* handle, void** d_ptr);                 /* in ANOTHER process than the allocating one */
int pt_peer_close(void* d_ptr);                                              /* pointer from pt_peer_open */
/* d_image_rgb: W * H * 3 bytes, local or peer device memory; NULL detaches.  Takes effect at the next render. */
int pt_frame_set_image_target(PtFrame* frame, uint8_t* d_image_rgb);

/* ---- k-d tree build on the device (the step right before the render loop) ----
 * Replaces KDLeaf::partitioned (src/kdtree/leaf.rs:89-231) as called by KDTreeScene::from (kdscene.rs:19-44: items =
 * flat instances, bounds = FlatSceneNode::bounds) and KDMesh::new (kdmesh.rs:37-58: items = triangles, bounds =
 * Triangle::bounds).  bounds: n x {min x, y, z, max x, y, z} (BoundingBox::min / max of item i), host memory for
 * pt_kd_build, device memory for pt_kd_build_device.  The result is the reference's tree — same planes, same member
 * lists in the same order — serialised breadth-first exactly like the records of PtSceneDesc.tlas_nodes / blas_nodes
 * (front child, then back child; leaf lists in the same order), so it can be copied straight into a scene
 * description, or spliced into an uploaded scene without leaving the device (pt_scene_set_tlas). */
typedef struct PtKdTree PtKdTree; /* opaque: the tree's records in device memory */
int pt_kd_build(const double* bounds, uint32_t n, const PtKdBuildConfig* config, PtKdTree** out);
int pt_kd_build_device(const double* d_bounds, uint32_t n, const PtKdBuildConfig* config, void* stream, PtKdTree** out);
void pt_kd_tree_free(PtKdTree* tree);
uint32_t pt_kd_tree_node_count(const PtKdTree* tree);
uint32_t pt_kd_tree_item_count(const PtKdTree* tree);
uint32_t pt_kd_tree_depth(const PtKdTree* tree);          /* depth of the deepest node, root = 0 */
/* root bounds {min x, y, z, max x, y, z} (KDTreeNode::bounds of the root) and BoundingBox::extent of it
 * (squared diagonal, bounding_box.rs:95-99) — PtSceneDesc.tlas_extent / PtMesh.extent */
int pt_kd_tree_root_bounds(const PtKdTree* tree, double bounds6_out[6], double* extent_out);
/* copy the records to host arrays of pt_kd_tree_node_count / pt_kd_tree_item_count entries (either may be NULL) */
int pt_kd_tree_download(const PtKdTree* tree, PtKdNode* nodes_out, uint32_t* items_out);
/* device time of the build (CUDA events), the number of kernels it launched, and the bytes the build has to move at
 * the least (its HBM roofline numerator: members x passes, see kd_build.cu); any may be NULL */
int pt_kd_tree_build_stats(const PtKdTree* tree, double* device_ms_out, uint32_t* launches_out, uint64_t* algorithmic_bytes_out);
/* Make `tree` the scene tree of an uploaded scene (device-to-device copy; the tree can be freed afterwards).
 * Its items must be flat-instance indices of that scene. */
int pt_scene_set_tlas(PtScene* scene, const PtKdTree* tree);

/* ---- scene flattening on the device (feeds the tree build) ----
 * Replaces FlatScene::from (src/flat_scene.rs:18-46) + FlatSceneNode::new (:101-108) + FlatSceneNode::bounds (:63-69).
 * The glue hands over the scene graph as it is — one PtHierNode per distinct SceneNode (a node shared through
 * Arc<SceneNode> appears once and is referenced from several child lists: instancing), child lists concatenated in
 * `children`, one PtGeometryRec per Geometry — and gets back, in the reference's breadth-first instance order, the
 * PtInstance / PtInstanceTrans records of PtSceneDesc and the world bounds of every instance (the input of
 * pt_kd_build_device), all resident in HBM. */
typedef struct PtHierNode {
    double trans[16];     /* SceneNode::trans(), row-major 4x4 (scene.rs:35,151-205) */
    uint32_t geometry;    /* index into geometries, 0xFFFFFFFF = none (scene.rs:24) */
    uint32_t first_child; /* this node's children are children[first_child .. first_child + child_count) */
    uint32_t child_count;
    uint32_t reserved;
} PtHierNode; /* 144 bytes */
typedef struct PtGeometryRec {
    double bounds[6];     /* primitive.bounds(): min x, y, z, max x, y, z in object space (primitive.rs:63-77) */
    uint32_t prim;        /* PtPrimType */
    uint32_t mesh;        /* as PtInstance.mesh */
    uint32_t material;    /* as PtInstance.material */
    uint32_t reserved;
} PtGeometryRec; /* 64 bytes */
typedef struct PtFlatScene PtFlatScene; /* opaque: the flat instances in device memory */
int pt_flatten(const PtHierNode* nodes, uint32_t n_nodes, const uint32_t* children, uint32_t n_children, uint32_t root,
               const PtGeometryRec* geometries, uint32_t n_geometries, PtFlatScene** out);
void pt_flat_free(PtFlatScene* flat);
uint32_t pt_flat_instance_count(const PtFlatScene* flat);
const double* pt_flat_bounds_device(const PtFlatScene* flat); /* n x 6 doubles, device memory */
/* copy the records to host arrays of pt_flat_instance_count entries (any may be NULL); bounds_out: n x 6 */
int pt_flat_download(const PtFlatScene* flat, PtInstance* instances_out, PtInstanceTrans* trans_out, double* bounds_out);
int pt_flat_build_stats(const PtFlatScene* flat, double* device_ms_out, uint32_t* launches_out);
/* Make `flat` the instances and `tree` (built over pt_flat_bounds_device) the scene tree of an uploaded scene whose
 * blob carried the meshes, materials, lights and textures (device-to-device copies; both can be freed afterwards). */
int pt_scene_set_instances(PtScene* scene, const PtFlatScene* flat, const PtKdTree* tree);

#ifdef __cplusplus
}
#endif
#endif /* PORTRAYER_GPU_H */
