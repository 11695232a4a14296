// C ABI of libportrayer_gpu.so (include/portrayer_gpu.h): scene upload, frame
// management and the batch loop that drives the wavefront kernels.
//
// Runtime design (DESIGN.md "host runtime"):
//   * one library context per process: a stream, a caching device arena (node pools and frame
//     buffers are reused between calls, no cudaMalloc/cudaFree in steady state), a pinned host
//     arena for control blocks and staging, and a texture residency cache keyed by PtTexture.key;
//   * a frame's kernels read everything from a __constant__ FrameState slot, so the frame owns
//     ONE CUDA graph — camera -> WHILE(level has rays){extend, shadow, shade} -> tree_eval ->
//     resolve — that is replayed for any scene / camera after one small async copy;
//   * pt_render enqueues background H2D, the graph(s), the control-block and image D2H on one
//     stream and synchronises once.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "flatten.h"
#include "kd_build.h"
#include "kernels.h"
#include "portrayer_gpu.h"

using namespace ptd;

namespace {

thread_local std::string g_error;
bool g_initialised = false;
std::recursive_mutex g_mu;  // the ABI is blocking and serialised: one render at a time per process
using Lock = std::lock_guard<std::recursive_mutex>;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t e_ = (expr);                                                                               \
        if (e_ != cudaSuccess) return fail(PT_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));        \
    } while (0)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// PT_DEBUG_SYNC=1: synchronise the library stream after every upload-time stage and say which one just finished
// (stderr) — for locating a hanging or faulting kernel on a box without a debugger
cudaStream_t current_stream();
void debug_sync(const char* stage) {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("PT_DEBUG_SYNC");
        on = (e && *e && *e != '0') ? 1 : 0;
    }
    if (!on) return;
    fprintf(stderr, "[pt] %s ...", stage);
    fflush(stderr);
    const cudaError_t e = cudaStreamSynchronize(current_stream());
    fprintf(stderr, " %s\n", cudaGetErrorString(e));
    fflush(stderr);
}

int ensure_init() {
    if (g_initialised) return PT_OK;
    return pt_init(-1);
}

// ------------------------------------------------------------------ caching arenas
// Freed blocks are kept and handed out again to requests of a similar size.  The hot path of a
// renderer allocates the same few shapes over and over (node pool, frame buffers, scene records).
template <bool PINNED>
class Arena {
public:
    void* alloc(size_t bytes, cudaError_t* err) {
        bytes = (std::max<size_t>(bytes, 1) + 511) & ~size_t(511);
        auto it = free_.lower_bound(bytes);
        if (it != free_.end() && it->first <= bytes + bytes / 4 + (size_t(1) << 16)) {
            void* p = it->second;
            live_[p] = it->first;
            cached_ -= it->first;
            free_.erase(it);
            *err = cudaSuccess;
            return p;
        }
        void* p = nullptr;
        *err = raw_alloc(&p, bytes);
        if (*err != cudaSuccess) {  // give cached blocks back to the driver and retry once
            cudaGetLastError();
            trim();
            *err = raw_alloc(&p, bytes);
        }
        if (*err != cudaSuccess) return nullptr;
        live_[p] = bytes;
        return p;
    }
    void release(void* p) {
        if (!p) return;
        auto it = live_.find(p);
        if (it == live_.end()) return;
        free_.emplace(it->second, p);
        cached_ += it->second;
        live_.erase(it);
        while (cached_ > limit_ && !free_.empty()) {  // drop the largest idle blocks first
            auto last = std::prev(free_.end());
            raw_free(last->second);
            cached_ -= last->first;
            free_.erase(last);
        }
    }
    void trim() {
        for (auto& kv : free_) raw_free(kv.second);
        free_.clear();
        cached_ = 0;
    }
    void set_limit(size_t bytes) { limit_ = bytes; }
    size_t cached_bytes() const { return cached_; }

private:
    static cudaError_t raw_alloc(void** p, size_t bytes) { return PINNED ? cudaMallocHost(p, bytes) : cudaMalloc(p, bytes); }
    static void raw_free(void* p) {
        if (PINNED) cudaFreeHost(p);
        else cudaFree(p);
    }
    std::multimap<size_t, void*> free_;
    std::unordered_map<void*, size_t> live_;
    size_t cached_ = 0;
    size_t limit_ = size_t(24) << 30;
};

Arena<true> g_pin;  // pinned host memory: one arena for the process (pinned allocations are usable from every device under UVA)

// ------------------------------------------------------------------ texture residency cache
struct ResidentTexture {
    uint8_t* d_texels = nullptr;
    uint64_t bytes = 0;
    uint32_t width = 0, height = 0;
    uint32_t refs = 0;
    uint64_t last_use = 0;
};

// ------------------------------------------------------------------ per-device state
// Everything the library keeps per GPU.  A process that never calls pt_init_devices has exactly one context (the device
// pt_init chose); a device GROUP (pt_init_devices) has one per member, and pt_scene_upload / pt_render fan out over them.
// The ABI is serialised by g_mu, so ONE "current context" pointer is enough: every entry point that takes a handle
// switches to the handle's context (CtxScope) for the duration of the call, and the g_* names below — the names this
// file used when it knew one device only — resolve through it.
struct DeviceCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    Arena<false> dev;
    std::unordered_map<uint64_t, ResidentTexture> textures;
    uint64_t texture_bytes = 0, texture_tick = 0;
    double* gamma_lut = nullptr;  // 256 doubles, filled once by the device
    uint32_t slots_used = 0;      // bitmap of __constant__ FrameState slots in use
    std::vector<PtFrame*> frame_cache;  // frames pt_render keeps between calls (LRU, small)
    uint8_t* group_image = nullptr;     // primary of a group: the full RGB8 image every member's resolve kernel stores into
    uint64_t group_image_bytes = 0;
    bool peer_to_primary = false;       // this member's kernels can store into the primary's memory
    bool ready = false;
};
constexpr int kMaxDevices = 16;
DeviceCtx g_ctxs[kMaxDevices];
DeviceCtx* g_cur = &g_ctxs[0];
int g_group_size = 1;  // contexts [0, g_group_size) form the device group; [0] is the primary
#define g_stream (g_cur->stream)
#define g_dev (g_cur->dev)
#define g_textures (g_cur->textures)
#define g_texture_bytes (g_cur->texture_bytes)
#define g_texture_tick (g_cur->texture_tick)
#define g_gamma_lut (g_cur->gamma_lut)
#define g_slots_used (g_cur->slots_used)
#define g_frame_cache (g_cur->frame_cache)
cudaStream_t current_stream() { return g_stream; }

// switch the current context (and the CUDA device) for the duration of an API call
struct CtxScope {
    DeviceCtx* prev;
    explicit CtxScope(DeviceCtx* ctx) : prev(g_cur) {
        if (ctx && ctx != g_cur) { g_cur = ctx; cudaSetDevice(ctx->device); }
    }
    ~CtxScope() {
        if (prev != g_cur) { g_cur = prev; if (prev->device >= 0) cudaSetDevice(prev->device); }
    }
    CtxScope(const CtxScope&) = delete;
    CtxScope& operator=(const CtxScope&) = delete;
};

uint64_t g_texture_limit = uint64_t(32) << 30;  // of 180 GB, per device

void evict_textures(uint64_t need) {
    while (g_texture_bytes + need > g_texture_limit) {
        uint64_t best_key = 0, best_tick = ~0ull;
        for (auto& kv : g_textures)
            if (kv.second.refs == 0 && kv.second.last_use < best_tick) { best_tick = kv.second.last_use; best_key = kv.first; }
        if (best_tick == ~0ull) return;
        auto it = g_textures.find(best_key);
        g_dev.release(it->second.d_texels);
        g_texture_bytes -= it->second.bytes;
        g_textures.erase(it);
    }
}

int take_slot() {
    for (int i = 0; i < kStateSlots - 1; ++i)
        if (!(g_slots_used & (1u << i))) { g_slots_used |= 1u << i; return i; }
    return kStateSlots - 1;  // shared overflow slot: state is re-uploaded before every render
}
void give_slot(int slot) {
    if (slot >= 0 && slot < kStateSlots - 1) g_slots_used &= ~(1u << slot);
}

}  // namespace

struct PtScene {
    DeviceCtx* ctx = nullptr;               // the device this copy lives on
    std::vector<PtScene*> replicas;         // device group: the copies on the other members (owned by this handle)
    std::vector<TextureDev> tex_table;      // host copy of d_textures (where each texture's texels lie on this device)
    unsigned char* d_records = nullptr;  // blob sections [0, off_texels) (+ inline texels when not cached)
    uint64_t bytes = 0;
    PtBlobHeader h{};
    DScene view{};
    bool has_reflective = false;
    TextureDev* d_textures = nullptr;
    float4* d_aabb = nullptr;               // padded FP32 world box per instance (conservative cull, traverse.cuh)
    float4* d_tl_cull = nullptr;            // leaf cull structure of the scene tree (leaf_cull.cu): two box sets + the root box (last 2 float4)
    float4* d_bl_cull = nullptr;            // ... of the KDMesh trees
    ptd::LeafCull tl_cull{}, bl_cull{};
    std::vector<ptd::LcTree> blas_trees;    // where each KDMesh tree lies in blas_nodes / blas_items / tri_pos
    float4* d_tri_aabb = nullptr;           // padded FP32 object box per triangle + the Mesh fold levels (one allocation)
    // Mesh fold structure (traverse.cuh mesh_fold): Morton order of every linear mesh's triangles, boxes in that order
    // and 4-ary group levels, all inside d_tri_aabb / d_fold_order
    uint32_t* d_fold_order = nullptr;
    uint32_t fold_off[ptd::kFoldLevels + 1] = {};  // offsets (in boxes) of the fold levels inside d_tri_aabb
    uint32_t fold_levels = 0;
    bool fold_sort = false;  // some linear mesh is big enough to be worth sorting, and the meshes' triangle ranges do not overlap partially
    std::vector<uint64_t> resident_keys;    // textures held in the residency cache (refs to drop)
    std::vector<uint8_t*> private_texels;   // unkeyed textures owned by this scene
    uint64_t h2d_bytes = 0;                 // bytes the upload copied to the device
    PtKdNode* d_own_tlas_nodes = nullptr;   // pt_scene_set_tlas: a device-built scene tree replacing the blob's
    uint32_t* d_own_tlas_items = nullptr;
    PtInstance* d_own_instances = nullptr;  // pt_scene_set_instances: device-flattened instances replacing the blob's
    PtInstanceTrans* d_own_instance_trans = nullptr;
};

struct PtKdTree {
    ptd::KdTreeDev* dev = nullptr;
    DeviceCtx* ctx = nullptr;
};
struct PtFlatScene {
    ptd::FlatSceneDev* dev = nullptr;
    DeviceCtx* ctx = nullptr;
};

struct PtFrame {
    DeviceCtx* ctx = nullptr;
    PtScene* scene = nullptr;
    PtCamera cam{};
    PtRenderParams params{};
    bool row_major = false;
    std::vector<uint32_t> pixel_index;  // owned slot -> global pixel
    uint32_t* d_pixel_index = nullptr;
    double* d_background = nullptr;
    uint64_t bg_doubles = 0;
    uint64_t out_pixels = 0;  // entries of the output buffers (owned pixels, or W*H when row-major)
    uint8_t* d_rgb = nullptr;
    uint8_t* image_target = nullptr;  // pt_frame_set_image_target: full-image RGB8, local or peer memory (not owned)
    uint32_t* d_hit_id = nullptr;
    double* d_hit_t = nullptr;
    // node pool
    unsigned char* d_pool = nullptr;
    NodePool pool{};
    uint32_t n_lights_cap = 0;
    bool reflective_cap = false;
    uint32_t batch_slots = 0;  // owned pixels per batch the frame (graph, node pool) was sized for
    // owned pixels per batch actually used (<= batch_slots): halved, for good, whenever a render of this frame ran out
    // of node pool, so that a scene with deep ray trees pays for the discovery once and not on every render
    uint32_t batch_slots_now = 0;
    BatchCtl* d_ctl = nullptr;
    BatchCtl* h_ctl = nullptr;  // pinned ring, one per batch
    uint32_t h_ctl_count = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::vector<cudaEvent_t> kernel_events;  // PT_RENDER_KERNEL_TIMES: begin/end pairs, reused per batch
    uint32_t max_depth = PT_MAX_RECURSION_DEPTH;
    int n_levels = 1;
    // __constant__ slot + graph
    int slot = -1;
    cudaGraph_t graph[3] = {nullptr, nullptr, nullptr};  // one per walk_mode(): exact, counting, pruned traversal kernels
    cudaGraphExec_t exec[3] = {nullptr, nullptr, nullptr};
    cudaGraphNode_t camera_node[3] = {nullptr, nullptr, nullptr};
    bool graph_failed = false;
    uint64_t last_use = 0;
    // a render that has been enqueued but not finished (pt_frame_enqueue / pt_frame_finish)
    enum { IDLE, IN_FLIGHT, FINISHED } pending = IDLE;
    cudaStream_t pending_stream = nullptr;
    std::vector<std::pair<uint32_t, uint32_t>> pending_batches;  // (first_slot, n_slots) of the batches in flight, in ring order
    int pending_rc = PT_OK;
    std::string pending_error;
    PtProgressFn progress = nullptr;  // of the pt_frame_render call in progress
    void* progress_user = nullptr;
    PtStats pending_stats{};
};

namespace {

// Ray-tree nodes the pool holds per path of a batch when a material reflects (129 B + 1 B / light each: 4.3 GB for the
// default 4 Mi-path batch — HBM is 180 GB).  The reference's scenes average 2-5 nodes per path, but a batch that looks at
// a dielectric needs far more, and every overflow costs that batch a second pass (pt_frame_finish).
constexpr uint64_t kPoolNodesPerPath = 8;
uint64_t g_frame_tick = 0;
constexpr size_t kFrameCacheMax = 6;

const char* panic_text(int code) {
    switch (code) {
        case PT_OK: return "ok";
        case PT_ERR_INVALID: return "invalid argument or malformed scene blob";
        case PT_ERR_CUDA: return "CUDA runtime failure";
        case PT_ERR_NO_TEXCOORD_NORMALMAP: return "Normal/Texture mapping is not supported for this primitive!";
        case PT_ERR_NO_TEXCOORD_TEXTURE: return "Texture mapping is not supported for this primitive!";
        case PT_ERR_KD_PLANE_MISS: return "bug: ray should definitely hit infinite plane";
        case PT_ERR_TIR_INSIDE: return "bug: should not have total internal reflection when casting inside surface";
        case PT_ERR_KD_TOO_DEEP: return "kd-tree deeper than PT_MAX_KD_STACK";
        case PT_ERR_OVERFLOW: return "ray-tree node pool exhausted at the minimum batch size";
        default: return "unknown error";
    }
}

int device_error_to_code(uint32_t bits, uint32_t flags = 0) {
    if (flags & PT_RENDER_TOLERATE_KD_PLANE) bits &= ~PT_DEVERR_KD_PLANE;
    if (bits & PT_DEVERR_NORMALMAP) return PT_ERR_NO_TEXCOORD_NORMALMAP;
    if (bits & PT_DEVERR_TEXTURE) return PT_ERR_NO_TEXCOORD_TEXTURE;
    if (bits & PT_DEVERR_KD_PLANE) return PT_ERR_KD_PLANE_MISS;
    if (bits & PT_DEVERR_TIR) return PT_ERR_TIR_INSIDE;
    return PT_OK;
}

// the reference's panic text + where on the image it fired
int fail_device_error(int code, const PtStats* st, uint32_t width) {
    if (!st || !st->err_bit || !width) return fail(code, "%s", panic_text(code));
    static const char* kKernel[3] = {"extend", "shadow", "shade"};
    return fail(code, "%s [first at pixel (%u, %u) sample %u path %u, %s kernel, recursion level %u]", panic_text(code),
                st->err_pixel % width, st->err_pixel / width, st->err_sample, st->err_pathid, kKernel[(st->err_where & 0xFF) % 3],
                (st->err_where >> 8) & 0xFF);
}

void fill_view(PtScene* s) {
    const PtBlobHeader& h = s->h;
    DScene& v = s->view;
    unsigned char* b = s->d_records;
    v.tlas_nodes = s->d_own_tlas_nodes ? s->d_own_tlas_nodes : reinterpret_cast<const PtKdNode*>(b + h.off_tlas_nodes);
    v.tlas_items = s->d_own_tlas_items ? s->d_own_tlas_items : reinterpret_cast<const uint32_t*>(b + h.off_tlas_items);
    v.instances = s->d_own_instances ? s->d_own_instances : reinterpret_cast<const PtInstance*>(b + h.off_instances);
    v.instance_trans = s->d_own_instance_trans ? s->d_own_instance_trans
                                               : reinterpret_cast<const PtInstanceTrans*>(b + h.off_instance_trans);
    v.meshes = reinterpret_cast<const PtMesh*>(b + h.off_meshes);
    v.blas_nodes = reinterpret_cast<const PtKdNode*>(b + h.off_blas_nodes);
    v.blas_items = reinterpret_cast<const uint32_t*>(b + h.off_blas_items);
    v.tri_pos = reinterpret_cast<const PtTriPos*>(b + h.off_tri_pos);
    v.tri_normals = reinterpret_cast<const PtTriNormals*>(b + h.off_tri_normals);
    v.tri_uvs = reinterpret_cast<const PtTriUvs*>(b + h.off_tri_uvs);
    v.materials = reinterpret_cast<const PtMaterial*>(b + h.off_materials);
    v.lights = reinterpret_cast<const PtLight*>(b + h.off_lights);
    v.textures = s->d_textures;
    v.inst_aabb = s->d_aabb;
    v.tl_cull = s->tl_cull;
    v.bl_cull = s->bl_cull;
    v.tl_root = s->d_tl_cull ? s->d_tl_cull + 2 * ptd::leaf_cull_sizes(s->h.n_tlas_nodes, s->h.n_tlas_items).set_float4 : nullptr;
    v.tri_aabb = s->d_tri_aabb;
    v.fold_order = s->d_fold_order;
    for (int l = 0; l <= ptd::kFoldLevels; ++l)
        v.fold_aabb[l] = s->d_tri_aabb && l <= (int)s->fold_levels ? s->d_tri_aabb + 2 * (size_t)s->fold_off[l] : nullptr;
    v.fold_levels = s->fold_levels;
    v.gamma_lut = g_gamma_lut;
    v.ambient[0] = h.ambient[0]; v.ambient[1] = h.ambient[1]; v.ambient[2] = h.ambient[2];
    v.tlas_extent = h.tlas_extent;
    v.n_lights = h.n_lights;
    v.n_instances = h.n_instances;
    v.n_tlas_nodes = h.n_tlas_nodes;
    v.n_tlas_items = h.n_tlas_items;
}

void free_scene(PtScene* s) {
    if (!s) return;
    for (PtScene* r : s->replicas) free_scene(r);
    s->replicas.clear();
    CtxScope scope(s->ctx);
    for (uint64_t key : s->resident_keys) {
        auto it = g_textures.find(key);
        if (it != g_textures.end() && it->second.refs > 0) --it->second.refs;
    }
    for (uint8_t* p : s->private_texels) g_dev.release(p);
    g_dev.release(s->d_textures);
    g_dev.release(s->d_aabb);
    g_dev.release(s->d_tri_aabb);
    g_dev.release(s->d_fold_order);
    g_dev.release(s->d_tl_cull);
    g_dev.release(s->d_bl_cull);
    g_dev.release(s->d_own_tlas_nodes);
    g_dev.release(s->d_own_tlas_items);
    g_dev.release(s->d_own_instances);
    g_dev.release(s->d_own_instance_trans);
    g_dev.release(s->d_records);
    delete s;
}

// Validate the records on the host; `texels_present` says whether the blob carries its texel section.
int adopt_header(PtScene* s, const void* host_blob, uint64_t bytes, bool* texels_present) {
    if (bytes < sizeof(PtBlobHeader)) return fail(PT_ERR_INVALID, "malformed scene blob");
    PtBlobHeader h;
    memcpy(&h, host_blob, sizeof h);
    if (h.magic != PT_BLOB_MAGIC || h.version != PT_BLOB_VERSION) return fail(PT_ERR_INVALID, "malformed scene blob (magic / version)");
    *texels_present = bytes >= h.total_bytes;
    PtSceneDesc d;
    int rc = *texels_present ? pt_scene_unpack(host_blob, bytes, &d) : pt_scene_unpack_records(host_blob, bytes, &d);
    if (rc != PT_OK) return fail(rc, "%s", rc == PT_ERR_KD_TOO_DEEP ? panic_text(rc) : "malformed scene blob");
    s->has_reflective = false;
    for (uint32_t i = 0; i < d.n_materials; ++i)
        if (d.materials[i].reflectivity > 0.0) s->has_reflective = true;
    // Mesh fold order: sort when a linear mesh is big enough to gain from it; the sort keys assume that two meshes'
    // triangle ranges are either the same range or disjoint (what the packer emits) — anything else keeps index order
    s->fold_sort = false;
    s->blas_trees.clear();
    for (uint32_t i = 0; i < d.n_meshes; ++i) {
        const PtMesh& m = d.meshes[i];
        if (m.kind != PT_MESH_KD || m.node_count == 0) continue;
        bool seen = false;  // several meshes may share one tree (instanced KDMesh): build its boxes once
        for (const ptd::LcTree& t : s->blas_trees) seen = seen || (t.node_first == m.node_first && t.item_first == m.item_first);
        if (!seen) s->blas_trees.push_back(ptd::LcTree{m.node_first, m.node_count, m.item_first, m.tri_first});
    }
    {
        std::vector<std::pair<uint32_t, uint32_t>> ranges;
        for (uint32_t i = 0; i < d.n_meshes; ++i) {
            if (d.meshes[i].kind == PT_MESH_LINEAR && d.meshes[i].tri_count >= 256) s->fold_sort = true;  // below that the sort costs more upload time than the fold gains
            ranges.emplace_back(d.meshes[i].tri_first, d.meshes[i].tri_count);
        }
        std::sort(ranges.begin(), ranges.end());
        for (size_t i = 1; i < ranges.size() && s->fold_sort; ++i) {
            const auto &a = ranges[i - 1], &b = ranges[i];
            if (a != b && (uint64_t)a.first + a.second > b.first) s->fold_sort = false;
        }
    }
    s->h = h;
    if (h.tlas_depth > PT_MAX_KD_STACK || h.blas_max_depth > PT_MAX_KD_STACK)
        return fail(PT_ERR_KD_TOO_DEEP, "kd-tree depth %u / %u exceeds PT_MAX_KD_STACK = %d", h.tlas_depth, h.blas_max_depth,
                    PT_MAX_KD_STACK);
    return PT_OK;
}

// Bind every texture of the scene to device texels: resident copy (by key), or a fresh copy from `texel_src` — host
// or device pointer to the blob's texel section; nullptr = records-only upload — or, per texture, from `srcs[i]`
// (replicating a scene from another device of the group: where the texture lies THERE; `kind` = cudaMemcpyDefault).
int bind_textures(PtScene* s, const PtTexture* tex, const unsigned char* texel_src, cudaMemcpyKind kind, const TextureDev* srcs = nullptr) {
    const uint32_t n = s->h.n_textures;
    if (n == 0) return PT_OK;
    std::vector<TextureDev>& table = s->tex_table;
    table.assign(n, TextureDev{});
    for (uint32_t i = 0; i < n; ++i) {
        const PtTexture& t = tex[i];
        const uint64_t bytes = (uint64_t)t.width * t.height * 3;
        if (t.width == 0 || t.height == 0) return fail(PT_ERR_INVALID, "texture %u is empty", i);
        uint8_t* d = nullptr;
        if (t.key != 0) {
            auto it = g_textures.find(t.key);
            if (it != g_textures.end() && it->second.width == t.width && it->second.height == t.height) {
                d = it->second.d_texels;
                ++it->second.refs;
                it->second.last_use = ++g_texture_tick;
                s->resident_keys.push_back(t.key);
            }
        }
        if (!d) {
            const unsigned char* src = nullptr;
            if (srcs) {
                src = srcs[i].texels;
            } else if (texel_src) {
                if (t.offset > s->h.n_texel_bytes || bytes > s->h.n_texel_bytes - t.offset)
                    return fail(PT_ERR_INVALID, "texture %u lies outside the texel pool", i);
                src = texel_src + t.offset;
            }
            if (!src) return fail(PT_ERR_INVALID, "texture %u (key %llx) is not resident and the blob carries no texels", i,
                                  (unsigned long long)t.key);
            if (t.key != 0) evict_textures(bytes);
            cudaError_t e;
            d = static_cast<uint8_t*>(g_dev.alloc(bytes, &e));
            if (!d) return fail(PT_ERR_CUDA, "texture allocation failed: %s", cudaGetErrorString(e));
            e = cudaMemcpyAsync(d, src, bytes, kind, g_stream);
            if (e != cudaSuccess) { g_dev.release(d); return fail(PT_ERR_CUDA, "texture upload failed: %s", cudaGetErrorString(e)); }
            if (kind == cudaMemcpyHostToDevice) s->h2d_bytes += bytes;
            if (t.key != 0 && g_textures.find(t.key) == g_textures.end()) {
                ResidentTexture r;
                r.d_texels = d; r.bytes = bytes; r.width = t.width; r.height = t.height; r.refs = 1; r.last_use = ++g_texture_tick;
                g_textures.emplace(t.key, r);
                g_texture_bytes += bytes;
                s->resident_keys.push_back(t.key);
            } else {
                s->private_texels.push_back(d);
            }
        }
        table[i] = TextureDev{t.width, t.height, d};
    }
    cudaError_t e;
    s->d_textures = static_cast<TextureDev*>(g_dev.alloc(n * sizeof(TextureDev), &e));
    if (!s->d_textures) return fail(PT_ERR_CUDA, "texture table allocation failed: %s", cudaGetErrorString(e));
    // pageable source: the copy is staged before the call returns
    CUDA_TRY(cudaMemcpyAsync(s->d_textures, table.data(), n * sizeof(TextureDev), cudaMemcpyHostToDevice, g_stream));
    s->h2d_bytes += n * sizeof(TextureDev);
    return PT_OK;
}

ptd::KdAllocator arena_allocator() {
    ptd::KdAllocator al;
    al.alloc = [](size_t bytes, cudaError_t* err) { return g_dev.alloc(bytes, err); };
    al.release = [](void* p) { g_dev.release(p); };
    return al;
}

// leaf cull structure (leaf_cull.cu) of one forest; `extra` float4 are appended to the storage (the scene tree keeps its root box there)
int build_leaf_cull(PtKdNode* d_nodes, uint32_t n_nodes, const uint32_t* d_items, uint32_t n_items, const std::vector<ptd::LcTree>& trees,
                    uint32_t max_depth, const float4* d_item_boxes, size_t extra, float4** d_storage, ptd::LeafCull* out) {
    const ptd::LeafCullSizes sz = ptd::leaf_cull_sizes(n_nodes, n_items);
    cudaError_t e;
    g_dev.release(*d_storage);
    *d_storage = static_cast<float4*>(g_dev.alloc((2 * sz.set_float4 + extra) * sizeof(float4), &e));
    if (!*d_storage) return fail(PT_ERR_CUDA, "leaf cull allocation failed: %s", cudaGetErrorString(e));
    void* scratch = g_dev.alloc(sz.scratch_bytes + std::max<size_t>(trees.size(), 1) * sizeof(ptd::LcTree), &e);
    if (!scratch) return fail(PT_ERR_CUDA, "leaf cull scratch allocation failed: %s", cudaGetErrorString(e));
    ptd::LcTree* d_trees = reinterpret_cast<ptd::LcTree*>(static_cast<unsigned char*>(scratch) + sz.scratch_bytes);
    if (!trees.empty()) {
        // pageable source: staged before the call returns
        e = cudaMemcpyAsync(d_trees, trees.data(), trees.size() * sizeof(ptd::LcTree), cudaMemcpyHostToDevice, g_stream);
        if (e != cudaSuccess) { g_dev.release(scratch); return fail(PT_ERR_CUDA, "leaf cull upload failed: %s", cudaGetErrorString(e)); }
    }
    debug_sync("leaf cull: before launch");
    e = ptd::launch_leaf_cull(d_nodes, n_nodes, d_items, n_items, d_trees, (uint32_t)trees.size(), max_depth + 1, d_item_boxes, *d_storage, scratch,
                              arena_allocator(), out, g_stream);
    debug_sync("leaf cull: built");
    g_dev.release(scratch);  // stream-ordered reuse
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "leaf cull build failed: %s", cudaGetErrorString(e));
    return PT_OK;
}

// the scene tree's cull structure + root box; needs d_aabb and the current tree in the view
int build_tlas_cull(PtScene* s) {
    fill_view(s);
    const std::vector<ptd::LcTree> one{ptd::LcTree{0u, s->h.n_tlas_nodes, 0u, 0u}};
    int rc = build_leaf_cull(const_cast<PtKdNode*>(s->view.tlas_nodes), s->h.n_tlas_nodes, s->view.tlas_items, s->h.n_tlas_items, one,
                             s->h.tlas_depth, s->d_aabb, 2, &s->d_tl_cull, &s->tl_cull);
    if (rc != PT_OK) return rc;
    ptd::launch_root_box(s->d_aabb, s->h.n_instances, s->d_tl_cull + 2 * ptd::leaf_cull_sizes(s->h.n_tlas_nodes, s->h.n_tlas_items).set_float4, g_stream);
    debug_sync("root box");
    fill_view(s);
    return PT_OK;
}

// instance boxes for the FP32 cull, computed on the device from the uploaded records
int build_instance_bounds(PtScene* s, bool with_tlas = true) {
    const uint32_t n = s->h.n_instances;
    cudaError_t e;
    s->d_aabb = static_cast<float4*>(g_dev.alloc(std::max<size_t>(n, 1) * 2 * sizeof(float4), &e));
    if (!s->d_aabb) return fail(PT_ERR_CUDA, "instance bounds allocation failed: %s", cudaGetErrorString(e));
    double* scratch = static_cast<double*>(g_dev.alloc(std::max<size_t>(s->h.n_meshes, 1) * 6 * sizeof(double), &e));
    if (!scratch) return fail(PT_ERR_CUDA, "mesh bounds allocation failed: %s", cudaGetErrorString(e));
    // triangle boxes: index order and the Mesh fold structure (Morton order + 4-ary group levels)
    const uint32_t nt = s->h.n_triangles;
    void* fold_scratch = nullptr;
    size_t fold_temp = 0;
    if (nt) {
        uint32_t off = nt, n_level = nt;
        s->fold_levels = 0;
        for (int l = 0; l <= ptd::kFoldLevels; ++l) {
            s->fold_off[l] = off;
            off += n_level;
            if (l >= 1) s->fold_levels = (uint32_t)l;
            if (n_level <= 1) break;  // a level with one box: nothing coarser to build
            n_level = (n_level + 3u) / 4u;
        }
        s->d_tri_aabb = static_cast<float4*>(g_dev.alloc((size_t)off * 2 * sizeof(float4), &e));
        if (!s->d_tri_aabb) { g_dev.release(scratch); return fail(PT_ERR_CUDA, "triangle bounds allocation failed: %s", cudaGetErrorString(e)); }
        s->d_fold_order = static_cast<uint32_t*>(g_dev.alloc((size_t)nt * sizeof(uint32_t), &e));
        fold_temp = s->fold_sort ? fold_sort_temp_bytes(nt) : 0;
        if (s->d_fold_order) fold_scratch = g_dev.alloc(fold_scratch_bytes(nt, fold_temp), &e);
        if (!fold_scratch) { g_dev.release(scratch); return fail(PT_ERR_CUDA, "fold order allocation failed: %s", cudaGetErrorString(e)); }
    }
    fill_view(s);
    debug_sync("records + textures uploaded");
    launch_instance_bounds(s->view, s->h.n_meshes, scratch, s->d_aabb, g_stream);
    debug_sync("instance bounds");
    if (nt) {
        launch_triangle_bounds(s->view.tri_pos, nt, s->d_tri_aabb, g_stream);
        int end_bit = 33;
        while (end_bit < 64 && (nt >> (end_bit - 32)) != 0) ++end_bit;
        const cudaError_t se = launch_fold_order(s->view.meshes, s->h.n_meshes, s->view.tri_pos, scratch, nt, s->fold_sort, s->d_fold_order,
                                                 fold_scratch, fold_temp, end_bit, g_stream);
        if (se != cudaSuccess) { g_dev.release(scratch); g_dev.release(fold_scratch); return fail(PT_ERR_CUDA, "fold order sort failed: %s", cudaGetErrorString(se)); }
        launch_gather_fold_boxes(s->d_tri_aabb, s->d_fold_order, nt, s->d_tri_aabb + 2 * (size_t)s->fold_off[0], g_stream);
        uint32_t n_level = nt;
        for (uint32_t l = 1; l <= s->fold_levels; ++l) {
            launch_group_bounds(s->d_tri_aabb + 2 * (size_t)s->fold_off[l - 1], n_level, 4u, s->d_tri_aabb + 2 * (size_t)s->fold_off[l], g_stream);
            n_level = (n_level + 3u) / 4u;
        }
        g_dev.release(fold_scratch);
        debug_sync("triangle bounds + fold structure");
    }
    g_dev.release(scratch);  // stream-ordered reuse: later users of the block run after this kernel on g_stream
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "instance bounds kernel failed: %s", cudaGetErrorString(e));
    // leaf cull structures: KDMesh trees (the blob's node records, patched in place), then the scene tree
    int rc = build_leaf_cull(const_cast<PtKdNode*>(s->view.blas_nodes), s->h.n_blas_nodes, s->view.blas_items, s->h.n_blas_items, s->blas_trees,
                             s->h.blas_max_depth, s->d_tri_aabb, 0, &s->d_bl_cull, &s->bl_cull);
    if (rc != PT_OK) return rc;
    if (with_tlas) rc = build_tlas_cull(s);
    else fill_view(s);
    return rc;
}

void destroy_graphs(PtFrame* f) {
    for (int k = 0; k < 3; ++k) {
        if (f->exec[k]) cudaGraphExecDestroy(f->exec[k]);
        if (f->graph[k]) cudaGraphDestroy(f->graph[k]);
        f->exec[k] = nullptr;
        f->graph[k] = nullptr;
    }
}

void free_frame(PtFrame* f) {
    if (!f) return;
    CtxScope scope(f->ctx);
    destroy_graphs(f);
    g_dev.release(f->d_pixel_index);
    g_dev.release(f->d_background);
    g_dev.release(f->d_rgb);
    g_dev.release(f->d_hit_id);
    g_dev.release(f->d_hit_t);
    g_dev.release(f->d_pool);
    g_dev.release(f->d_ctl);
    g_pin.release(f->h_ctl);
    if (f->ev_start) cudaEventDestroy(f->ev_start);
    if (f->ev_stop) cudaEventDestroy(f->ev_stop);
    for (cudaEvent_t e : f->kernel_events) cudaEventDestroy(e);
    give_slot(f->slot);
    delete f;
}

// bytes of a node pool of `capacity` nodes (alloc_pool's layout)
size_t pool_bytes(uint64_t capacity, uint32_t n_lights) {
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    return 12 * up(capacity * sizeof(double)) + 6 * up(capacity * sizeof(uint32_t)) + up(capacity) + up(capacity * std::max<uint32_t>(n_lights, 1));
}
size_t free_device_bytes() {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
    return free_b + g_dev.cached_bytes();
}
// Paths per batch when the caller does not say.  Every recursion level of a batch is one launch of each kernel, and a
// launch ends when its slowest warp does: the more rays a level holds, the smaller the share of that tail.  Measured on
// graphics-castle 3840x2160 x 16 (profiles/r02_batch_size.txt): 1 Mi paths 965 Mrays/s (castle-hd), 4 Mi 1 465, 8 Mi 1 543,
// 16 Mi 1 561, 32 Mi 1 575.  16 Mi paths = a 17 GB node pool when materials reflect (8 nodes per path): a tenth of a
// B200's 180 GB.  Smaller devices (and a B200 that is nearly full, see create_frame) get the 4 Mi batch.
constexpr uint64_t kSmallBatchPaths = 1ull << 22, kLargeBatchPaths = 1ull << 24;
uint64_t default_batch_paths() {
    static uint64_t v = 0;
    if (!v) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
        v = total_b >= (96ull << 30) ? kLargeBatchPaths : kSmallBatchPaths;
        if (const char* e = getenv("PT_BATCH_PATHS")) { const long long n = atoll(e); if (n > 0) v = (uint64_t)n; }
    }
    return v;
}

// carve the SoA node pool out of one allocation
int alloc_pool(unsigned char** d_pool, NodePool* pool, uint32_t capacity, uint32_t n_lights) {
    const size_t cap = capacity;
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    size_t off = 0;
    size_t o_f64[12], o_u32[6], o_mode, o_occl;
    for (int i = 0; i < 12; ++i) { o_f64[i] = off; off = up(off + cap * sizeof(double)); }
    for (int i = 0; i < 6; ++i) { o_u32[i] = off; off = up(off + cap * sizeof(uint32_t)); }
    o_mode = off; off = up(off + cap);
    o_occl = off; off = up(off + cap * std::max<uint32_t>(n_lights, 1));
    cudaError_t e;
    *d_pool = static_cast<unsigned char*>(g_dev.alloc(off, &e));
    if (!*d_pool) return fail(PT_ERR_CUDA, "node pool allocation (%zu bytes) failed: %s", off, cudaGetErrorString(e));
    unsigned char* b = *d_pool;
    double** f64[12] = {&pool->ox, &pool->oy, &pool->oz, &pool->dx, &pool->dy, &pool->dz,
                        &pool->t,  &pool->cr, &pool->cg, &pool->cb, &pool->refl, &pool->fres};
    for (int i = 0; i < 12; ++i) *f64[i] = reinterpret_cast<double*>(b + o_f64[i]);
    uint32_t** u32[6] = {&pool->inst, &pool->sub, &pool->root, &pool->pathid, &pool->child0, &pool->child1};
    for (int i = 0; i < 6; ++i) *u32[i] = reinterpret_cast<uint32_t*>(b + o_u32[i]);
    pool->mode = b + o_mode;
    pool->occl = b + o_occl;
    pool->capacity = capacity;
    return PT_OK;
}

uint32_t effective_max_depth(const PtRenderParams& p) { return p.max_depth ? p.max_depth : PT_MAX_RECURSION_DEPTH; }

FrameParams frame_params(const PtFrame* f) {
    FrameParams fp{};
    fp.cam = f->cam;
    fp.pixel_index = f->d_pixel_index;
    fp.background = f->d_background;
    fp.bg_mode = f->params.bg_mode;
    fp.width = f->params.width;
    fp.height = f->params.height;
    fp.samples = f->params.samples;
    fp.rng_mode = f->params.rng_mode;
    fp.seed = f->params.seed;
    fp.max_depth = f->max_depth;
    return fp;
}

FrameState frame_state(const PtFrame* f) {
    FrameState st{};
    st.sc = f->scene->view;
    st.fp = frame_params(f);
    st.pool = f->pool;
    st.ctl = f->d_ctl;
    st.rgb = f->d_rgb;
    st.rgb_image = f->image_target;
    st.hit_id = f->d_hit_id;
    st.hit_t = f->d_hit_t;
    st.row_major = f->row_major ? 1u : 0u;
    st.n_levels = (uint32_t)f->n_levels;
    return st;
}

// Optional per-kernel timing: one begin/end event pair per extend / shadow / shade launch.
struct KernelTimer {
    std::vector<cudaEvent_t>* events = nullptr;  // null: timing off
    size_t used = 0;
    struct Span { int kind; size_t begin; int level; };
    std::vector<Span> spans;
    int level = 0;  // recursion level of the spans begun from now on
    void begin(int kind, cudaStream_t st) {
        if (!events) return;
        while (events->size() < used + 2) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            events->push_back(e);
        }
        spans.push_back({kind, used, level});
        cudaEventRecord((*events)[used], st);
    }
    void end(cudaStream_t st) {
        if (!events) return;
        cudaEventRecord((*events)[used + 1], st);
        used += 2;
    }
    // after the stream has been synchronised
    void collect(PtStats* stats) {
        if (!events || !stats) return;
        for (const Span& s : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, (*events)[s.begin], (*events)[s.begin + 1]);
            const int lv = std::min(std::max(s.level, 0), 15);
            if (s.kind == 0) { stats->ms_extend += ms; ++stats->n_extend; stats->ms_extend_level[lv] += ms; }
            else if (s.kind == 1) { stats->ms_shadow += ms; ++stats->n_shadow; stats->ms_shadow_level[lv] += ms; }
            else { stats->ms_shade += ms; ++stats->n_shade; }
        }
        spans.clear();
        used = 0;
    }
};

// Stream path: run the levels of one batch kernel by kernel; the host looks at the control block after every
// level and stops at the first level without rays.  Leaves the final control block in *h_ctl.
int run_levels_stream(int slot, uint32_t n_lights, uint64_t capacity, BatchCtl* d_ctl, BatchCtl* h_ctl, uint32_t n_paths,
                      int n_levels, int mode, cudaStream_t st, uint32_t* launches, KernelTimer* timer, bool linear = false) {
    for (int level = 0; level < n_levels; ++level) {
        // level d holds at most n_paths * 2^d rays, and never more than the pool
        const uint64_t bound = (uint64_t)n_paths << std::min(level, 31);
        const uint64_t max_items = std::min<uint64_t>(bound, capacity);
        timer->level = level;
        timer->begin(0, st);
        launch_extend(slot, max_items, mode, linear, st);
        timer->end(st);
        if (n_lights) {
            timer->begin(1, st);
            launch_shadow(slot, max_items, n_lights, mode, linear, st);
            timer->end(st);
        }
        timer->begin(2, st);
        launch_shade(slot, max_items, shade_big_group(n_paths), 0, st);
        timer->end(st);
        *launches += n_lights ? 3 : 2;
        CUDA_TRY(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(BatchCtl), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (h_ctl->error_bits & PT_DEVERR_OVERFLOW) break;
        if (h_ctl->level_start[level + 2] <= h_ctl->level_start[level + 1] && h_ctl->level_hi[level + 1] <= h_ctl->level_hi[level]) break;  // next level is empty
    }
    CUDA_TRY(cudaGetLastError());
    return PT_OK;
}

void accumulate_stats(PtStats* stats, const BatchCtl& c, uint32_t n_paths) {
    if (!stats) return;
    stats->rays_primary += n_paths;
    stats->rays_shadow += c.rays_shadow;
    stats->rays_reflect += c.rays_reflect;
    stats->rays_refract += c.rays_refract;
    stats->rays_depth_cut += c.rays_depth_cut;
    for (int k = 0; k < 2; ++k) {
        stats->k_kd_splits[k] += c.work[k][0];
        stats->k_instance_tests[k] += c.work[k][1];
        stats->k_triangle_tests[k] += c.work[k][2];
        stats->k_bbox_gates[k] += c.work[k][3];
        stats->k_prim_flops[k] += c.work[k][4];
        stats->x_box_tests[k] += c.work[k][5];
        stats->x_instance_tests[k] += c.work[k][6];
        stats->x_triangle_tests[k] += c.work[k][7];
        stats->x_bbox_gates[k] += c.work[k][8];
        stats->x_prim_flops[k] += c.work[k][9];
        stats->kd_splits += c.work[k][0];
        stats->instance_tests += c.work[k][1];
        stats->triangle_tests += c.work[k][2];
        stats->bbox_gates += c.work[k][3];
    }
    stats->shaded_hits += c.shaded_hits;
    stats->texel_lookups += c.texel_lookups;
    stats->nodes_total += (uint64_t)c.pool_count + c.hi_count;
    stats->device_error_bits |= c.error_bits & ~PT_DEVERR_OVERFLOW;
    if (c.err_info[0] && !stats->err_bit) {
        stats->err_bit = c.err_info[0];
        stats->err_pixel = c.err_info[1];
        stats->err_sample = c.err_info[2];
        stats->err_pathid = c.err_info[3];
        stats->err_where = c.err_info[4];
    }
    for (uint32_t d = 0; d + 1 < 16; ++d)
        if ((c.level_start[d + 1] > c.level_start[d] || (d > 0 && c.level_hi[d] > c.level_hi[d - 1])) && d > stats->max_level) stats->max_level = d;
}

bool graphs_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PT_DISABLE_GRAPHS");
        v = (e && *e && *e != '0') ? 0 : 1;
    }
    return v == 1;
}

// which traversal kernels a render with these flags runs (kernels.h)
int walk_mode(uint32_t flags) {
    if (flags & PT_RENDER_COUNTERS) return ptd::kWalkCount;  // the counters are the REFERENCE's work: the full walk
    if (flags & PT_RENDER_EXACT_WALK) return ptd::kWalkExact;
    static int forced = -1;  // PT_EXACT_WALK=1 in the environment: the flag for every render (A/B runs)
    if (forced < 0) { const char* e = getenv("PT_EXACT_WALK"); forced = (e && *e && *e != '0') ? 1 : 0; }
    return forced ? ptd::kWalkExact : ptd::kWalkPrune;
}

int ensure_graph(PtFrame* f, int mode) {
    const int k = mode;
    if (f->exec[k]) return PT_OK;
    cudaError_t e = build_frame_graph(f->slot, f->batch_slots, f->params.samples, std::max<uint32_t>(f->n_lights_cap, 1),
                                      f->pool.capacity, mode, &f->graph[k], &f->exec[k], &f->camera_node[k]);
    if (e != cudaSuccess) {
        cudaGetLastError();
        f->graph_failed = true;
        return fail(PT_ERR_CUDA, "frame graph construction failed: %s", cudaGetErrorString(e));
    }
    return PT_OK;
}

int create_frame(PtScene* scene, const PtCamera* camera, const PtRenderParams* params, PtFrame** out) {
    const PtRenderParams& p = *params;
    if (p.width == 0 || p.height == 0 || p.samples == 0) return fail(PT_ERR_INVALID, "empty image or zero samples");
    if (p.x1 >= p.width || p.x2 >= p.width || p.y1 >= p.height || p.y2 >= p.height)
        return fail(PT_ERR_INVALID, "The positions {x: %u, y: %u} and/or {x: %u, y: %u} are not within an image with width = %u and height = %u",
                    p.x1, p.y1, p.x2, p.y2, p.width, p.height);
    if ((uint64_t)p.width * p.height > 0xFFFFFFF0ull) return fail(PT_ERR_INVALID, "image too large");
    if (p.world > 1 && p.rank >= p.world) return fail(PT_ERR_INVALID, "rank %u out of range for world %u", p.rank, p.world);
    if (p.bg_mode > PT_BG_CONSTANT || p.rng_mode > PT_RNG_HASH) return fail(PT_ERR_INVALID, "bad bg_mode / rng_mode");
    if (effective_max_depth(p) > PT_MAX_DEPTH_SUPPORTED) return fail(PT_ERR_INVALID, "max_depth > %u is not supported", PT_MAX_DEPTH_SUPPORTED);
    if ((p.flags & PT_RENDER_ROW_MAJOR) && p.world > 1)
        return fail(PT_ERR_INVALID, "PT_RENDER_ROW_MAJOR needs world <= 1 (ranks own interleaved tiles)");

    PtFrame* f = new PtFrame();
    f->ctx = scene->ctx;
    f->scene = scene;
    f->cam = *camera;
    f->params = p;
    f->row_major = (p.flags & PT_RENDER_ROW_MAJOR) != 0;
    f->max_depth = effective_max_depth(p);
    f->reflective_cap = scene->has_reflective;
    f->n_lights_cap = scene->h.n_lights;
    f->n_levels = scene->has_reflective ? (int)f->max_depth + 1 : 1;
    f->slot = take_slot();

    // owned pixels: interleaved tiles, 8x4 micro-tiles inside a tile (tiles.c)
    f->pixel_index.resize(pt_owned_pixels(&p, nullptr, 0));
    pt_owned_pixels(&p, f->pixel_index.data(), f->pixel_index.size());
    const uint64_t owned = f->pixel_index.size();
    f->out_pixels = f->row_major ? (uint64_t)p.width * p.height : owned;

    f->bg_doubles = p.bg_mode == PT_BG_PER_PIXEL ? (uint64_t)p.width * p.height * 3
                    : p.bg_mode == PT_BG_PER_ROW ? (uint64_t)p.height * 3
                                                 : 3;
    // batch geometry: whole pixels per batch so a pixel's samples are summed in one place, in order
    uint64_t max_paths = p.max_batch_paths ? p.max_batch_paths : default_batch_paths();
    uint64_t slots = 0, capacity = 0;
    auto size_batch = [&]() -> int {
        slots = std::max<uint64_t>(1, max_paths / p.samples);
        slots = std::min<uint64_t>(slots, std::max<uint64_t>(owned, 1));
        if (slots * p.samples > 0x7FFFFFFFull) slots = 0x7FFFFFFFull / p.samples;
        if (slots == 0) return fail(PT_ERR_INVALID, "samples too large");
        const uint64_t batch_paths = slots * p.samples;
        capacity = p.node_pool_capacity ? p.node_pool_capacity : (scene->has_reflective ? batch_paths * kPoolNodesPerPath : batch_paths);
        capacity = std::max<uint64_t>(capacity, batch_paths);
        // the shade kernel keeps counting refused allocations after the pool is full (pool_count is how the retry sizes its
        // batches), and the graph path runs one more level after the pool has filled: up to 5 x capacity in all, which must
        // not wrap the 32-bit counter
        capacity = std::min<uint64_t>(capacity, 0x30000000ull);
        // the shadow kernel's 32-bit work cursor counts (hit, light) pairs
        capacity = std::min<uint64_t>(capacity, 0xFFF00000ull / std::max<uint32_t>(scene->h.n_lights, 1));
        if (capacity < batch_paths) return fail(PT_ERR_INVALID, "batch too large for %u lights: lower max_batch_paths", scene->h.n_lights);
        return PT_OK;
    };
    {
        int rc = size_batch();
        // the library's own default may be too much for what is free right now: fall back to the small batch
        while (rc == PT_OK && !p.max_batch_paths && max_paths > kSmallBatchPaths && pool_bytes(capacity, scene->h.n_lights) > free_device_bytes() / 2) {
            max_paths /= 2;
            rc = size_batch();
        }
        if (rc != PT_OK) { free_frame(f); return rc; }
    }
    f->batch_slots = (uint32_t)slots;
    f->batch_slots_now = f->batch_slots;
    const uint64_t n_batches = owned ? (owned + slots - 1) / slots : 1;
    f->h_ctl_count = (uint32_t)std::min<uint64_t>(n_batches, 1u << 16);

    cudaError_t e = cudaSuccess;
    auto dev = [&](size_t bytes) -> void* {
        if (e != cudaSuccess) return nullptr;
        return g_dev.alloc(std::max<size_t>(bytes, 16), &e);
    };
    f->d_pixel_index = static_cast<uint32_t*>(dev(owned * sizeof(uint32_t)));
    f->d_background = static_cast<double*>(dev(f->bg_doubles * sizeof(double)));
    f->d_rgb = static_cast<uint8_t*>(dev(f->out_pixels * 3));
    f->d_ctl = static_cast<BatchCtl*>(dev(sizeof(BatchCtl)));
    if (e == cudaSuccess) f->h_ctl = static_cast<BatchCtl*>(g_pin.alloc(sizeof(BatchCtl) * f->h_ctl_count, &e));
    if (e == cudaSuccess) e = cudaEventCreate(&f->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&f->ev_stop);
    if (e == cudaSuccess && owned)
        e = cudaMemcpyAsync(f->d_pixel_index, f->pixel_index.data(), owned * sizeof(uint32_t), cudaMemcpyHostToDevice, g_stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        free_frame(f);
        return fail(PT_ERR_CUDA, "frame allocation failed: %s", cudaGetErrorString(e));
    }
    int rc = alloc_pool(&f->d_pool, &f->pool, (uint32_t)capacity, scene->h.n_lights);
    if (rc != PT_OK) { free_frame(f); return rc; }
    *out = f;
    return PT_OK;
}

// the hit-id / hit-t outputs are optional: allocate them the first time a caller asks
int ensure_id_buffers(PtFrame* f) {
    cudaError_t e = cudaSuccess;
    if (!f->d_hit_id) f->d_hit_id = static_cast<uint32_t*>(g_dev.alloc(std::max<uint64_t>(f->out_pixels, 1) * 2 * sizeof(uint32_t), &e));
    if (e == cudaSuccess && !f->d_hit_t) f->d_hit_t = static_cast<double*>(g_dev.alloc(std::max<uint64_t>(f->out_pixels, 1) * sizeof(double), &e));
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "hit-id buffer allocation failed: %s", cudaGetErrorString(e));
    return PT_OK;
}

// careful mode: one batch at a time, host check after every level, halve the batch on node-pool overflow
int render_stream_path(PtFrame* f, cudaStream_t st, PtProgressFn progress, void* user, PtStats* stats, uint32_t* launches_out,
                       uint32_t* batches_out, uint32_t* retries_out, uint32_t* error_bits_out) {
    const int count = walk_mode(f->params.flags);
    const uint32_t owned = (uint32_t)f->pixel_index.size();
    const uint32_t S = f->params.samples;
    KernelTimer timer;
    if (f->params.flags & PT_RENDER_KERNEL_TIMES) timer.events = &f->kernel_events;
    uint32_t first_slot = 0;
    uint32_t batch_slots = std::max<uint32_t>(1, std::min(f->batch_slots_now, f->batch_slots));
    while (first_slot < owned) {
        const uint32_t n_slots = std::min(batch_slots, owned - first_slot);
        const uint32_t n_paths = n_slots * S;
        launch_camera(f->slot, first_slot, n_slots, S, st);
        *launches_out += 1;
        int rc = run_levels_stream(f->slot, f->scene->h.n_lights, f->pool.capacity, f->d_ctl, f->h_ctl, n_paths, f->n_levels, count,
                                   st, launches_out, &timer, (f->params.flags & PT_RENDER_LINEAR_TLAS) != 0);
        if (rc != PT_OK) return rc;
        timer.collect(stats);
        if (f->h_ctl->error_bits & PT_DEVERR_OVERFLOW) {
            // the ray trees of this batch do not fit: halve the batch and redo it (results do not depend on batching)
            if (n_slots == 1) return fail(PT_ERR_OVERFLOW, "%s", panic_text(PT_ERR_OVERFLOW));
            batch_slots = std::max<uint32_t>(1, n_slots / 2);
            ++*retries_out;
            continue;
        }
        launch_tree_eval(f->slot, n_paths, st);
        launch_resolve(f->slot, n_slots, st);
        *launches_out += 2;
        *error_bits_out |= f->h_ctl->error_bits;
        accumulate_stats(stats, *f->h_ctl, n_paths);
        ++*batches_out;
        first_slot += n_slots;
        if (progress) progress(user, n_slots);  // reporter.report_finished_pixels, render.rs:149
    }
    return PT_OK;
}

// graph path, part 1: every batch is one replay of the frame graph; control blocks come back through a pinned ring.
// Nothing here waits for the device.
int enqueue_ranges(PtFrame* f, cudaStream_t st, const std::vector<std::pair<uint32_t, uint32_t>>& ranges) {
    const int k = walk_mode(f->params.flags);
    int rc = ensure_graph(f, k);
    if (rc != PT_OK) return rc;
    const uint32_t S = f->params.samples;
    const size_t n_batches = ranges.size();
    if (n_batches > f->h_ctl_count) {  // smaller batches than the frame was created with: a longer control-block ring
        if (n_batches > (1u << 16)) return fail(PT_ERR_INVALID, "frame has more batches than control-block slots");
        cudaError_t e = cudaSuccess;
        BatchCtl* ring = static_cast<BatchCtl*>(g_pin.alloc(sizeof(BatchCtl) * n_batches, &e));
        if (!ring) return fail(PT_ERR_CUDA, "control-block ring allocation failed: %s", cudaGetErrorString(e));
        CUDA_TRY(cudaStreamSynchronize(st));  // nothing may still be copying into the old ring
        g_pin.release(f->h_ctl);
        f->h_ctl = ring;
        f->h_ctl_count = (uint32_t)n_batches;
    }
    f->pending_batches = ranges;
    for (size_t b = 0; b < n_batches; ++b) {
        CUDA_TRY(set_graph_batch(f->exec[k], f->camera_node[k], f->slot, ranges[b].first, ranges[b].second, f->batch_slots, S));
        CUDA_TRY(cudaGraphLaunch(f->exec[k], st));
        CUDA_TRY(cudaMemcpyAsync(&f->h_ctl[b], f->d_ctl, sizeof(BatchCtl), cudaMemcpyDeviceToHost, st));
    }
    return PT_OK;
}

// [first, first + n) cut into batches of the frame's current batch size
void split_range(const PtFrame* f, uint32_t first, uint32_t n, std::vector<std::pair<uint32_t, uint32_t>>* out) {
    const uint32_t per_batch = std::max<uint32_t>(1, std::min(f->batch_slots_now, f->batch_slots));
    for (uint32_t done = 0; done < n; done += per_batch) out->emplace_back(first + done, std::min(per_batch, n - done));
}

// A batch ran out of node pool.  pool_count keeps counting the allocations that were refused, so pool_count / capacity
// is a LOWER bound of the shortfall (the refused children's own children never asked): the batch is cut by twice that,
// rounded up to a power of two, and at least halved (a tighter 1.25 x was measured: more retry rounds, slower overall).
uint32_t overflow_divisor(const PtFrame* f, const BatchCtl& c) {
    const double want = 2.0 * ((double)c.pool_count + (double)c.hi_count) / (double)std::max<uint32_t>(f->pool.capacity, 1);
    uint32_t div = 2;
    while ((double)div < want && div < (1u << 16)) div *= 2;
    return div;
}
// ... and the frame keeps the smaller batch for its later renders: a scene with deep ray trees pays for finding that
// out once, not on every render
void shrink_frame_batch(PtFrame* f, uint32_t div) {
    f->batch_slots_now = std::max<uint32_t>(1, std::min(f->batch_slots_now, f->batch_slots) / std::max<uint32_t>(div, 1));
}

int enqueue_graph_path(PtFrame* f, cudaStream_t st) {
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    split_range(f, 0, (uint32_t)f->pixel_index.size(), &ranges);
    return enqueue_ranges(f, st, ranges);
}

// graph path, part 2 (after the stream / event has been waited for): read the control blocks of the batches in flight.
// Batches that ran out of node pool are returned in `failed` (their pixels are not final, nothing of them is counted),
// after the frame's batch size has been shrunk by what they asked for.
struct FailedBatch { uint32_t first, n, div; };
void collect_graph_path(PtFrame* f, PtProgressFn progress, void* user, PtStats* stats, uint32_t* launches_out,
                        uint32_t* batches_out, uint32_t* error_bits_out, std::vector<FailedBatch>* failed, bool whole_frame) {
    failed->clear();
    const uint32_t S = f->params.samples;
    uint32_t max_div = 0;
    for (size_t b = 0; b < f->pending_batches.size(); ++b) {
        const BatchCtl& c = f->h_ctl[b];
        *launches_out += 3 + c.levels_run * 3;
        if (c.error_bits & PT_DEVERR_OVERFLOW) {
            const uint32_t div = overflow_divisor(f, c);
            failed->push_back({f->pending_batches[b].first, f->pending_batches[b].second, div});
            max_div = std::max(max_div, div);
            continue;
        }
        *error_bits_out |= c.error_bits;
        accumulate_stats(stats, c, f->pending_batches[b].second * S);
        ++*batches_out;
        if (progress) progress(user, f->pending_batches[b].second);
    }
    if (whole_frame && max_div) shrink_frame_batch(f, max_div);
}

}  // namespace

// =================================================================== library
extern "C" {

// bring one context up on `device` (no-op when it already is)
static int init_context(DeviceCtx* c, int device) {
    if (c->ready && c->device == device) return PT_OK;
    CUDA_TRY(cudaSetDevice(device));
    c->device = device;
    CtxScope scope(c);
    if (!c->stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (!c->gamma_lut) {
        CUDA_TRY(cudaMalloc(&c->gamma_lut, 256 * sizeof(double)));
        launch_gamma_lut(c->gamma_lut, c->stream);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    // per DEVICE: the persistent grids (occupancy queries) and the kernels' function attributes — the shade kernel's
    // 165 KB tile needs cudaFuncAttributeMaxDynamicSharedMemorySize set on every device that launches it
    CUDA_TRY(cudaSetDevice(device));
    kernels_init();
    CUDA_TRY(cudaGetLastError());
    c->ready = true;
    return PT_OK;
}

static void shutdown_context(DeviceCtx* c) {
    if (c->device < 0) return;
    CtxScope scope(c);
    cudaSetDevice(c->device);
    for (PtFrame* f : c->frame_cache) free_frame(f);
    c->frame_cache.clear();
    for (auto& kv : c->textures) c->dev.release(kv.second.d_texels);  // (scenes still alive keep dangling texel pointers: shutdown is final)
    c->textures.clear();
    c->texture_bytes = 0;
    c->dev.release(c->group_image);
    c->group_image = nullptr;
    c->group_image_bytes = 0;
    c->dev.trim();
    if (c->stream) cudaStreamDestroy(c->stream);
    c->stream = nullptr;
    cudaFree(c->gamma_lut);
    c->gamma_lut = nullptr;
    c->slots_used = 0;
    c->ready = false;
    c->device = -1;
}

static void shutdown_all() {
    for (int i = 0; i < kMaxDevices; ++i) shutdown_context(&g_ctxs[i]);
    g_pin.trim();
    g_cur = &g_ctxs[0];
    g_group_size = 1;
    g_initialised = false;
}

int pt_init(int device) {
    Lock lock(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PT_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) return fail(PT_ERR_INVALID, "device %d out of range (%d devices)", device, count);
    if (g_initialised && (g_group_size != 1 || g_ctxs[0].device != device)) shutdown_all();  // another configuration: start over
    g_cur = &g_ctxs[0];
    int rc = init_context(&g_ctxs[0], device);
    if (rc != PT_OK) return rc;
    CUDA_TRY(cudaSetDevice(device));
    kernels_init();
    g_group_size = 1;
    g_initialised = true;
    return PT_OK;
}

// A device GROUP in one process (the reference's one `Image::render` call keeps every core of the box busy through
// rayon, render.rs:127,216-223; this is that for the GPUs of a box): ids[0] becomes the primary — the device every
// handle-less call and every pt_frame_* / pt_kd_* / pt_flatten call runs on — and from then on
//   * pt_scene_upload puts the scene on the primary from the host blob and REPLICATES records, textures and the
//     upload-time acceleration data to the other members device-to-device (NVLink / NVSwitch when peer access exists);
//   * pt_render splits its tiles over the members (tile k -> member k mod n, the ownership rule of PtRenderParams.rank /
//     world, nested inside the caller's own rank / world), runs all members concurrently, and every member's resolve
//     kernel stores its RGB8 pixels straight into ONE image on the primary (peer stores) that is then copied to the host.
// Results are bit-identical to a one-device render (RNG draws are keyed by the global pixel index).
int pt_init_devices(const int* ids, int n) {
    if (!ids || n < 1 || n > kMaxDevices) return fail(PT_ERR_INVALID, "pt_init_devices: need 1..%d device ids", kMaxDevices);
    Lock lock(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PT_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || ids[i] >= count) return fail(PT_ERR_INVALID, "device %d out of range (%d devices)", ids[i], count);
        for (int j = 0; j < i; ++j)
            if (ids[j] == ids[i]) return fail(PT_ERR_INVALID, "device %d listed twice", ids[i]);
    }
    bool same = g_initialised && g_group_size == n;
    for (int i = 0; same && i < n; ++i) same = g_ctxs[i].ready && g_ctxs[i].device == ids[i];
    if (same) return PT_OK;
    if (g_initialised) shutdown_all();
    g_cur = &g_ctxs[0];
    for (int i = 0; i < n; ++i) {
        int rc = init_context(&g_ctxs[i], ids[i]);
        if (rc != PT_OK) { shutdown_all(); return rc; }
    }
    // peer access both ways between the primary and every other member (scene replication reads the primary's memory,
    // resolve kernels store into it); without it copies are staged by the driver and the image is gathered on the host
    for (int i = 1; i < n; ++i) {
        int can_a = 0, can_b = 0;
        cudaDeviceCanAccessPeer(&can_a, ids[i], ids[0]);
        cudaDeviceCanAccessPeer(&can_b, ids[0], ids[i]);
        g_ctxs[i].peer_to_primary = false;
        if (can_a) {
            cudaSetDevice(ids[i]);
            e = cudaDeviceEnablePeerAccess(ids[0], 0);
            g_ctxs[i].peer_to_primary = (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled);
        }
        if (can_b) {
            cudaSetDevice(ids[0]);
            cudaDeviceEnablePeerAccess(ids[i], 0);
        }
        cudaGetLastError();
    }
    g_ctxs[0].peer_to_primary = true;
    CUDA_TRY(cudaSetDevice(ids[0]));
    kernels_init();
    g_group_size = n;
    g_initialised = true;
    return PT_OK;
}

int pt_device_group_size(void) {
    Lock lock(g_mu);
    return g_initialised ? g_group_size : 0;
}

void pt_release_cached_memory(void) {
    Lock lock(g_mu);
    for (int i = 0; i < kMaxDevices; ++i) {
        DeviceCtx* c = &g_ctxs[i];
        if (!c->ready) continue;
        CtxScope scope(c);
        for (PtFrame* f : g_frame_cache) free_frame(f);
        g_frame_cache.clear();
        for (auto it = g_textures.begin(); it != g_textures.end();) {
            if (it->second.refs == 0) {
                g_dev.release(it->second.d_texels);
                g_texture_bytes -= it->second.bytes;
                it = g_textures.erase(it);
            } else {
                ++it;
            }
        }
        g_dev.release(c->group_image);
        c->group_image = nullptr;
        c->group_image_bytes = 0;
        g_dev.trim();
    }
    g_pin.trim();
}

int pt_measure_fp64_rate(double milliseconds, double* tflops_out) {
    if (!tflops_out) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    cudaError_t e;
    double* sink = static_cast<double*>(g_dev.alloc(64, &e));
    if (!sink) return fail(PT_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    int iters = 2048;
    double best = 0.0;
    launch_fp64_rate(sink, 256, g_stream);  // warm-up
    for (int rep = 0; rep < 6; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, g_stream));
        const double flops = launch_fp64_rate(sink, iters, g_stream);
        CUDA_TRY(cudaEventRecord(e1, g_stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        if (ms < milliseconds * 0.5 && iters < (1 << 24)) iters *= 2;  // grow until one launch lasts about the requested time
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    g_dev.release(sink);
    CUDA_TRY(cudaGetLastError());
    *tflops_out = best;
    return PT_OK;
}

uint64_t pt_abi_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(PtCamera);
        case 1: return sizeof(PtRenderParams);
        case 2: return sizeof(PtStats);
        case 3: return sizeof(PtBlobHeader);
        case 4: return sizeof(PtSceneDesc);
        default: return 0;
    }
}

uint64_t pt_resident_texture_bytes(void) {
    Lock lock(g_mu);
    return g_texture_bytes;
}

void pt_shutdown(void) {
    Lock lock(g_mu);
    shutdown_all();
}

const char* pt_last_error(void) { return g_error.c_str(); }
const char* pt_error_string(int code) { return panic_text(code); }

int pt_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

// =================================================================== scene
// the scene on the CURRENT context from a host blob.  `texel_from` (replication inside a device group): the copy of the
// same scene on the primary — its textures are copied device to device instead of crossing PCIe once per member.
// `sync`: wait for the upload before returning (the caller may then free / reuse `blob`).
static int upload_scene_host(const void* blob, uint64_t bytes, const PtScene* texel_from, bool sync, PtScene** out) {
    PtScene* s = new PtScene();
    s->ctx = g_cur;
    bool texels_present = false;
    int rc = PT_OK;
    if (texel_from) {  // validated once, on the primary
        s->h = texel_from->h;
        s->has_reflective = texel_from->has_reflective;
        s->fold_sort = texel_from->fold_sort;
        s->blas_trees = texel_from->blas_trees;
    } else {
        rc = adopt_header(s, blob, bytes, &texels_present);
        if (rc != PT_OK) { delete s; return rc; }
    }
    // records: everything before the texel section, verbatim
    s->bytes = std::min<uint64_t>(s->h.off_texels, s->h.total_bytes);
    cudaError_t e;
    s->d_records = static_cast<unsigned char*>(g_dev.alloc(s->bytes, &e));
    if (s->d_records) e = cudaMemcpyAsync(s->d_records, blob, s->bytes, cudaMemcpyHostToDevice, g_stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        free_scene(s);
        return fail(PT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    s->h2d_bytes = s->bytes;
    const unsigned char* base = static_cast<const unsigned char*>(blob);
    const PtTexture* tex = reinterpret_cast<const PtTexture*>(base + s->h.off_textures);
    if (texel_from) rc = bind_textures(s, tex, nullptr, cudaMemcpyDefault, texel_from->tex_table.data());
    else rc = bind_textures(s, tex, texels_present ? base + s->h.off_texels : nullptr, cudaMemcpyHostToDevice);
    if (rc == PT_OK) rc = build_instance_bounds(s);
    if (rc == PT_OK && sync) {
        e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    if (rc != PT_OK) { free_scene(s); return rc; }
    fill_view(s);
    *out = s;
    return PT_OK;
}

int pt_scene_upload(const void* blob, uint64_t bytes, PtScene** out) {
    if (!blob || !out) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    CtxScope scope(&g_ctxs[0]);
    PtScene* s = nullptr;
    rc = upload_scene_host(blob, bytes, nullptr, true, &s);
    if (rc != PT_OK) return rc;
    // device group: the other members' copies; every member builds its acceleration data on its own stream, at once
    for (int i = 1; i < g_group_size && rc == PT_OK; ++i) {
        CtxScope member(&g_ctxs[i]);
        PtScene* r = nullptr;
        rc = upload_scene_host(blob, bytes, s, false, &r);
        if (rc == PT_OK) s->replicas.push_back(r);
    }
    for (PtScene* r : s->replicas) {
        CtxScope member(r->ctx);
        const cudaError_t e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess && rc == PT_OK) rc = fail(PT_ERR_CUDA, "scene replication failed: %s", cudaGetErrorString(e));
        s->h2d_bytes += r->h2d_bytes;
    }
    if (rc != PT_OK) { free_scene(s); return rc; }
    *out = s;
    return PT_OK;
}

int pt_scene_upload_device(const void* d_blob, uint64_t bytes, PtScene** out) {
    if (!d_blob || !out || bytes < sizeof(PtBlobHeader)) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    // The records are validated on a host copy (a few MB at most for the reference's scenes; the texel pool is skipped).
    PtBlobHeader h;
    CUDA_TRY(cudaMemcpy(&h, d_blob, sizeof h, cudaMemcpyDeviceToHost));
    if (h.magic != PT_BLOB_MAGIC || h.version != PT_BLOB_VERSION || h.off_texels > h.total_bytes || bytes < h.off_texels)
        return fail(PT_ERR_INVALID, "malformed scene blob");
    const bool texels_present = bytes >= h.total_bytes;
    std::vector<unsigned char> host(texels_present ? h.total_bytes : h.off_texels);
    CUDA_TRY(cudaMemcpy(host.data(), d_blob, h.off_texels, cudaMemcpyDeviceToHost));
    PtScene* s = new PtScene();
    s->ctx = g_cur;
    bool present2 = false;
    rc = adopt_header(s, host.data(), host.size(), &present2);
    if (rc != PT_OK) { delete s; return rc; }
    s->bytes = h.off_texels;
    cudaError_t e;
    s->d_records = static_cast<unsigned char*>(g_dev.alloc(s->bytes, &e));
    if (s->d_records) e = cudaMemcpyAsync(s->d_records, d_blob, s->bytes, cudaMemcpyDeviceToDevice, g_stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        free_scene(s);
        return fail(PT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    rc = bind_textures(s, reinterpret_cast<const PtTexture*>(host.data() + h.off_textures),
                       texels_present ? static_cast<const unsigned char*>(d_blob) + h.off_texels : nullptr,
                       cudaMemcpyDeviceToDevice);
    if (rc == PT_OK) rc = build_instance_bounds(s);
    if (rc == PT_OK) {
        e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    if (rc != PT_OK) { free_scene(s); return rc; }
    fill_view(s);
    *out = s;
    return PT_OK;
}

uint64_t pt_scene_uploaded_bytes(const PtScene* scene) { return scene ? scene->h2d_bytes : 0; }

void pt_scene_free(PtScene* scene) {
    if (!scene) return;
    Lock lock(g_mu);
    auto forget = [](PtScene* sc) {
        for (PtFrame* f : sc->ctx->frame_cache)
            if (f->scene == sc) f->scene = nullptr;
    };
    forget(scene);
    for (PtScene* r : scene->replicas) forget(r);
    free_scene(scene);
}

// =================================================================== frames
int pt_frame_create(PtScene* scene, const PtCamera* camera, const PtRenderParams* params, PtFrame** out) {
    if (!scene || !camera || !params || !out) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    CtxScope scope(scene->ctx);
    PtFrame* f = nullptr;
    int rc = create_frame(scene, camera, params, &f);
    if (rc != PT_OK) return rc;
    rc = ensure_id_buffers(f);  // explicit frames always expose hit ids
    if (rc != PT_OK) { free_frame(f); return rc; }
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    *out = f;
    return PT_OK;
}

int pt_frame_rebind(PtFrame* frame, PtScene* scene, const PtCamera* camera, const uint64_t* seed, const uint32_t* rng_mode) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    if (frame->pending == PtFrame::IN_FLIGHT) return fail(PT_ERR_INVALID, "the frame has a render in flight");
    if (scene) {
        if (scene->ctx != frame->ctx) return fail(PT_ERR_INVALID, "the scene lives on another device than the frame");
        if (scene->h.n_lights != frame->n_lights_cap || scene->has_reflective != frame->reflective_cap)
            return fail(PT_ERR_INVALID, "scene has %u lights / reflective=%d, the frame was sized for %u / %d", scene->h.n_lights,
                        (int)scene->has_reflective, frame->n_lights_cap, (int)frame->reflective_cap);
        frame->scene = scene;
    }
    if (camera) frame->cam = *camera;
    if (seed) frame->params.seed = *seed;
    if (rng_mode) {
        if (*rng_mode > PT_RNG_HASH) return fail(PT_ERR_INVALID, "bad rng_mode");
        frame->params.rng_mode = *rng_mode;
    }
    return PT_OK;
}

void pt_frame_free(PtFrame* frame) {
    Lock lock(g_mu);
    free_frame(frame);  // switches to the frame's device itself
}
uint64_t pt_frame_owned_pixels(const PtFrame* frame) { return frame ? frame->pixel_index.size() : 0; }
uint64_t pt_frame_background_doubles(const PtFrame* frame) { return frame ? frame->bg_doubles : 0; }

int pt_frame_set_background(PtFrame* frame, const double* background) {
    if (!frame || !background) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    CUDA_TRY(cudaMemcpyAsync(frame->d_background, background, frame->bg_doubles * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PT_OK;
}
int pt_frame_set_background_device(PtFrame* frame, const double* d_background) {
    if (!frame || !d_background) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    CUDA_TRY(cudaMemcpyAsync(frame->d_background, d_background, frame->bg_doubles * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PT_OK;
}

// the blocking careful render (stream path), also the fallback of the graph path
static int render_blocking_stream(PtFrame* f, cudaStream_t st, PtProgressFn progress, void* user, PtStats* stats, uint32_t retries0) {
    uint32_t launches = 0, batches = 0, retries = retries0, error_bits = 0;
    int rc = render_stream_path(f, st, progress, user, stats, &launches, &batches, &retries, &error_bits);
    if (rc != PT_OK) return rc;
    CUDA_TRY(cudaEventRecord(f->ev_stop, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, f->ev_start, f->ev_stop));
    if (stats) {
        stats->device_ms = ms;
        stats->batches = batches;
        stats->retries = retries;
        stats->kernel_launches = launches;
        stats->device_error_bits = error_bits & ~PT_DEVERR_OVERFLOW;
    }
    const int code = device_error_to_code(error_bits, f->params.flags);
    if (code != PT_OK) return fail_device_error(code, stats, f->params.width);
    return PT_OK;
}

int pt_frame_enqueue(PtFrame* frame, void* stream) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    PtFrame* f = frame;
    if (!f->scene) return fail(PT_ERR_INVALID, "the frame's scene has been freed");
    if (f->pending == PtFrame::IN_FLIGHT) return fail(PT_ERR_INVALID, "the frame already has a render in flight: call pt_frame_finish first");
    cudaStream_t st = stream ? (cudaStream_t)stream : g_stream;
    f->pending_stream = st;
    f->pending_stats = PtStats{};
    f->pending_rc = PT_OK;
    const uint32_t owned = (uint32_t)f->pixel_index.size();
    const FrameState state = frame_state(f);
    CUDA_TRY(cudaEventRecord(f->ev_start, st));
    CUDA_TRY(upload_state(f->slot, state, st));
    const bool want_stream = (f->params.flags & (PT_RENDER_KERNEL_TIMES | PT_RENDER_NO_GRAPH | PT_RENDER_LINEAR_TLAS)) != 0 || !graphs_enabled() ||
                             f->graph_failed;
    if (owned != 0 && !want_stream) {
        int rc = enqueue_graph_path(f, st);
        if (rc == PT_OK) {
            CUDA_TRY(cudaEventRecord(f->ev_stop, st));
            f->pending = PtFrame::IN_FLIGHT;
            return PT_OK;
        }
        if (!f->graph_failed) return rc;  // a real error; otherwise: no graph support on this driver, use the stream path
    }
    // stream path (per-kernel timing, PT_RENDER_NO_GRAPH, fallback) or an empty frame: done synchronously here
    if (owned == 0) {
        CUDA_TRY(cudaEventRecord(f->ev_stop, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        f->pending_rc = render_blocking_stream(f, st, f->progress, f->progress_user, &f->pending_stats, 0);
        if (f->pending_rc != PT_OK) f->pending_error = g_error;
    }
    f->pending = PtFrame::FINISHED;
    return PT_OK;
}

int pt_frame_finish(PtFrame* frame, PtStats* stats) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    PtFrame* f = frame;
    if (f->pending == PtFrame::IDLE) return fail(PT_ERR_INVALID, "no render was enqueued on this frame");
    const double h2d_ms = stats ? stats->h2d_ms : 0.0, d2h_ms = stats ? stats->d2h_ms : 0.0;
    const uint64_t h2d_b = stats ? stats->h2d_bytes : 0, d2h_b = stats ? stats->d2h_bytes : 0;
    int rc = PT_OK;
    if (f->pending == PtFrame::IN_FLIGHT) {
        f->pending = PtFrame::IDLE;
        CUDA_TRY(cudaEventSynchronize(f->ev_stop));
        CUDA_TRY(cudaGetLastError());
        PtStats local{};
        uint32_t launches = 0, batches = 0, error_bits = 0, retries = 0;
        std::vector<FailedBatch> failed;
        std::vector<std::pair<uint32_t, uint32_t>> again;
        collect_graph_path(f, f->progress, f->progress_user, &local, &launches, &batches, &error_bits, &failed, true);
        // Node pool overflow: the ray trees of some batches did not fit.  Results do not depend on batching and batches
        // are independent, so ONLY those batches are redone, each cut by what it asked for (overflow_divisor); when most
        // of the frame overflowed, the frame also keeps a smaller batch for its later renders (collect_graph_path).  The
        // careful path (host check after every level) remains as the last resort.
        while (rc == PT_OK && !failed.empty() && retries < 8) {
            ++retries;
            again.clear();
            bool stuck = false;
            for (const FailedBatch& r : failed) {
                if (r.n <= 1) { stuck = true; break; }  // a single pixel whose ray trees do not fit
                const uint32_t piece = std::max<uint32_t>(1, r.n / r.div);
                for (uint32_t done = 0; done < r.n; done += piece) again.emplace_back(r.first + done, std::min(piece, r.n - done));
            }
            if (stuck) break;
            rc = enqueue_ranges(f, f->pending_stream, again);
            if (rc != PT_OK) break;
            CUDA_TRY(cudaEventRecord(f->ev_stop, f->pending_stream));
            CUDA_TRY(cudaEventSynchronize(f->ev_stop));
            CUDA_TRY(cudaGetLastError());
            collect_graph_path(f, f->progress, f->progress_user, &local, &launches, &batches, &error_bits, &failed, false);
        }
        if (rc != PT_OK) {
            // enqueue failed above: reported below
        } else if (!failed.empty()) {
            local = PtStats{};
            CUDA_TRY(cudaEventRecord(f->ev_start, f->pending_stream));
            rc = render_blocking_stream(f, f->pending_stream, f->progress, f->progress_user, &local, 1 + retries);
        } else {
            local.retries = retries;
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, f->ev_start, f->ev_stop));
            local.device_ms = ms;
            local.batches = batches;
            local.kernel_launches = launches;
            local.device_error_bits = error_bits & ~PT_DEVERR_OVERFLOW;
            const int code = device_error_to_code(error_bits, f->params.flags);
            if (code != PT_OK) rc = fail_device_error(code, &local, f->params.width);
        }
        if (stats) *stats = local;
    } else {
        f->pending = PtFrame::IDLE;
        rc = f->pending_rc;
        if (rc != PT_OK) g_error = f->pending_error;
        if (stats) *stats = f->pending_stats;
    }
    if (stats) { stats->h2d_ms = h2d_ms; stats->d2h_ms = d2h_ms; stats->h2d_bytes = h2d_b; stats->d2h_bytes = d2h_b; }
    return rc;
}

int pt_frame_render(PtFrame* frame, void* stream, PtProgressFn progress, void* user, PtStats* stats) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    frame->progress = progress;
    frame->progress_user = user;
    int rc = pt_frame_enqueue(frame, stream);
    if (rc == PT_OK) rc = pt_frame_finish(frame, stats);
    frame->progress = nullptr;
    frame->progress_user = nullptr;
    return rc;
}

const uint8_t* pt_frame_rgb_device(const PtFrame* frame) { return frame ? frame->d_rgb : nullptr; }
const uint32_t* pt_frame_hit_id_device(const PtFrame* frame) { return frame ? frame->d_hit_id : nullptr; }
const double* pt_frame_hit_t_device(const PtFrame* frame) { return frame ? frame->d_hit_t : nullptr; }

int pt_frame_set_image_target(PtFrame* frame, uint8_t* d_image_rgb) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    Lock lock(g_mu);
    if (frame->pending == PtFrame::IN_FLIGHT) return fail(PT_ERR_INVALID, "a render of this frame is in flight");
    frame->image_target = d_image_rgb;  // read by frame_state() when the next render uploads the frame's constants
    return PT_OK;
}

int pt_peer_alloc(uint64_t bytes, void** d_ptr, PtPeerHandle* handle_out) {
    if (!d_ptr || !handle_out || bytes == 0) return fail(PT_ERR_INVALID, "null argument or zero size");
    static_assert(sizeof(PtPeerHandle) == sizeof(cudaIpcMemHandle_t), "PtPeerHandle carries a cudaIpcMemHandle_t");
    Lock lock(g_mu);
    if (int rc = ensure_init()) return rc;
    void* p = nullptr;
    // plain cudaMalloc, not the caching arena: an IPC handle names a whole allocation
    CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(PT_ERR_CUDA, "pt_peer_alloc: %s", cudaGetErrorString(e)); }
    memcpy(handle_out->bytes, &h, sizeof(h));
    *d_ptr = p;
    return PT_OK;
}

int pt_peer_free(void* d_ptr) {
    if (!d_ptr) return PT_OK;
    Lock lock(g_mu);
    CUDA_TRY(cudaFree(d_ptr));
    return PT_OK;
}

int pt_peer_open(const PtPeerHandle* handle, void** d_ptr) {
    if (!handle || !d_ptr) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    if (int rc = ensure_init()) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof(h));
    void* p = nullptr;
    // enables peer access between this device and the owner's when they differ (NVLink / NVSwitch on an HGX board)
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "pt_peer_open: %s", cudaGetErrorString(e));
    *d_ptr = p;
    return PT_OK;
}

int pt_peer_close(void* d_ptr) {
    if (!d_ptr) return PT_OK;
    Lock lock(g_mu);
    CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return PT_OK;
}

int pt_frame_pixel_index(const PtFrame* frame, uint32_t* index_out) {
    if (!frame || !index_out) return fail(PT_ERR_INVALID, "null argument");
    memcpy(index_out, frame->pixel_index.data(), frame->pixel_index.size() * sizeof(uint32_t));
    return PT_OK;
}

int pt_frame_read(PtFrame* frame, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out, PtStats* stats) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    const size_t owned = frame->pixel_index.size();
    const double t0 = now_ms();
    uint64_t bytes = 0;
    if ((hit_id_out && !frame->d_hit_id) || (hit_t_out && !frame->d_hit_t))
        return fail(PT_ERR_INVALID, "hit ids were not rendered for this frame");
    if (owned == 0) return PT_OK;
    if (frame->row_major) {
        // the slice rectangle goes straight into the caller's image: only slice pixels are written (render.rs:136-138)
        const PtRenderParams& p = frame->params;
        const size_t W = p.width, w = p.x2 - p.x1 + 1, hgt = p.y2 - p.y1 + 1;
        const size_t first = (size_t)p.y1 * W + p.x1;
        if (rgb_inout) { CUDA_TRY(cudaMemcpy2DAsync(rgb_inout + first * 3, W * 3, frame->d_rgb + first * 3, W * 3, w * 3, hgt, cudaMemcpyDeviceToHost, g_stream)); bytes += w * hgt * 3; }
        if (hit_id_out) { CUDA_TRY(cudaMemcpy2DAsync(hit_id_out + first * 2, W * 8, frame->d_hit_id + first * 2, W * 8, w * 8, hgt, cudaMemcpyDeviceToHost, g_stream)); bytes += w * hgt * 8; }
        if (hit_t_out) { CUDA_TRY(cudaMemcpy2DAsync(hit_t_out + first, W * 8, frame->d_hit_t + first, W * 8, w * 8, hgt, cudaMemcpyDeviceToHost, g_stream)); bytes += w * hgt * 8; }
        CUDA_TRY(cudaStreamSynchronize(g_stream));
    } else {
        // compact outputs -> pinned staging -> scatter into the caller's images
        cudaError_t e = cudaSuccess;
        uint8_t* rgb = nullptr;
        uint32_t* ids = nullptr;
        double* ts = nullptr;
        if (rgb_inout) { rgb = static_cast<uint8_t*>(g_pin.alloc(owned * 3, &e)); if (rgb) e = cudaMemcpyAsync(rgb, frame->d_rgb, owned * 3, cudaMemcpyDeviceToHost, g_stream); bytes += owned * 3; }
        if (e == cudaSuccess && hit_id_out) { ids = static_cast<uint32_t*>(g_pin.alloc(owned * 8, &e)); if (ids) e = cudaMemcpyAsync(ids, frame->d_hit_id, owned * 8, cudaMemcpyDeviceToHost, g_stream); bytes += owned * 8; }
        if (e == cudaSuccess && hit_t_out) { ts = static_cast<double*>(g_pin.alloc(owned * 8, &e)); if (ts) e = cudaMemcpyAsync(ts, frame->d_hit_t, owned * 8, cudaMemcpyDeviceToHost, g_stream); bytes += owned * 8; }
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
        if (e == cudaSuccess) {
            // write only the pixels this call owns (render.rs:136-138)
            const uint32_t* index = frame->pixel_index.data();
            for (size_t k = 0; k < owned; ++k) {
                const size_t px = index[k];
                if (rgb) { rgb_inout[px * 3] = rgb[k * 3]; rgb_inout[px * 3 + 1] = rgb[k * 3 + 1]; rgb_inout[px * 3 + 2] = rgb[k * 3 + 2]; }
                if (ids) { hit_id_out[px * 2] = ids[k * 2]; hit_id_out[px * 2 + 1] = ids[k * 2 + 1]; }
                if (ts) hit_t_out[px] = ts[k];
            }
        }
        g_pin.release(rgb);
        g_pin.release(ids);
        g_pin.release(ts);
        if (e != cudaSuccess) return fail(PT_ERR_CUDA, "frame read failed: %s", cudaGetErrorString(e));
    }
    if (stats) { stats->d2h_ms += now_ms() - t0; stats->d2h_bytes += bytes; }
    return PT_OK;
}

// =================================================================== one-call render (the Rust shim's entry)
static bool same_geometry(const PtRenderParams& a, const PtRenderParams& b) {
    return a.width == b.width && a.height == b.height && a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2 &&
           a.samples == b.samples && a.bg_mode == b.bg_mode && a.max_depth == b.max_depth && a.tile_w == b.tile_w &&
           a.tile_h == b.tile_h && a.rank == b.rank && a.world == b.world && a.max_batch_paths == b.max_batch_paths &&
           a.node_pool_capacity == b.node_pool_capacity && ((a.flags ^ b.flags) & PT_RENDER_ROW_MAJOR) == 0;
}

// The frame pt_render uses for (scene, params) on the CURRENT context.  Frames (device buffers + CUDA graph) are kept
// between calls and re-bound to whatever scene comes next: a program that renders many scenes at one resolution
// (examples/normal-mapping.rs) allocates once.
static int cached_frame(PtScene* scene, const PtCamera* camera, const PtRenderParams& p, PtFrame** out) {
    PtFrame* f = nullptr;
    for (PtFrame* c : g_frame_cache)
        if (same_geometry(c->params, p) && c->n_lights_cap == scene->h.n_lights && c->reflective_cap == scene->has_reflective) { f = c; break; }
    if (f) {
        f->scene = scene;
        f->cam = *camera;
        f->params = p;
    } else {
        int rc = create_frame(scene, camera, &p, &f);
        if (rc != PT_OK) return rc;
        if (g_frame_cache.size() >= kFrameCacheMax) {
            size_t oldest = 0;
            for (size_t i = 1; i < g_frame_cache.size(); ++i)
                if (g_frame_cache[i]->last_use < g_frame_cache[oldest]->last_use) oldest = i;
            free_frame(g_frame_cache[oldest]);
            g_frame_cache.erase(g_frame_cache.begin() + oldest);
        }
        g_frame_cache.push_back(f);
        // ... and by bytes: node pools of big frames are GBs each; idle frames may hold a quarter of the device at most
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); total_b = 0; }
        for (;;) {
            size_t held = 0, oldest = g_frame_cache.size();
            for (size_t i = 0; i < g_frame_cache.size(); ++i) {
                const PtFrame* c = g_frame_cache[i];
                held += pool_bytes(c->pool.capacity, c->n_lights_cap);
                if (c != f && (oldest == g_frame_cache.size() || c->last_use < g_frame_cache[oldest]->last_use)) oldest = i;
            }
            if (held <= total_b / 4 || oldest == g_frame_cache.size()) break;
            free_frame(g_frame_cache[oldest]);
            g_frame_cache.erase(g_frame_cache.begin() + oldest);
        }
    }
    f->last_use = ++g_frame_tick;
    f->image_target = nullptr;
    *out = f;
    return PT_OK;
}

// sums / maxima of the members' statistics of one group render
static void merge_stats(PtStats* into, const PtStats& s) {
    into->rays_primary += s.rays_primary; into->rays_shadow += s.rays_shadow; into->rays_reflect += s.rays_reflect;
    into->rays_refract += s.rays_refract; into->rays_depth_cut += s.rays_depth_cut;
    into->kd_splits += s.kd_splits; into->instance_tests += s.instance_tests; into->triangle_tests += s.triangle_tests;
    into->bbox_gates += s.bbox_gates; into->shaded_hits += s.shaded_hits; into->texel_lookups += s.texel_lookups;
    into->nodes_total += s.nodes_total; into->batches += s.batches; into->retries += s.retries;
    into->max_level = std::max(into->max_level, s.max_level);
    into->device_error_bits |= s.device_error_bits;
    into->kernel_launches += s.kernel_launches;
    into->device_ms = std::max(into->device_ms, s.device_ms);  // the members run concurrently
    into->h2d_ms += s.h2d_ms; into->d2h_ms += s.d2h_ms; into->h2d_bytes += s.h2d_bytes; into->d2h_bytes += s.d2h_bytes;
    for (int k = 0; k < 2; ++k) {
        into->k_kd_splits[k] += s.k_kd_splits[k]; into->k_instance_tests[k] += s.k_instance_tests[k];
        into->k_triangle_tests[k] += s.k_triangle_tests[k]; into->k_bbox_gates[k] += s.k_bbox_gates[k];
        into->k_prim_flops[k] += s.k_prim_flops[k];
        into->x_box_tests[k] += s.x_box_tests[k]; into->x_instance_tests[k] += s.x_instance_tests[k];
        into->x_triangle_tests[k] += s.x_triangle_tests[k]; into->x_bbox_gates[k] += s.x_bbox_gates[k];
        into->x_prim_flops[k] += s.x_prim_flops[k];
    }
    into->ms_extend = std::max(into->ms_extend, s.ms_extend); into->ms_shadow = std::max(into->ms_shadow, s.ms_shadow);
    into->ms_shade = std::max(into->ms_shade, s.ms_shade);
    into->n_extend += s.n_extend; into->n_shadow += s.n_shadow; into->n_shade += s.n_shade;
    for (int l = 0; l < 16; ++l) {
        into->ms_extend_level[l] = std::max(into->ms_extend_level[l], s.ms_extend_level[l]);
        into->ms_shadow_level[l] = std::max(into->ms_shadow_level[l], s.ms_shadow_level[l]);
    }
    if (s.err_bit && !into->err_bit) {
        into->err_bit = s.err_bit; into->err_pixel = s.err_pixel; into->err_sample = s.err_sample;
        into->err_pathid = s.err_pathid; into->err_where = s.err_where;
    }
}

// pt_render over a device group: the scene's copies render interleaved tiles concurrently (member i = rank r0 + w0 * i
// of world w0 * n, r0 / w0 the caller's own rank / world), every resolve kernel stores its pixels into one image on the
// primary, one D2H brings the slice rectangle to the host.
static int render_group(PtScene* scene, const PtCamera* camera, const PtRenderParams& p0, const double* background,
                        uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out, PtProgressFn progress, void* user,
                        PtStats* stats) {
    const int n = 1 + (int)scene->replicas.size();
    const uint32_t w0 = std::max<uint32_t>(p0.world, 1), r0 = p0.world > 1 ? p0.rank : 0;
    std::vector<PtScene*> members(n);
    std::vector<PtFrame*> frames(n, nullptr);
    members[0] = scene;
    for (int i = 1; i < n; ++i) members[i] = scene->replicas[i - 1];
    // one image on the primary needs every member to reach it, and every slice pixel to be rendered by this call
    bool image_path = p0.world <= 1;
    for (int i = 0; i < n; ++i) image_path = image_path && members[i]->ctx->peer_to_primary;
    PtStats total{};
    for (int i = 0; i < n; ++i) {
        CtxScope member(members[i]->ctx);
        PtRenderParams pi = p0;
        pi.world = w0 * (uint32_t)n;
        pi.rank = r0 + w0 * (uint32_t)i;  // the caller's tiles are k = r0 (mod w0): member i takes every n-th of THOSE
        pi.flags &= ~PT_RENDER_ROW_MAJOR;
        int rc = cached_frame(members[i], camera, pi, &frames[i]);
        if (rc == PT_OK && (hit_id_out || hit_t_out)) rc = ensure_id_buffers(frames[i]);
        if (rc != PT_OK) return rc;
    }
    DeviceCtx* primary = scene->ctx;
    const uint64_t image_bytes = (uint64_t)p0.width * p0.height * 3;
    if (image_path) {
        CtxScope scope(primary);
        if (primary->group_image_bytes < image_bytes) {
            g_dev.release(primary->group_image);
            cudaError_t e;
            primary->group_image = static_cast<uint8_t*>(g_dev.alloc(image_bytes, &e));
            primary->group_image_bytes = primary->group_image ? image_bytes : 0;
            if (!primary->group_image) return fail(PT_ERR_CUDA, "group image allocation failed: %s", cudaGetErrorString(e));
        }
        for (PtFrame* f : frames) f->image_target = primary->group_image;
    }
    // background: host -> primary, primary -> the other members device to device (a per-pixel 4K background is 199 MB)
    const double t0 = now_ms();
    const uint64_t bg_bytes = frames[0]->bg_doubles * sizeof(double);
    cudaEvent_t bg_ready = nullptr;
    {
        CtxScope scope(primary);
        CUDA_TRY(cudaMemcpyAsync(frames[0]->d_background, background, bg_bytes, cudaMemcpyHostToDevice, g_stream));
        CUDA_TRY(cudaEventCreateWithFlags(&bg_ready, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(bg_ready, g_stream));
    }
    for (int i = 1; i < n; ++i) {
        CtxScope member(members[i]->ctx);
        cudaError_t e = cudaStreamWaitEvent(g_stream, bg_ready, 0);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(frames[i]->d_background, g_cur->device, frames[0]->d_background, primary->device, bg_bytes, g_stream);
        if (e != cudaSuccess) { cudaEventDestroy(bg_ready); return fail(PT_ERR_CUDA, "background replication failed: %s", cudaGetErrorString(e)); }
    }
    total.h2d_ms = now_ms() - t0;
    total.h2d_bytes = bg_bytes;
    // all members enqueue before anyone waits
    int rc = PT_OK;
    int enqueued = 0;
    for (int i = 0; i < n && rc == PT_OK; ++i) {
        frames[i]->progress = progress;
        frames[i]->progress_user = user;
        rc = pt_frame_enqueue(frames[i], nullptr);
        if (rc == PT_OK) ++enqueued;
    }
    for (int i = 0; i < enqueued; ++i) {
        PtStats st{};
        const int rc_i = pt_frame_finish(frames[i], &st);
        if (rc_i != PT_OK && rc == PT_OK) rc = rc_i;
        merge_stats(&total, st);
    }
    for (PtFrame* f : frames) { f->progress = nullptr; f->progress_user = nullptr; }
    cudaEventDestroy(bg_ready);
    int rc2 = PT_OK;
    if (image_path) {
        CtxScope scope(primary);
        const double t1 = now_ms();
        const size_t W = p0.width, w = p0.x2 - p0.x1 + 1, hgt = p0.y2 - p0.y1 + 1;
        const size_t first = (size_t)p0.y1 * W + p0.x1;
        cudaError_t e = cudaMemcpy2DAsync(rgb_inout + first * 3, W * 3, primary->group_image + first * 3, W * 3, w * 3, hgt, cudaMemcpyDeviceToHost, g_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) rc2 = fail(PT_ERR_CUDA, "group image read failed: %s", cudaGetErrorString(e));
        total.d2h_ms += now_ms() - t1;
        total.d2h_bytes += w * hgt * 3;
    }
    for (int i = 0; i < n && rc2 == PT_OK; ++i)
        if (!image_path || hit_id_out || hit_t_out) rc2 = pt_frame_read(frames[i], image_path ? nullptr : rgb_inout, hit_id_out, hit_t_out, &total);
    if (stats) *stats = total;
    return rc != PT_OK ? rc : rc2;
}

int pt_render(PtScene* scene, const PtCamera* camera, const PtRenderParams* params, const double* background,
              uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out, PtProgressFn progress, void* user,
              PtStats* stats) {
    if (!scene || !camera || !params || !background || !rgb_inout) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    if (!scene->replicas.empty())
        return render_group(scene, camera, *params, background, rgb_inout, hit_id_out, hit_t_out, progress, user, stats);
    CtxScope scope(scene->ctx);
    PtRenderParams p = *params;
    if (p.world <= 1) p.flags |= PT_RENDER_ROW_MAJOR;  // single rank: resolve writes the image in place, D2H is one 2-D copy
    else p.flags &= ~PT_RENDER_ROW_MAJOR;
    PtFrame* f = nullptr;
    int rc = cached_frame(scene, camera, p, &f);
    if (rc != PT_OK) return rc;
    if (hit_id_out || hit_t_out) {
        rc = ensure_id_buffers(f);
        if (rc != PT_OK) return rc;
    }
    PtStats local{};
    const double t0 = now_ms();
    CUDA_TRY(cudaMemcpyAsync(f->d_background, background, f->bg_doubles * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    local.h2d_ms = now_ms() - t0;
    local.h2d_bytes = f->bg_doubles * sizeof(double);
    rc = pt_frame_render(f, nullptr, progress, user, &local);
    int rc2 = pt_frame_read(f, rgb_inout, hit_id_out, hit_t_out, &local);
    if (stats) *stats = local;
    return rc != PT_OK ? rc : rc2;
}

// =================================================================== texture ingest / PNG encode (image_io.cu)
int pt_texture_ingest(const void* pixels, uint32_t width, uint32_t height, uint32_t layout, uint64_t key) {
    if (!pixels || width == 0 || height == 0 || key == 0) return fail(PT_ERR_INVALID, "null pixels, empty image or zero key");
    uint32_t channels = 0, r_at = 0, g_at = 0, b_at = 0;
    switch (layout) {
        case PT_PIXELS_LUMA8: channels = 1; break;
        case PT_PIXELS_LUMAA8: channels = 2; break;
        case PT_PIXELS_RGB8: channels = 3; g_at = 1; b_at = 2; break;
        case PT_PIXELS_RGBA8: channels = 4; g_at = 1; b_at = 2; break;
        case PT_PIXELS_BGR8: channels = 3; r_at = 2; g_at = 1; break;
        case PT_PIXELS_BGRA8: channels = 4; r_at = 2; g_at = 1; break;
        default: return fail(PT_ERR_INVALID, "unknown pixel layout %u", layout);
    }
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    const uint64_t n = (uint64_t)width * height, bytes = n * 3;
    uint8_t* d_primary = nullptr;
    for (int i = 0; i < g_group_size; ++i) {
        CtxScope member(&g_ctxs[i]);
        auto it = g_textures.find(key);
        if (it != g_textures.end()) {
            if (it->second.width == width && it->second.height == height) { if (i == 0) d_primary = it->second.d_texels; continue; }  // already resident
            if (it->second.refs) return fail(PT_ERR_INVALID, "texture key %llx is in use with another size", (unsigned long long)key);
            g_dev.release(it->second.d_texels);
            g_texture_bytes -= it->second.bytes;
            g_textures.erase(it);
        }
        evict_textures(bytes);
        cudaError_t e;
        uint8_t* d = static_cast<uint8_t*>(g_dev.alloc(bytes, &e));
        if (!d) return fail(PT_ERR_CUDA, "texture allocation failed: %s", cudaGetErrorString(e));
        if (i == 0 || !d_primary) {
            uint8_t* d_src = static_cast<uint8_t*>(g_dev.alloc(n * channels, &e));
            if (!d_src) { g_dev.release(d); return fail(PT_ERR_CUDA, "texture staging allocation failed: %s", cudaGetErrorString(e)); }
            e = cudaMemcpyAsync(d_src, pixels, n * channels, cudaMemcpyHostToDevice, g_stream);
            if (e == cudaSuccess) e = ptd::launch_to_rgb(d_src, n, channels, r_at, g_at, b_at, d, g_stream);
            g_dev.release(d_src);  // stream-ordered reuse
            if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);  // the caller may free `pixels`; the group's copies read d
            if (e != cudaSuccess) { g_dev.release(d); return fail(PT_ERR_CUDA, "texture ingest failed: %s", cudaGetErrorString(e)); }
            if (i == 0) d_primary = d;
        } else {  // device group: the converted texels from the primary, device to device
            e = cudaMemcpyAsync(d, d_primary, bytes, cudaMemcpyDefault, g_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
            if (e != cudaSuccess) { g_dev.release(d); return fail(PT_ERR_CUDA, "texture replication failed: %s", cudaGetErrorString(e)); }
        }
        ResidentTexture r;
        r.d_texels = d; r.bytes = bytes; r.width = width; r.height = height; r.refs = 0; r.last_use = ++g_texture_tick;
        g_textures.emplace(key, r);
        g_texture_bytes += bytes;
    }
    return PT_OK;
}

int pt_texture_read(uint64_t key, uint8_t* rgb_out, uint64_t capacity, uint32_t* width_out, uint32_t* height_out) {
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    CtxScope scope(&g_ctxs[0]);
    auto it = g_textures.find(key);
    if (it == g_textures.end()) return fail(PT_ERR_INVALID, "texture key %llx is not resident", (unsigned long long)key);
    if (width_out) *width_out = it->second.width;
    if (height_out) *height_out = it->second.height;
    if (!rgb_out) return PT_OK;
    if (capacity < it->second.bytes) return fail(PT_ERR_INVALID, "buffer too small: %llu bytes needed", (unsigned long long)it->second.bytes);
    CUDA_TRY(cudaMemcpyAsync(rgb_out, it->second.d_texels, it->second.bytes, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PT_OK;
}

uint64_t pt_png_size(uint32_t width, uint32_t height) { return (width && height) ? ptd::png_file_bytes(width, height) : 0; }

int pt_png_encode_device(const uint8_t* d_rgb, uint32_t width, uint32_t height, uint8_t* d_png_out, void* stream) {
    if (!d_rgb || !d_png_out || width == 0 || height == 0) return fail(PT_ERR_INVALID, "null argument or empty image");
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : g_stream;
    cudaError_t e;
    void* scratch = g_dev.alloc(ptd::png_scratch_bytes(width, height), &e);
    if (!scratch) return fail(PT_ERR_CUDA, "PNG scratch allocation failed: %s", cudaGetErrorString(e));
    e = ptd::launch_png_encode(d_rgb, width, height, d_png_out, scratch, st);
    if (e == cudaSuccess && (!stream || st != g_stream)) e = cudaStreamSynchronize(st);  // the scratch goes back to the arena, which orders reuse on g_stream only
    g_dev.release(scratch);
    if (e == cudaErrorInvalidValue) return fail(PT_ERR_INVALID, "image too large for a single-IDAT PNG");
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "PNG encode failed: %s", cudaGetErrorString(e));
    return PT_OK;
}

// device image -> host file through a pinned staging buffer
static int png_to_host(const uint8_t* d_rgb, uint32_t width, uint32_t height, uint8_t* png_out, uint64_t capacity, uint64_t* png_bytes_out) {
    const uint64_t bytes = ptd::png_file_bytes(width, height);
    if (png_bytes_out) *png_bytes_out = bytes;
    if (!png_out || capacity < bytes) return fail(PT_ERR_INVALID, "PNG buffer too small: %llu bytes needed", (unsigned long long)bytes);
    cudaError_t e;
    uint8_t* d_png = static_cast<uint8_t*>(g_dev.alloc(bytes, &e));
    if (!d_png) return fail(PT_ERR_CUDA, "PNG allocation failed: %s", cudaGetErrorString(e));
    int rc = pt_png_encode_device(d_rgb, width, height, d_png, g_stream);
    if (rc == PT_OK) {
        e = cudaMemcpyAsync(png_out, d_png, bytes, cudaMemcpyDeviceToHost, g_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
        if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, "PNG download failed: %s", cudaGetErrorString(e));
    }
    g_dev.release(d_png);
    return rc;
}

int pt_png_encode(const uint8_t* rgb, uint32_t width, uint32_t height, uint8_t* png_out, uint64_t capacity, uint64_t* png_bytes_out) {
    if (!rgb || width == 0 || height == 0) return fail(PT_ERR_INVALID, "null argument or empty image");
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    const uint64_t n = (uint64_t)width * height * 3;
    cudaError_t e;
    uint8_t* d_rgb = static_cast<uint8_t*>(g_dev.alloc(n, &e));
    if (!d_rgb) return fail(PT_ERR_CUDA, "image allocation failed: %s", cudaGetErrorString(e));
    e = cudaMemcpyAsync(d_rgb, rgb, n, cudaMemcpyHostToDevice, g_stream);
    rc = e == cudaSuccess ? png_to_host(d_rgb, width, height, png_out, capacity, png_bytes_out) : fail(PT_ERR_CUDA, "image upload failed: %s", cudaGetErrorString(e));
    g_dev.release(d_rgb);
    return rc;
}

int pt_frame_encode_png(PtFrame* frame, uint8_t* png_out, uint64_t capacity, uint64_t* png_bytes_out) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    Lock lock(g_mu);
    CtxScope scope(frame->ctx);
    if (frame->pending == PtFrame::IN_FLIGHT) return fail(PT_ERR_INVALID, "a render of this frame is in flight");
    if (!frame->row_major) return fail(PT_ERR_INVALID, "the frame's device image is in compact owned-pixel order (world > 1): encode the gathered image instead");
    return png_to_host(frame->d_rgb, frame->params.width, frame->params.height, png_out, capacity, png_bytes_out);
}

// Ray::color(scene, background, 0) for explicit rays (ray.rs:139-148) — what the
// reference's own mesh_equivalence test evaluates per ray (kdmesh.rs:155-163).
int pt_trace_rays(PtScene* scene, uint64_t n, const double* origins, const double* dirs, const double* background3,
                  uint32_t rng_mode, uint64_t seed, uint32_t max_depth, uint32_t flags, double* color_out,
                  uint32_t* hit_id_out, double* hit_t_out, PtStats* stats) {
    if (!scene || !origins || !dirs || !background3) return fail(PT_ERR_INVALID, "null argument");
    if (n == 0) return PT_OK;
    if (n > 0x7FFFFFFFull) return fail(PT_ERR_INVALID, "too many rays");
    Lock lock(g_mu);
    CtxScope scope(scene->ctx);
    const uint32_t depth = max_depth ? max_depth : PT_MAX_RECURSION_DEPTH;
    if (depth > PT_MAX_DEPTH_SUPPORTED) return fail(PT_ERR_INVALID, "max_depth > %u is not supported", PT_MAX_DEPTH_SUPPORTED);
    const int n_levels = scene->has_reflective ? (int)depth + 1 : 1;
    const int count = walk_mode(flags);
    const uint32_t batch = (uint32_t)std::min<uint64_t>(n, 1u << 20);
    const uint32_t capacity = scene->has_reflective ? batch * 8 : batch;

    double *d_o = nullptr, *d_d = nullptr, *d_bg = nullptr, *d_color = nullptr, *d_t = nullptr;
    uint32_t* d_id = nullptr;
    unsigned char* d_pool = nullptr;
    BatchCtl *d_ctl = nullptr, *h_ctl = nullptr;
    NodePool pool{};
    int rc = PT_OK;
    cudaError_t e = cudaSuccess;
    auto dev = [&](size_t bytes) -> void* { return e == cudaSuccess ? g_dev.alloc(bytes, &e) : nullptr; };
    d_o = static_cast<double*>(dev(n * 24)); d_d = static_cast<double*>(dev(n * 24)); d_bg = static_cast<double*>(dev(24));
    d_color = static_cast<double*>(dev(n * 24)); d_t = static_cast<double*>(dev(n * 8)); d_id = static_cast<uint32_t*>(dev(n * 8));
    d_ctl = static_cast<BatchCtl*>(dev(sizeof(BatchCtl)));
    if (e == cudaSuccess) h_ctl = static_cast<BatchCtl*>(g_pin.alloc(sizeof(BatchCtl), &e));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_o, origins, n * 24, cudaMemcpyHostToDevice, g_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_d, dirs, n * 24, cudaMemcpyHostToDevice, g_stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_bg, background3, 24, cudaMemcpyHostToDevice, g_stream);
    if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, "trace_rays allocation failed: %s", cudaGetErrorString(e));
    if (rc == PT_OK) rc = alloc_pool(&d_pool, &pool, capacity, scene->h.n_lights);

    const int slot = take_slot();
    FrameState state{};
    state.sc = scene->view;
    state.fp.pixel_index = nullptr;
    state.fp.background = d_bg;
    state.fp.bg_mode = PT_BG_CONSTANT;
    state.fp.width = 1; state.fp.height = 1; state.fp.samples = 1;
    state.fp.rng_mode = rng_mode; state.fp.seed = seed; state.fp.max_depth = depth;
    state.pool = pool;
    state.ctl = d_ctl;
    state.hit_id = d_id;
    state.hit_t = d_t;
    state.n_levels = (uint32_t)n_levels;
    state.ray_origins = d_o;
    state.ray_dirs = d_d;
    state.ray_color = d_color;
    if (rc == PT_OK) {
        e = upload_state(slot, state, g_stream);
        if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, "trace_rays state upload failed: %s", cudaGetErrorString(e));
    }
    if (stats) memset(stats, 0, sizeof *stats);
    uint32_t launches = 0, error_bits = 0;
    uint32_t step = batch;
    for (uint64_t first = 0; rc == PT_OK && first < n;) {
        const uint32_t n_paths = (uint32_t)std::min<uint64_t>(step, n - first);
        launch_load_rays(slot, (uint32_t)first, n_paths, g_stream);
        ++launches;
        KernelTimer timer;
        rc = run_levels_stream(slot, scene->h.n_lights, capacity, d_ctl, h_ctl, n_paths, n_levels, count, g_stream, &launches, &timer, (flags & PT_RENDER_LINEAR_TLAS) != 0);
        if (rc != PT_OK) break;
        if (h_ctl->error_bits & PT_DEVERR_OVERFLOW) {
            if (n_paths == 1) { rc = fail(PT_ERR_OVERFLOW, "%s", panic_text(PT_ERR_OVERFLOW)); break; }
            step = std::max<uint32_t>(1, n_paths / 2);
            continue;
        }
        launch_tree_eval(slot, n_paths, g_stream);
        launch_export_rays(slot, n_paths, g_stream);
        launches += 2;
        error_bits |= h_ctl->error_bits;
        accumulate_stats(stats, *h_ctl, n_paths);
        first += n_paths;
    }
    if (rc == PT_OK) {
        e = cudaStreamSynchronize(g_stream);
        if (e == cudaSuccess && color_out) e = cudaMemcpy(color_out, d_color, n * 24, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && hit_id_out) e = cudaMemcpy(hit_id_out, d_id, n * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && hit_t_out) e = cudaMemcpy(hit_t_out, d_t, n * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, "trace_rays copy back failed: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(g_stream);
    }
    give_slot(slot);
    g_dev.release(d_o); g_dev.release(d_d); g_dev.release(d_bg); g_dev.release(d_color); g_dev.release(d_t); g_dev.release(d_id);
    g_dev.release(d_pool); g_dev.release(d_ctl);
    g_pin.release(h_ctl);
    if (stats) { stats->kernel_launches = launches; stats->device_error_bits = error_bits & ~PT_DEVERR_OVERFLOW; }
    if (rc != PT_OK) return rc;
    const int code = device_error_to_code(error_bits);
    if (code != PT_OK) return fail(code, "%s", panic_text(code));
    return PT_OK;
}


// ------------------------------------------------------------------- k-d tree build (kd_build.cu)
static int kd_config_ok(const PtKdBuildConfig* c) {
    if (!c) return fail(PT_ERR_INVALID, "null config");
    if (c->max_depth > PT_MAX_KD_STACK) return fail(PT_ERR_KD_TOO_DEEP, "%s", panic_text(PT_ERR_KD_TOO_DEEP));
    return PT_OK;
}

int pt_kd_build_device(const double* d_bounds, uint32_t n, const PtKdBuildConfig* config, void* stream, PtKdTree** out) {
    if (!out || (n && !d_bounds)) return fail(PT_ERR_INVALID, "null argument");
    int rc = kd_config_ok(config);
    if (rc != PT_OK) return rc;
    Lock lock(g_mu);
    rc = ensure_init();
    if (rc != PT_OK) return rc;
    ptd::KdTreeDev* dev = nullptr;
    ptd::KdAllocator al = arena_allocator();
    const cudaError_t e = ptd::kd_build_device(d_bounds, n, *config, al, stream ? (cudaStream_t)stream : g_stream, &dev);
    if (e == cudaErrorInvalidValue) return fail(PT_ERR_INVALID, "k-d tree does not fit 30-bit node / item indices");
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "k-d tree build failed: %s", cudaGetErrorString(e));
    *out = new PtKdTree{dev, g_cur};
    return PT_OK;
}

int pt_kd_build(const double* bounds, uint32_t n, const PtKdBuildConfig* config, PtKdTree** out) {
    if (!out || (n && !bounds)) return fail(PT_ERR_INVALID, "null argument");
    int rc = kd_config_ok(config);
    if (rc != PT_OK) return rc;
    Lock lock(g_mu);
    rc = ensure_init();
    if (rc != PT_OK) return rc;
    cudaError_t e = cudaSuccess;
    const size_t bytes = std::max<size_t>((size_t)n * 6 * sizeof(double), 8);
    double* d_bounds = static_cast<double*>(g_dev.alloc(bytes, &e));
    if (!d_bounds) return fail(PT_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e));
    if (n) e = cudaMemcpyAsync(d_bounds, bounds, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, g_stream);
    if (e != cudaSuccess) { g_dev.release(d_bounds); return fail(PT_ERR_CUDA, "bounds upload failed: %s", cudaGetErrorString(e)); }
    rc = pt_kd_build_device(d_bounds, n, config, g_stream, out);
    g_dev.release(d_bounds);
    return rc;
}

void pt_kd_tree_free(PtKdTree* tree) {
    if (!tree) return;
    Lock lock(g_mu);
    ptd::kd_tree_release(tree->dev);
    delete tree;
}
uint32_t pt_kd_tree_node_count(const PtKdTree* tree) { return tree ? ptd::kd_tree_node_count(tree->dev) : 0; }
uint32_t pt_kd_tree_item_count(const PtKdTree* tree) { return tree ? ptd::kd_tree_item_count(tree->dev) : 0; }
uint32_t pt_kd_tree_depth(const PtKdTree* tree) { return tree ? ptd::kd_tree_depth(tree->dev) : 0; }

int pt_kd_tree_root_bounds(const PtKdTree* tree, double bounds6_out[6], double* extent_out) {
    if (!tree) return fail(PT_ERR_INVALID, "null tree");
    const double* b = ptd::kd_tree_root_bounds(tree->dev);
    if (bounds6_out) memcpy(bounds6_out, b, 6 * sizeof(double));
    if (extent_out) {
        // (max - min).magnitude_squared(), bounding_box.rs:95-99: x*x + y*y + z*z, left to right, no fused multiply-add
        const volatile double dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
        const volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;
        const volatile double xy = xx + yy;
        *extent_out = xy + zz;
    }
    return PT_OK;
}

int pt_kd_tree_download(const PtKdTree* tree, PtKdNode* nodes_out, uint32_t* items_out) {
    if (!tree) return fail(PT_ERR_INVALID, "null tree");
    Lock lock(g_mu);
    const uint32_t nn = ptd::kd_tree_node_count(tree->dev), ni = ptd::kd_tree_item_count(tree->dev);
    if (nodes_out && nn) CUDA_TRY(cudaMemcpy(nodes_out, ptd::kd_tree_nodes_device(tree->dev), (size_t)nn * sizeof(PtKdNode), cudaMemcpyDeviceToHost));
    if (items_out && ni) CUDA_TRY(cudaMemcpy(items_out, ptd::kd_tree_items_device(tree->dev), (size_t)ni * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return PT_OK;
}

int pt_kd_tree_build_stats(const PtKdTree* tree, double* device_ms_out, uint32_t* launches_out, uint64_t* algorithmic_bytes_out) {
    if (!tree) return fail(PT_ERR_INVALID, "null tree");
    if (device_ms_out) *device_ms_out = ptd::kd_tree_device_ms(tree->dev);
    if (launches_out) *launches_out = ptd::kd_tree_launches(tree->dev);
    if (algorithmic_bytes_out) *algorithmic_bytes_out = ptd::kd_tree_algorithmic_bytes(tree->dev);
    return PT_OK;
}

int pt_scene_set_tlas(PtScene* scene, const PtKdTree* tree) {
    if (!scene || !tree) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    if (!scene->replicas.empty()) return fail(PT_ERR_INVALID, "not supported for a scene replicated over a device group (pt_init_devices): upload it on one device");
    CtxScope scope(scene->ctx);
    const uint32_t nn = ptd::kd_tree_node_count(tree->dev), ni = ptd::kd_tree_item_count(tree->dev);
    if (ptd::kd_tree_depth(tree->dev) > PT_MAX_KD_STACK) return fail(PT_ERR_KD_TOO_DEEP, "%s", panic_text(PT_ERR_KD_TOO_DEEP));
    cudaError_t e = cudaSuccess;
    PtKdNode* d_nodes = static_cast<PtKdNode*>(g_dev.alloc(std::max<size_t>(nn, 1) * sizeof(PtKdNode), &e));
    uint32_t* d_items = d_nodes ? static_cast<uint32_t*>(g_dev.alloc(std::max<size_t>(ni, 1) * sizeof(uint32_t), &e)) : nullptr;
    if (!d_nodes || !d_items) { g_dev.release(d_nodes); return fail(PT_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(d_nodes, ptd::kd_tree_nodes_device(tree->dev), (size_t)nn * sizeof(PtKdNode), cudaMemcpyDeviceToDevice, g_stream);
    if (e == cudaSuccess && ni)
        e = cudaMemcpyAsync(d_items, ptd::kd_tree_items_device(tree->dev), (size_t)ni * sizeof(uint32_t), cudaMemcpyDeviceToDevice, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) { g_dev.release(d_nodes); g_dev.release(d_items); return fail(PT_ERR_CUDA, "tree copy failed: %s", cudaGetErrorString(e)); }
    g_dev.release(scene->d_own_tlas_nodes);
    g_dev.release(scene->d_own_tlas_items);
    scene->d_own_tlas_nodes = d_nodes;
    scene->d_own_tlas_items = d_items;
    double extent = 0.0;
    pt_kd_tree_root_bounds(tree, nullptr, &extent);
    scene->h.tlas_extent = extent;
    scene->h.tlas_depth = ptd::kd_tree_depth(tree->dev);
    scene->h.n_tlas_nodes = nn;
    scene->h.n_tlas_items = ni;
    int rc = build_tlas_cull(scene);
    if (rc != PT_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    fill_view(scene);
    return PT_OK;
}


// ------------------------------------------------------------------- scene flattening (flatten.cu)
int pt_flatten(const PtHierNode* nodes, uint32_t n_nodes, const uint32_t* children, uint32_t n_children, uint32_t root,
               const PtGeometryRec* geometries, uint32_t n_geometries, PtFlatScene** out) {
    if (!out || !nodes || n_nodes == 0 || root >= n_nodes || (n_children && !children) || (n_geometries && !geometries))
        return fail(PT_ERR_INVALID, "null or empty argument");
    for (uint32_t i = 0; i < n_nodes; ++i) {  // the device trusts these indices
        const PtHierNode& n = nodes[i];
        if ((uint64_t)n.first_child + n.child_count > n_children) return fail(PT_ERR_INVALID, "node %u: child range outside the child list", i);
        if (n.geometry != 0xFFFFFFFFu && n.geometry >= n_geometries) return fail(PT_ERR_INVALID, "node %u: geometry index out of range", i);
    }
    for (uint32_t i = 0; i < n_children; ++i)
        if (children[i] >= n_nodes) return fail(PT_ERR_INVALID, "child %u: node index out of range", i);
    Lock lock(g_mu);
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    cudaError_t e = cudaSuccess;
    const size_t nb = (size_t)n_nodes * sizeof(PtHierNode), cb = std::max<size_t>((size_t)n_children * 4, 4),
                 gb = std::max<size_t>((size_t)n_geometries * sizeof(PtGeometryRec), 8);
    PtHierNode* d_nodes = static_cast<PtHierNode*>(g_dev.alloc(nb, &e));
    uint32_t* d_children = d_nodes ? static_cast<uint32_t*>(g_dev.alloc(cb, &e)) : nullptr;
    PtGeometryRec* d_geoms = d_children ? static_cast<PtGeometryRec*>(g_dev.alloc(gb, &e)) : nullptr;
    auto cleanup = [&] { g_dev.release(d_nodes); g_dev.release(d_children); g_dev.release(d_geoms); };
    if (!d_geoms) { cleanup(); return fail(PT_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    e = cudaMemcpyAsync(d_nodes, nodes, nb, cudaMemcpyHostToDevice, g_stream);
    if (e == cudaSuccess && n_children) e = cudaMemcpyAsync(d_children, children, (size_t)n_children * 4, cudaMemcpyHostToDevice, g_stream);
    if (e == cudaSuccess && n_geometries)
        e = cudaMemcpyAsync(d_geoms, geometries, (size_t)n_geometries * sizeof(PtGeometryRec), cudaMemcpyHostToDevice, g_stream);
    if (e != cudaSuccess) { cleanup(); return fail(PT_ERR_CUDA, "hierarchy upload failed: %s", cudaGetErrorString(e)); }
    ptd::KdAllocator al = arena_allocator();
    ptd::FlatSceneDev* dev = nullptr;
    e = ptd::flatten_device(d_nodes, n_nodes, d_children, n_children, root, d_geoms, /*max_levels=*/n_nodes, al, g_stream, &dev);
    cleanup();
    if (e == cudaErrorInvalidValue) return fail(PT_ERR_INVALID, "the hierarchy has a cycle or expands to more than 2^31 instances");
    if (e != cudaSuccess) return fail(PT_ERR_CUDA, "flatten failed: %s", cudaGetErrorString(e));
    *out = new PtFlatScene{dev, g_cur};
    return PT_OK;
}

void pt_flat_free(PtFlatScene* flat) {
    if (!flat) return;
    Lock lock(g_mu);
    ptd::flat_release(flat->dev);
    delete flat;
}
uint32_t pt_flat_instance_count(const PtFlatScene* flat) { return flat ? ptd::flat_instance_count(flat->dev) : 0; }
const double* pt_flat_bounds_device(const PtFlatScene* flat) { return flat ? ptd::flat_bounds_device(flat->dev) : nullptr; }

int pt_flat_download(const PtFlatScene* flat, PtInstance* instances_out, PtInstanceTrans* trans_out, double* bounds_out) {
    if (!flat) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    const size_t n = ptd::flat_instance_count(flat->dev);
    if (!n) return PT_OK;
    if (instances_out) CUDA_TRY(cudaMemcpy(instances_out, ptd::flat_instances_device(flat->dev), n * sizeof(PtInstance), cudaMemcpyDeviceToHost));
    if (trans_out) CUDA_TRY(cudaMemcpy(trans_out, ptd::flat_trans_device(flat->dev), n * sizeof(PtInstanceTrans), cudaMemcpyDeviceToHost));
    if (bounds_out) CUDA_TRY(cudaMemcpy(bounds_out, ptd::flat_bounds_device(flat->dev), n * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    return PT_OK;
}

int pt_flat_build_stats(const PtFlatScene* flat, double* device_ms_out, uint32_t* launches_out) {
    if (!flat) return fail(PT_ERR_INVALID, "null argument");
    if (device_ms_out) *device_ms_out = ptd::flat_device_ms(flat->dev);
    if (launches_out) *launches_out = ptd::flat_launches(flat->dev);
    return PT_OK;
}

int pt_scene_set_instances(PtScene* scene, const PtFlatScene* flat, const PtKdTree* tree) {
    if (!scene || !flat || !tree) return fail(PT_ERR_INVALID, "null argument");
    Lock lock(g_mu);
    if (!scene->replicas.empty()) return fail(PT_ERR_INVALID, "not supported for a scene replicated over a device group (pt_init_devices): upload it on one device");
    CtxScope scope(scene->ctx);
    const uint32_t n = ptd::flat_instance_count(flat->dev);
    cudaError_t e = cudaSuccess;
    PtInstance* d_inst = static_cast<PtInstance*>(g_dev.alloc(std::max<size_t>(n, 1) * sizeof(PtInstance), &e));
    PtInstanceTrans* d_tr = d_inst ? static_cast<PtInstanceTrans*>(g_dev.alloc(std::max<size_t>(n, 1) * sizeof(PtInstanceTrans), &e)) : nullptr;
    if (!d_tr) { g_dev.release(d_inst); return fail(PT_ERR_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    if (n) {
        e = cudaMemcpyAsync(d_inst, ptd::flat_instances_device(flat->dev), (size_t)n * sizeof(PtInstance), cudaMemcpyDeviceToDevice, g_stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_tr, ptd::flat_trans_device(flat->dev), (size_t)n * sizeof(PtInstanceTrans), cudaMemcpyDeviceToDevice, g_stream);
    }
    if (e != cudaSuccess) { g_dev.release(d_inst); g_dev.release(d_tr); return fail(PT_ERR_CUDA, "instance copy failed: %s", cudaGetErrorString(e)); }
    g_dev.release(scene->d_own_instances);
    g_dev.release(scene->d_own_instance_trans);
    scene->d_own_instances = d_inst;
    scene->d_own_instance_trans = d_tr;
    scene->h.n_instances = n;
    // the FP32 instance boxes depend on the instances: rebuild them, then splice the tree in (which re-gathers the leaf boxes)
    g_dev.release(scene->d_aabb);
    g_dev.release(scene->d_tri_aabb);
    g_dev.release(scene->d_fold_order);
    scene->d_aabb = scene->d_tri_aabb = nullptr;
    scene->d_fold_order = nullptr;
    int rc = build_instance_bounds(scene, /*with_tlas=*/false);  // the old tree's items index the old instances
    if (rc != PT_OK) return rc;
    return pt_scene_set_tlas(scene, tree);
}

}  // extern "C"
