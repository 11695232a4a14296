// C ABI of libportrayer_gpu.so (include/portrayer_gpu.h): scene upload, frame
// management and the batch loop that drives the wavefront kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"
#include "portrayer_gpu.h"

using namespace ptd;

namespace {

thread_local std::string g_error;
cudaStream_t g_stream = nullptr;
bool g_initialised = false;

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

int ensure_init() {
    if (g_initialised) return PT_OK;
    return pt_init(-1);
}

}  // namespace

struct PtScene {
    unsigned char* d_blob = nullptr;
    uint64_t bytes = 0;
    PtBlobHeader h{};
    DScene view{};
    bool has_reflective = false;
    PtFrame* cached_frame = nullptr;
    PtRenderParams cached_params{};
    PtCamera cached_cam{};
};

struct PtFrame {
    PtScene* scene = nullptr;
    PtCamera cam{};
    PtRenderParams params{};
    std::vector<uint32_t> pixel_index;  // owned slot -> global pixel
    uint32_t* d_pixel_index = nullptr;
    double* d_background = nullptr;
    uint64_t bg_doubles = 0;
    uint8_t* d_rgb = nullptr;
    uint32_t* d_hit_id = nullptr;
    double* d_hit_t = nullptr;
    // node pool
    unsigned char* d_pool = nullptr;
    NodePool pool{};
    uint32_t batch_slots = 0;  // owned pixels per batch
    BatchCtl* d_ctl = nullptr;
    BatchCtl* h_ctl = nullptr;  // pinned
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::vector<cudaEvent_t> kernel_events;  // PT_RENDER_KERNEL_TIMES: begin/end pairs, reused per batch
    uint32_t max_depth = PT_MAX_RECURSION_DEPTH;
    int n_levels = 1;
};

namespace {

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

int device_error_to_code(uint32_t bits) {
    if (bits & PT_DEVERR_NORMALMAP) return PT_ERR_NO_TEXCOORD_NORMALMAP;
    if (bits & PT_DEVERR_TEXTURE) return PT_ERR_NO_TEXCOORD_TEXTURE;
    if (bits & PT_DEVERR_KD_PLANE) return PT_ERR_KD_PLANE_MISS;
    if (bits & PT_DEVERR_TIR) return PT_ERR_TIR_INSIDE;
    return PT_OK;
}

void fill_view(PtScene* s) {
    const PtBlobHeader& h = s->h;
    DScene& v = s->view;
    unsigned char* b = s->d_blob;
    v.tlas_nodes = reinterpret_cast<const PtKdNode*>(b + h.off_tlas_nodes);
    v.tlas_items = reinterpret_cast<const uint32_t*>(b + h.off_tlas_items);
    v.instances = reinterpret_cast<const PtInstance*>(b + h.off_instances);
    v.instance_trans = reinterpret_cast<const PtInstanceTrans*>(b + h.off_instance_trans);
    v.meshes = reinterpret_cast<const PtMesh*>(b + h.off_meshes);
    v.blas_nodes = reinterpret_cast<const PtKdNode*>(b + h.off_blas_nodes);
    v.blas_items = reinterpret_cast<const uint32_t*>(b + h.off_blas_items);
    v.tri_pos = reinterpret_cast<const PtTriPos*>(b + h.off_tri_pos);
    v.tri_normals = reinterpret_cast<const PtTriNormals*>(b + h.off_tri_normals);
    v.tri_uvs = reinterpret_cast<const PtTriUvs*>(b + h.off_tri_uvs);
    v.materials = reinterpret_cast<const PtMaterial*>(b + h.off_materials);
    v.lights = reinterpret_cast<const PtLight*>(b + h.off_lights);
    v.textures = reinterpret_cast<const PtTexture*>(b + h.off_textures);
    v.texels = b + h.off_texels;
    v.ambient[0] = h.ambient[0]; v.ambient[1] = h.ambient[1]; v.ambient[2] = h.ambient[2];
    v.tlas_extent = h.tlas_extent;
    v.n_lights = h.n_lights;
    v.n_instances = h.n_instances;
    v.n_tlas_nodes = h.n_tlas_nodes;
    v.n_tlas_items = h.n_tlas_items;
}

// validate on the host copy, then keep what the host needs to know about the scene
int adopt_header(PtScene* s, const void* host_blob, uint64_t bytes) {
    PtSceneDesc d;
    int rc = pt_scene_unpack(host_blob, bytes, &d);
    if (rc != PT_OK) return fail(rc, "malformed scene blob");
    memcpy(&s->h, host_blob, sizeof s->h);
    if (s->h.tlas_depth > PT_MAX_KD_STACK || s->h.blas_max_depth > PT_MAX_KD_STACK)
        return fail(PT_ERR_KD_TOO_DEEP, "kd-tree depth %u / %u exceeds PT_MAX_KD_STACK = %d", s->h.tlas_depth,
                    s->h.blas_max_depth, PT_MAX_KD_STACK);
    s->has_reflective = false;
    for (uint32_t i = 0; i < d.n_materials; ++i)
        if (d.materials[i].reflectivity > 0.0) s->has_reflective = true;
    return PT_OK;
}

void free_frame(PtFrame* f) {
    if (!f) return;
    cudaFree(f->d_pixel_index);
    cudaFree(f->d_background);
    cudaFree(f->d_rgb);
    cudaFree(f->d_hit_id);
    cudaFree(f->d_hit_t);
    cudaFree(f->d_pool);
    cudaFree(f->d_ctl);
    if (f->h_ctl) cudaFreeHost(f->h_ctl);
    if (f->ev_start) cudaEventDestroy(f->ev_start);
    if (f->ev_stop) cudaEventDestroy(f->ev_stop);
    for (cudaEvent_t e : f->kernel_events) cudaEventDestroy(e);
    delete f;
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
    CUDA_TRY(cudaMalloc(d_pool, off));
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

// Optional per-kernel timing: one begin/end event pair per extend / shadow / shade launch.
struct KernelTimer {
    std::vector<cudaEvent_t>* events = nullptr;  // null: timing off
    size_t used = 0;
    struct Span { int kind; size_t begin; };
    std::vector<Span> spans;
    void begin(int kind, cudaStream_t st) {
        if (!events) return;
        while (events->size() < used + 2) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            events->push_back(e);
        }
        spans.push_back({kind, used});
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
            if (s.kind == 0) { stats->ms_extend += ms; ++stats->n_extend; }
            else if (s.kind == 1) { stats->ms_shadow += ms; ++stats->n_shadow; }
            else { stats->ms_shade += ms; ++stats->n_shade; }
        }
        spans.clear();
        used = 0;
    }
};

// Run one batch of `n_paths` root rays (already written into level 0) through every level.
// Returns the control block in h_ctl (after a stream sync).
int run_levels(const DScene& sc, const FrameParams& fp, const NodePool& pool, BatchCtl* d_ctl, BatchCtl* h_ctl,
               uint32_t first_slot, uint32_t n_paths, int n_levels, bool count, cudaStream_t st, uint32_t* launches,
               KernelTimer* timer) {
    for (int level = 0; level < n_levels; ++level) {
        // level d holds at most n_paths * 2^d rays, and never more than the pool
        unsigned long long bound = (unsigned long long)n_paths << std::min(level, 31);
        const uint32_t max_items = (uint32_t)std::min<unsigned long long>(bound, pool.capacity);
        timer->begin(0, st);
        launch_extend(sc, pool, d_ctl, level, max_items, count, st);
        timer->end(st);
        if (sc.n_lights) {
            timer->begin(1, st);
            launch_shadow(sc, fp, pool, d_ctl, level, first_slot, max_items, count, st);
            timer->end(st);
        }
        timer->begin(2, st);
        launch_shade(sc, fp, pool, d_ctl, level, first_slot, max_items, st);
        timer->end(st);
        *launches += sc.n_lights ? 3 : 2;
    }
    CUDA_TRY(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(BatchCtl), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
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
        stats->kd_splits += c.work[k][0];
        stats->instance_tests += c.work[k][1];
        stats->triangle_tests += c.work[k][2];
        stats->bbox_gates += c.work[k][3];
    }
    stats->shaded_hits += c.shaded_hits;
    stats->texel_lookups += c.texel_lookups;
    stats->nodes_total += std::min<uint32_t>(c.pool_count, 0xFFFFFFFFu);
    stats->device_error_bits |= c.error_bits & ~PT_DEVERR_OVERFLOW;
    for (uint32_t d = 0; d + 1 < 16; ++d)
        if (c.level_start[d + 1] > c.level_start[d] && d > stats->max_level) stats->max_level = d;
}

}  // namespace

// =================================================================== library
extern "C" {

int pt_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PT_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device >= 0) CUDA_TRY(cudaSetDevice(device));
    if (!g_stream) CUDA_TRY(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    kernels_init();
    CUDA_TRY(cudaGetLastError());
    g_initialised = true;
    return PT_OK;
}

void pt_shutdown(void) {
    if (g_stream) cudaStreamDestroy(g_stream);
    g_stream = nullptr;
    g_initialised = false;
}

const char* pt_last_error(void) { return g_error.c_str(); }
const char* pt_error_string(int code) { return panic_text(code); }

int pt_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

// =================================================================== scene
int pt_scene_upload(const void* blob, uint64_t bytes, PtScene** out) {
    if (!blob || !out) return fail(PT_ERR_INVALID, "null argument");
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    PtScene* s = new PtScene();
    rc = adopt_header(s, blob, bytes);
    if (rc != PT_OK) { delete s; return rc; }
    s->bytes = s->h.total_bytes;
    cudaError_t e = cudaMalloc(&s->d_blob, s->bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_blob, blob, s->bytes, cudaMemcpyHostToDevice, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) {
        cudaFree(s->d_blob);
        delete s;
        return fail(PT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    fill_view(s);
    *out = s;
    return PT_OK;
}

int pt_scene_upload_device(const void* d_blob, uint64_t bytes, PtScene** out) {
    if (!d_blob || !out || bytes < sizeof(PtBlobHeader)) return fail(PT_ERR_INVALID, "null argument");
    int rc = ensure_init();
    if (rc != PT_OK) return rc;
    // The records are validated on a host copy (a few MB at most for the reference's scenes; the texel pool is skipped).
    PtBlobHeader h;
    CUDA_TRY(cudaMemcpy(&h, d_blob, sizeof h, cudaMemcpyDeviceToHost));
    if (h.magic != PT_BLOB_MAGIC || h.version != PT_BLOB_VERSION || h.total_bytes > bytes)
        return fail(PT_ERR_INVALID, "malformed scene blob");
    std::vector<unsigned char> host(h.total_bytes);
    const uint64_t records = std::min<uint64_t>(h.off_texels, h.total_bytes);
    CUDA_TRY(cudaMemcpy(host.data(), d_blob, records, cudaMemcpyDeviceToHost));
    PtScene* s = new PtScene();
    rc = adopt_header(s, host.data(), host.size());
    if (rc != PT_OK) { delete s; return rc; }
    s->bytes = h.total_bytes;
    cudaError_t e = cudaMalloc(&s->d_blob, s->bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_blob, d_blob, s->bytes, cudaMemcpyDeviceToDevice, g_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_stream);
    if (e != cudaSuccess) {
        cudaFree(s->d_blob);
        delete s;
        return fail(PT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e));
    }
    fill_view(s);
    *out = s;
    return PT_OK;
}

void pt_scene_free(PtScene* scene) {
    if (!scene) return;
    free_frame(scene->cached_frame);
    cudaFree(scene->d_blob);
    delete scene;
}

// =================================================================== frames
int pt_frame_create(PtScene* scene, const PtCamera* camera, const PtRenderParams* params, PtFrame** out) {
    if (!scene || !camera || !params || !out) return fail(PT_ERR_INVALID, "null argument");
    const PtRenderParams& p = *params;
    if (p.width == 0 || p.height == 0 || p.samples == 0) return fail(PT_ERR_INVALID, "empty image or zero samples");
    if (p.x1 >= p.width || p.x2 >= p.width || p.y1 >= p.height || p.y2 >= p.height)
        return fail(PT_ERR_INVALID, "The positions {x: %u, y: %u} and/or {x: %u, y: %u} are not within an image with width = %u and height = %u",
                    p.x1, p.y1, p.x2, p.y2, p.width, p.height);
    if ((uint64_t)p.width * p.height > 0xFFFFFFF0ull) return fail(PT_ERR_INVALID, "image too large");
    if (p.world > 1 && p.rank >= p.world) return fail(PT_ERR_INVALID, "rank %u out of range for world %u", p.rank, p.world);
    if (p.bg_mode > PT_BG_CONSTANT || p.rng_mode > PT_RNG_HASH) return fail(PT_ERR_INVALID, "bad bg_mode / rng_mode");
    if (effective_max_depth(p) > 13) return fail(PT_ERR_INVALID, "max_depth > 13 is not supported");

    PtFrame* f = new PtFrame();
    f->scene = scene;
    f->cam = *camera;
    f->params = p;
    f->max_depth = effective_max_depth(p);
    f->n_levels = scene->has_reflective ? (int)f->max_depth + 1 : 1;

    // owned pixels: interleaved tiles, 8x4 micro-tiles inside a tile (tiles.c)
    f->pixel_index.resize(pt_owned_pixels(&p, nullptr, 0));
    pt_owned_pixels(&p, f->pixel_index.data(), f->pixel_index.size());
    const uint64_t owned = f->pixel_index.size();

    f->bg_doubles = p.bg_mode == PT_BG_PER_PIXEL ? (uint64_t)p.width * p.height * 3
                    : p.bg_mode == PT_BG_PER_ROW ? (uint64_t)p.height * 3
                                                 : 3;
    // batch geometry: whole pixels per batch so a pixel's samples are summed in one place, in order
    uint64_t max_paths = p.max_batch_paths ? p.max_batch_paths : (1ull << 22);
    uint64_t slots = std::max<uint64_t>(1, max_paths / p.samples);
    slots = std::min<uint64_t>(slots, std::max<uint64_t>(owned, 1));
    if (slots * p.samples > 0x7FFFFFFFull) slots = 0x7FFFFFFFull / p.samples;
    if (slots == 0) { delete f; return fail(PT_ERR_INVALID, "samples too large"); }
    f->batch_slots = (uint32_t)slots;
    const uint64_t batch_paths = slots * p.samples;
    uint64_t capacity = p.node_pool_capacity ? p.node_pool_capacity : (scene->has_reflective ? batch_paths * 4 : batch_paths);
    capacity = std::max<uint64_t>(capacity, batch_paths);
    capacity = std::min<uint64_t>(capacity, 0xFFFFFF00ull);

    cudaError_t e = cudaSuccess;
    auto try_alloc = [&](void** ptr, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(ptr, std::max<size_t>(bytes, 16));
    };
    try_alloc((void**)&f->d_pixel_index, owned * sizeof(uint32_t));
    try_alloc((void**)&f->d_background, f->bg_doubles * sizeof(double));
    try_alloc((void**)&f->d_rgb, owned * 3);
    try_alloc((void**)&f->d_hit_id, owned * 2 * sizeof(uint32_t));
    try_alloc((void**)&f->d_hit_t, owned * sizeof(double));
    try_alloc((void**)&f->d_ctl, sizeof(BatchCtl));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&f->h_ctl, sizeof(BatchCtl));
    if (e == cudaSuccess) e = cudaEventCreate(&f->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&f->ev_stop);
    if (e == cudaSuccess && owned)
        e = cudaMemcpy(f->d_pixel_index, f->pixel_index.data(), owned * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        free_frame(f);
        return fail(PT_ERR_CUDA, "frame allocation failed: %s", cudaGetErrorString(e));
    }
    int rc = alloc_pool(&f->d_pool, &f->pool, (uint32_t)capacity, scene->h.n_lights);
    if (rc != PT_OK) { free_frame(f); return rc; }
    *out = f;
    return PT_OK;
}

void pt_frame_free(PtFrame* frame) { free_frame(frame); }
uint64_t pt_frame_owned_pixels(const PtFrame* frame) { return frame ? frame->pixel_index.size() : 0; }
uint64_t pt_frame_background_doubles(const PtFrame* frame) { return frame ? frame->bg_doubles : 0; }

int pt_frame_set_background(PtFrame* frame, const double* background) {
    if (!frame || !background) return fail(PT_ERR_INVALID, "null argument");
    CUDA_TRY(cudaMemcpyAsync(frame->d_background, background, frame->bg_doubles * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PT_OK;
}
int pt_frame_set_background_device(PtFrame* frame, const double* d_background) {
    if (!frame || !d_background) return fail(PT_ERR_INVALID, "null argument");
    CUDA_TRY(cudaMemcpyAsync(frame->d_background, d_background, frame->bg_doubles * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return PT_OK;
}

int pt_frame_render(PtFrame* frame, void* stream, PtProgressFn progress, void* user, PtStats* stats) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    PtFrame* f = frame;
    cudaStream_t st = stream ? (cudaStream_t)stream : g_stream;
    const DScene& sc = f->scene->view;
    const FrameParams fp = frame_params(f);
    const bool count = (f->params.flags & PT_RENDER_COUNTERS) != 0;
    const uint32_t owned = (uint32_t)f->pixel_index.size();
    const uint32_t S = f->params.samples;
    if (stats) {
        const double h2d_ms = stats->h2d_ms, d2h_ms = stats->d2h_ms;
        const uint64_t h2d_b = stats->h2d_bytes, d2h_b = stats->d2h_bytes;
        memset(stats, 0, sizeof *stats);
        stats->h2d_ms = h2d_ms; stats->d2h_ms = d2h_ms; stats->h2d_bytes = h2d_b; stats->d2h_bytes = d2h_b;
    }
    uint32_t launches = 0, batches = 0, retries = 0, error_bits = 0;
    KernelTimer timer;
    if (f->params.flags & PT_RENDER_KERNEL_TIMES) timer.events = &f->kernel_events;

    CUDA_TRY(cudaEventRecord(f->ev_start, st));
    uint32_t first_slot = 0;
    uint32_t batch_slots = f->batch_slots;
    while (first_slot < owned) {
        const uint32_t n_slots = std::min(batch_slots, owned - first_slot);
        const uint32_t n_paths = n_slots * S;
        launch_begin_batch(f->d_ctl, n_paths, st);
        launch_camera(fp, f->pool, first_slot, n_paths, st);
        launches += 2;
        int rc = run_levels(sc, fp, f->pool, f->d_ctl, f->h_ctl, first_slot, n_paths, f->n_levels, count, st, &launches, &timer);
        if (rc != PT_OK) return rc;
        timer.collect(stats);
        if (f->h_ctl->error_bits & PT_DEVERR_OVERFLOW) {
            // the ray trees of this batch do not fit: halve the batch and redo it (results do not depend on batching)
            if (n_slots == 1) return fail(PT_ERR_OVERFLOW, "%s", panic_text(PT_ERR_OVERFLOW));
            batch_slots = std::max<uint32_t>(1, n_slots / 2);
            ++retries;
            continue;
        }
        launch_tree_eval(fp, f->pool, first_slot, n_paths, st);
        launch_resolve(fp, f->pool, first_slot, n_slots, f->d_rgb, f->d_hit_id, f->d_hit_t, st);
        launches += 2;
        error_bits |= f->h_ctl->error_bits;
        accumulate_stats(stats, *f->h_ctl, n_paths);
        ++batches;
        first_slot += n_slots;
        if (progress) progress(user, n_slots);  // reporter.report_finished_pixels, render.rs:149
    }
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
    const int code = device_error_to_code(error_bits);
    if (code != PT_OK) return fail(code, "%s", panic_text(code));
    return PT_OK;
}

const uint8_t* pt_frame_rgb_device(const PtFrame* frame) { return frame ? frame->d_rgb : nullptr; }
const uint32_t* pt_frame_hit_id_device(const PtFrame* frame) { return frame ? frame->d_hit_id : nullptr; }
const double* pt_frame_hit_t_device(const PtFrame* frame) { return frame ? frame->d_hit_t : nullptr; }

int pt_frame_pixel_index(const PtFrame* frame, uint32_t* index_out) {
    if (!frame || !index_out) return fail(PT_ERR_INVALID, "null argument");
    memcpy(index_out, frame->pixel_index.data(), frame->pixel_index.size() * sizeof(uint32_t));
    return PT_OK;
}

int pt_frame_read(PtFrame* frame, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out, PtStats* stats) {
    if (!frame) return fail(PT_ERR_INVALID, "null frame");
    const size_t owned = frame->pixel_index.size();
    const double t0 = now_ms();
    uint64_t bytes = 0;
    std::vector<uint8_t> rgb;
    std::vector<uint32_t> ids;
    std::vector<double> ts;
    if (rgb_inout) { rgb.resize(owned * 3); CUDA_TRY(cudaMemcpyAsync(rgb.data(), frame->d_rgb, owned * 3, cudaMemcpyDeviceToHost, g_stream)); bytes += owned * 3; }
    if (hit_id_out) { ids.resize(owned * 2); CUDA_TRY(cudaMemcpyAsync(ids.data(), frame->d_hit_id, owned * 8, cudaMemcpyDeviceToHost, g_stream)); bytes += owned * 8; }
    if (hit_t_out) { ts.resize(owned); CUDA_TRY(cudaMemcpyAsync(ts.data(), frame->d_hit_t, owned * 8, cudaMemcpyDeviceToHost, g_stream)); bytes += owned * 8; }
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    // write only the pixels this call owns (render.rs:136-138)
    for (size_t k = 0; k < owned; ++k) {
        const size_t px = frame->pixel_index[k];
        if (rgb_inout) { rgb_inout[px * 3] = rgb[k * 3]; rgb_inout[px * 3 + 1] = rgb[k * 3 + 1]; rgb_inout[px * 3 + 2] = rgb[k * 3 + 2]; }
        if (hit_id_out) { hit_id_out[px * 2] = ids[k * 2]; hit_id_out[px * 2 + 1] = ids[k * 2 + 1]; }
        if (hit_t_out) hit_t_out[px] = ts[k];
    }
    if (stats) { stats->d2h_ms += now_ms() - t0; stats->d2h_bytes += bytes; }
    return PT_OK;
}

// =================================================================== one-call render (the Rust shim's entry)
static bool same_geometry(const PtRenderParams& a, const PtRenderParams& b) {
    return a.width == b.width && a.height == b.height && a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2 &&
           a.samples == b.samples && a.bg_mode == b.bg_mode && a.max_depth == b.max_depth && a.tile_w == b.tile_w &&
           a.tile_h == b.tile_h && a.rank == b.rank && a.world == b.world && a.max_batch_paths == b.max_batch_paths &&
           a.node_pool_capacity == b.node_pool_capacity;
}

int pt_render(PtScene* scene, const PtCamera* camera, const PtRenderParams* params, const double* background,
              uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out, PtProgressFn progress, void* user,
              PtStats* stats) {
    if (!scene || !camera || !params || !background || !rgb_inout) return fail(PT_ERR_INVALID, "null argument");
    // device buffers are kept between calls on the same scene with the same geometry
    PtFrame* f = scene->cached_frame;
    if (f && same_geometry(scene->cached_params, *params)) {
        f->cam = *camera;
        f->params = *params;
    } else {
        free_frame(f);
        scene->cached_frame = nullptr;
        int rc = pt_frame_create(scene, camera, params, &f);
        if (rc != PT_OK) return rc;
        scene->cached_frame = f;
        scene->cached_params = *params;
    }
    PtStats local{};
    const double t0 = now_ms();
    int rc = pt_frame_set_background(f, background);
    if (rc != PT_OK) return rc;
    local.h2d_ms = now_ms() - t0;
    local.h2d_bytes = f->bg_doubles * sizeof(double);
    rc = pt_frame_render(f, nullptr, progress, user, &local);
    int rc2 = pt_frame_read(f, rgb_inout, hit_id_out, hit_t_out, &local);
    if (stats) *stats = local;
    return rc != PT_OK ? rc : rc2;
}

// Ray::color(scene, background, 0) for explicit rays (ray.rs:139-148) — what the
// reference's own mesh_equivalence test evaluates per ray (kdmesh.rs:155-163).
int pt_trace_rays(PtScene* scene, uint64_t n, const double* origins, const double* dirs, const double* background3,
                  uint32_t rng_mode, uint64_t seed, uint32_t max_depth, uint32_t flags, double* color_out,
                  uint32_t* hit_id_out, double* hit_t_out, PtStats* stats) {
    if (!scene || !origins || !dirs || !background3) return fail(PT_ERR_INVALID, "null argument");
    if (n == 0) return PT_OK;
    if (n > 0x7FFFFFFFull) return fail(PT_ERR_INVALID, "too many rays");
    const uint32_t depth = max_depth ? max_depth : PT_MAX_RECURSION_DEPTH;
    if (depth > 13) return fail(PT_ERR_INVALID, "max_depth > 13 is not supported");
    const int n_levels = scene->has_reflective ? (int)depth + 1 : 1;
    const bool count = (flags & PT_RENDER_COUNTERS) != 0;
    const uint32_t batch = (uint32_t)std::min<uint64_t>(n, 1u << 20);
    const uint32_t capacity = scene->has_reflective ? batch * 8 : batch;

    double *d_o = nullptr, *d_d = nullptr, *d_bg = nullptr, *d_color = nullptr, *d_t = nullptr;
    uint32_t* d_id = nullptr;
    unsigned char* d_pool = nullptr;
    BatchCtl *d_ctl = nullptr, *h_ctl = nullptr;
    NodePool pool{};
    int rc = PT_OK;
    cudaError_t e = cudaSuccess;
    auto try_alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
    try_alloc((void**)&d_o, n * 24); try_alloc((void**)&d_d, n * 24); try_alloc((void**)&d_bg, 24);
    try_alloc((void**)&d_color, n * 24); try_alloc((void**)&d_t, n * 8); try_alloc((void**)&d_id, n * 8);
    try_alloc((void**)&d_ctl, sizeof(BatchCtl));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h_ctl, sizeof(BatchCtl));
    if (e == cudaSuccess) e = cudaMemcpy(d_o, origins, n * 24, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_d, dirs, n * 24, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_bg, background3, 24, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail(PT_ERR_CUDA, "trace_rays allocation failed: %s", cudaGetErrorString(e));
    if (rc == PT_OK) rc = alloc_pool(&d_pool, &pool, capacity, scene->h.n_lights);

    FrameParams fp{};
    fp.pixel_index = nullptr;
    fp.background = d_bg;
    fp.bg_mode = PT_BG_CONSTANT;
    fp.width = 1; fp.height = 1; fp.samples = 1;
    fp.rng_mode = rng_mode; fp.seed = seed; fp.max_depth = depth;
    if (stats) memset(stats, 0, sizeof *stats);
    uint32_t launches = 0, error_bits = 0;
    uint32_t step = batch;
    for (uint64_t first = 0; rc == PT_OK && first < n;) {
        const uint32_t n_paths = (uint32_t)std::min<uint64_t>(step, n - first);
        launch_begin_batch(d_ctl, n_paths, g_stream);
        launch_load_rays(d_o, d_d, pool, (uint32_t)first, n_paths, g_stream);
        KernelTimer timer;
        rc = run_levels(scene->view, fp, pool, d_ctl, h_ctl, (uint32_t)first, n_paths, n_levels, count, g_stream, &launches, &timer);
        if (rc != PT_OK) break;
        if (h_ctl->error_bits & PT_DEVERR_OVERFLOW) {
            if (n_paths == 1) { rc = fail(PT_ERR_OVERFLOW, "%s", panic_text(PT_ERR_OVERFLOW)); break; }
            step = std::max<uint32_t>(1, n_paths / 2);
            continue;
        }
        launch_tree_eval(fp, pool, (uint32_t)first, n_paths, g_stream);
        launch_export_rays(pool, (uint32_t)first, n_paths, d_color, d_id, d_t, g_stream);
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
    }
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_bg); cudaFree(d_color); cudaFree(d_t); cudaFree(d_id);
    cudaFree(d_pool); cudaFree(d_ctl);
    if (h_ctl) cudaFreeHost(h_ctl);
    if (stats) { stats->kernel_launches = launches; stats->device_error_bits = error_bits & ~PT_DEVERR_OVERFLOW; }
    if (rc != PT_OK) return rc;
    const int code = device_error_to_code(error_bits);
    if (code != PT_OK) return fail(code, "%s", panic_text(code));
    return PT_OK;
}

}  // extern "C"
