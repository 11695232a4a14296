/*
 * scene_blob.c — pack a PtSceneDesc (pointers) into the pointer-free blob that
 * is uploaded / NCCL-broadcast, and view a blob as a desc again.
 * Pure host C: no CUDA, works on a machine without a GPU.
 */
#include <stdlib.h>
#include <string.h>

#include "portrayer_gpu.h"

#define PT_ALIGN 128ull

static uint64_t align_up(uint64_t v) { return (v + (PT_ALIGN - 1)) & ~(PT_ALIGN - 1); }

typedef struct Layout {
    uint64_t off[14];
    uint64_t len[14];
    uint64_t total;
} Layout;

static void compute_layout(const PtSceneDesc* d, Layout* L) {
    const uint64_t len[14] = {
        (uint64_t)d->n_tlas_nodes * sizeof(PtKdNode),
        (uint64_t)d->n_tlas_items * sizeof(uint32_t),
        (uint64_t)d->n_instances * sizeof(PtInstance),
        (uint64_t)d->n_instances * sizeof(PtInstanceTrans),
        (uint64_t)d->n_meshes * sizeof(PtMesh),
        (uint64_t)d->n_blas_nodes * sizeof(PtKdNode),
        (uint64_t)d->n_blas_items * sizeof(uint32_t),
        (uint64_t)d->n_triangles * sizeof(PtTriPos),
        (uint64_t)d->n_tri_normals * sizeof(PtTriNormals),
        (uint64_t)d->n_tri_uvs * sizeof(PtTriUvs),
        (uint64_t)d->n_materials * sizeof(PtMaterial),
        (uint64_t)d->n_lights * sizeof(PtLight),
        (uint64_t)d->n_textures * sizeof(PtTexture),
        (uint64_t)d->n_texel_bytes,
    };
    uint64_t cur = align_up(sizeof(PtBlobHeader));
    for (int i = 0; i < 14; ++i) {
        L->off[i] = cur;
        L->len[i] = len[i];
        cur = align_up(cur + len[i]);
    }
    L->total = cur;
}

uint64_t pt_scene_blob_size(const PtSceneDesc* desc) {
    if (!desc) return 0;
    Layout L;
    compute_layout(desc, &L);
    return L.total;
}

static uint32_t max_mesh_depth(const PtSceneDesc* d) {
    uint32_t m = 0;
    for (uint32_t i = 0; i < d->n_meshes; ++i)
        if (d->meshes[i].kind == PT_MESH_KD && d->meshes[i].kd_depth > m) m = d->meshes[i].kd_depth;
    return m;
}

int pt_scene_pack(const PtSceneDesc* d, void* blob_out, uint64_t capacity) {
    if (!d || !blob_out) return PT_ERR_INVALID;
    Layout L;
    compute_layout(d, &L);
    if (capacity < L.total) return PT_ERR_INVALID;
    const void* src[14] = {d->tlas_nodes, d->tlas_items,  d->instances, d->instance_trans, d->meshes,
                           d->blas_nodes, d->blas_items,  d->tri_pos,   d->tri_normals,    d->tri_uvs,
                           d->materials,  d->lights,      d->textures,  d->texels};
    for (int i = 0; i < 14; ++i)
        if (L.len[i] && !src[i]) return PT_ERR_INVALID;

    unsigned char* out = (unsigned char*)blob_out;
    memset(out, 0, L.total);
    PtBlobHeader h;
    memset(&h, 0, sizeof h);
    h.magic = PT_BLOB_MAGIC;
    h.version = PT_BLOB_VERSION;
    h.total_bytes = L.total;
    memcpy(h.ambient, d->ambient, sizeof h.ambient);
    h.tlas_extent = d->tlas_extent;
    h.tlas_depth = d->tlas_depth;
    h.blas_max_depth = max_mesh_depth(d);
    h.n_tlas_nodes = d->n_tlas_nodes;   h.n_tlas_items = d->n_tlas_items;
    h.n_instances = d->n_instances;     h.n_meshes = d->n_meshes;
    h.n_blas_nodes = d->n_blas_nodes;   h.n_blas_items = d->n_blas_items;
    h.n_triangles = d->n_triangles;     h.n_tri_normals = d->n_tri_normals;
    h.n_tri_uvs = d->n_tri_uvs;         h.n_materials = d->n_materials;
    h.n_lights = d->n_lights;           h.n_textures = d->n_textures;
    h.n_texel_bytes = d->n_texel_bytes;
    h.off_tlas_nodes = L.off[0];  h.off_tlas_items = L.off[1];  h.off_instances = L.off[2];
    h.off_instance_trans = L.off[3]; h.off_meshes = L.off[4];   h.off_blas_nodes = L.off[5];
    h.off_blas_items = L.off[6];  h.off_tri_pos = L.off[7];     h.off_tri_normals = L.off[8];
    h.off_tri_uvs = L.off[9];     h.off_materials = L.off[10];  h.off_lights = L.off[11];
    h.off_textures = L.off[12];   h.off_texels = L.off[13];
    memcpy(out, &h, sizeof h);
    for (int i = 0; i < 14; ++i)
        if (L.len[i]) memcpy(out + L.off[i], src[i], L.len[i]);
    return PT_OK;
}

static int section_ok(uint64_t off, uint64_t len, uint64_t total) {
    return (off % PT_ALIGN) == 0 && off <= total && len <= total - off;
}

/* Structure check of one tree: indices in range, children after their parent (no cycles), every node the child of at
 * most ONE split (the leaf cull structure derives a leaf's cell from its chain of parents, leaf_cull.cu), and the TRUE
 * depth of the deepest node — the device walk's explicit stack holds PT_MAX_KD_STACK entries and is not bounds-checked,
 * so the depth a blob declares in its header is not trusted.  Returns 1 / 0; -1 when the tree is too deep. */
static int kd_tree_ok(const PtKdNode* nodes, uint32_t n_nodes, uint32_t item_limit, uint32_t* depth_out) {
    uint8_t* depth = (uint8_t*)calloc(n_nodes ? n_nodes : 1, 2); /* depth, then reference count (saturating) */
    if (!depth) return 0;
    uint8_t* refs = depth + n_nodes;
    int ok = 1;
    uint32_t deepest = 0;
    for (uint32_t i = 0; i < n_nodes && ok == 1; ++i) {
        uint32_t axis = nodes[i].a & 3u, hi = nodes[i].a >> 2;
        if (axis == 3u) {
            if ((uint64_t)hi + nodes[i].b > item_limit) ok = 0;
        } else {
            const uint32_t child[2] = {hi, nodes[i].b};
            for (int c = 0; c < 2 && ok == 1; ++c) {
                if (child[c] >= n_nodes || child[c] <= i) { ok = 0; break; } /* children come after their parent: no cycles */
                if (refs[child[c]]++) { ok = 0; break; }                      /* a second parent: not a tree */
                if (depth[i] + 1u > PT_MAX_KD_STACK) { ok = -1; break; }
                depth[child[c]] = (uint8_t)(depth[i] + 1u);
                if (depth[child[c]] > deepest) deepest = depth[child[c]];
            }
        }
    }
    free(depth);
    if (depth_out) *depth_out = deepest;
    return ok;
}

static int unpack_impl(const void* blob, uint64_t bytes, PtSceneDesc* d, int records_only) {
    if (!blob || !d || bytes < sizeof(PtBlobHeader)) return PT_ERR_INVALID;
    PtBlobHeader h;
    memcpy(&h, blob, sizeof h);
    if (h.magic != PT_BLOB_MAGIC || h.version != PT_BLOB_VERSION) return PT_ERR_INVALID;
    if (records_only) {
        /* the texel section may be missing: everything before it must be there */
        if (h.off_texels > h.total_bytes || h.off_texels > bytes) return PT_ERR_INVALID;
    } else if (h.total_bytes > bytes) {
        return PT_ERR_INVALID;
    }
    const uint64_t T = records_only ? h.off_texels : h.total_bytes;
    const unsigned char* base = (const unsigned char*)blob;
    if (!section_ok(h.off_tlas_nodes, (uint64_t)h.n_tlas_nodes * sizeof(PtKdNode), T) ||
        !section_ok(h.off_tlas_items, (uint64_t)h.n_tlas_items * sizeof(uint32_t), T) ||
        !section_ok(h.off_instances, (uint64_t)h.n_instances * sizeof(PtInstance), T) ||
        !section_ok(h.off_instance_trans, (uint64_t)h.n_instances * sizeof(PtInstanceTrans), T) ||
        !section_ok(h.off_meshes, (uint64_t)h.n_meshes * sizeof(PtMesh), T) ||
        !section_ok(h.off_blas_nodes, (uint64_t)h.n_blas_nodes * sizeof(PtKdNode), T) ||
        !section_ok(h.off_blas_items, (uint64_t)h.n_blas_items * sizeof(uint32_t), T) ||
        !section_ok(h.off_tri_pos, (uint64_t)h.n_triangles * sizeof(PtTriPos), T) ||
        !section_ok(h.off_tri_normals, (uint64_t)h.n_tri_normals * sizeof(PtTriNormals), T) ||
        !section_ok(h.off_tri_uvs, (uint64_t)h.n_tri_uvs * sizeof(PtTriUvs), T) ||
        !section_ok(h.off_materials, (uint64_t)h.n_materials * sizeof(PtMaterial), T) ||
        !section_ok(h.off_lights, (uint64_t)h.n_lights * sizeof(PtLight), T) ||
        !section_ok(h.off_textures, (uint64_t)h.n_textures * sizeof(PtTexture), T) ||
        (!records_only && !section_ok(h.off_texels, h.n_texel_bytes, T)))
        return PT_ERR_INVALID;

    memset(d, 0, sizeof *d);
    memcpy(d->ambient, h.ambient, sizeof d->ambient);
    d->tlas_extent = h.tlas_extent;
    d->tlas_depth = h.tlas_depth;
    d->n_tlas_nodes = h.n_tlas_nodes;   d->tlas_nodes = (const PtKdNode*)(base + h.off_tlas_nodes);
    d->n_tlas_items = h.n_tlas_items;   d->tlas_items = (const uint32_t*)(base + h.off_tlas_items);
    d->n_instances = h.n_instances;     d->instances = (const PtInstance*)(base + h.off_instances);
    d->instance_trans = (const PtInstanceTrans*)(base + h.off_instance_trans);
    d->n_meshes = h.n_meshes;           d->meshes = (const PtMesh*)(base + h.off_meshes);
    d->n_blas_nodes = h.n_blas_nodes;   d->blas_nodes = (const PtKdNode*)(base + h.off_blas_nodes);
    d->n_blas_items = h.n_blas_items;   d->blas_items = (const uint32_t*)(base + h.off_blas_items);
    d->n_triangles = h.n_triangles;     d->tri_pos = (const PtTriPos*)(base + h.off_tri_pos);
    d->n_tri_normals = h.n_tri_normals; d->tri_normals = (const PtTriNormals*)(base + h.off_tri_normals);
    d->n_tri_uvs = h.n_tri_uvs;         d->tri_uvs = (const PtTriUvs*)(base + h.off_tri_uvs);
    d->n_materials = h.n_materials;     d->materials = (const PtMaterial*)(base + h.off_materials);
    d->n_lights = h.n_lights;           d->lights = (const PtLight*)(base + h.off_lights);
    d->n_textures = h.n_textures;       d->textures = (const PtTexture*)(base + h.off_textures);
    d->n_texel_bytes = h.n_texel_bytes; d->texels = records_only ? NULL : base + h.off_texels;

    /* semantic validation: every index the kernels will follow must be in range */
    if (d->n_tlas_nodes == 0 || d->n_lights > PT_MAX_LIGHTS) return PT_ERR_INVALID;
    {
        uint32_t depth = 0;
        const int ok = kd_tree_ok(d->tlas_nodes, d->n_tlas_nodes, d->n_tlas_items, &depth);
        if (ok < 0 || depth > PT_MAX_KD_STACK) return PT_ERR_KD_TOO_DEEP;
        if (!ok) return PT_ERR_INVALID;
    }
    for (uint32_t i = 0; i < d->n_tlas_items; ++i)
        if (d->tlas_items[i] >= d->n_instances) return PT_ERR_INVALID;
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        const PtInstance* in = &d->instances[i];
        if (in->prim > PT_PRIM_CONE || in->material >= d->n_materials) return PT_ERR_INVALID;
        if (in->prim == PT_PRIM_TRIANGLE || in->prim == PT_PRIM_MESH || in->prim == PT_PRIM_KDMESH) {
            if (in->mesh >= d->n_meshes) return PT_ERR_INVALID;
            const uint32_t kind = d->meshes[in->mesh].kind;
            if ((in->prim == PT_PRIM_TRIANGLE && kind != PT_MESH_TRIANGLE) ||
                (in->prim == PT_PRIM_MESH && kind != PT_MESH_LINEAR) ||
                (in->prim == PT_PRIM_KDMESH && kind != PT_MESH_KD))
                return PT_ERR_INVALID;
        }
    }
    for (uint32_t i = 0; i < d->n_meshes; ++i) {
        const PtMesh* m = &d->meshes[i];
        if ((uint64_t)m->tri_first + m->tri_count > d->n_triangles) return PT_ERR_INVALID;
        if ((m->flags & PT_MESH_FLAG_NORMALS) && (uint64_t)m->nrm_first + m->tri_count > d->n_tri_normals)
            return PT_ERR_INVALID;
        if ((m->flags & PT_MESH_FLAG_UVS) && (uint64_t)m->uv_first + m->tri_count > d->n_tri_uvs)
            return PT_ERR_INVALID;
        if (m->kind == PT_MESH_KD) {
            if (m->node_count == 0 || (uint64_t)m->node_first + m->node_count > d->n_blas_nodes) return PT_ERR_INVALID;
            if ((uint64_t)m->item_first + m->item_count > d->n_blas_items) return PT_ERR_INVALID;
            uint32_t depth = 0;
            const int ok = kd_tree_ok(d->blas_nodes + m->node_first, m->node_count, m->item_count, &depth);
            if (ok < 0 || depth > PT_MAX_KD_STACK) return PT_ERR_KD_TOO_DEEP;
            if (!ok) return PT_ERR_INVALID;
            for (uint32_t k = 0; k < m->item_count; ++k)
                if (d->blas_items[m->item_first + k] >= m->tri_count) return PT_ERR_INVALID;
        } else if (m->kind == PT_MESH_TRIANGLE) {
            if (m->tri_count != 1) return PT_ERR_INVALID;
        } else if (m->kind != PT_MESH_LINEAR) {
            return PT_ERR_INVALID;
        }
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        if (d->materials[i].texture >= (int32_t)d->n_textures || d->materials[i].normals >= (int32_t)d->n_textures)
            return PT_ERR_INVALID;
    }
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const PtTexture* t = &d->textures[i];
        if (t->width == 0 || t->height == 0) return PT_ERR_INVALID;
        if (t->offset > d->n_texel_bytes || (uint64_t)t->width * t->height * 3 > d->n_texel_bytes - t->offset)
            return PT_ERR_INVALID;
    }
    return PT_OK;
}

int pt_scene_unpack(const void* blob, uint64_t bytes, PtSceneDesc* d) { return unpack_impl(blob, bytes, d, 0); }
int pt_scene_unpack_records(const void* blob, uint64_t bytes, PtSceneDesc* d) { return unpack_impl(blob, bytes, d, 1); }
