/*
 * leafstats — design-space counter for the scene-tree leaf cull (ANALYSIS TOOL, not product code, not a test).
 *
 * Replays a render on the CPU oracle (it #includes oracle/oracle.c with its analysis hooks defined) and counts, per
 * ray class and recursion depth, what different leaf-cull schemes WOULD cost on the device: FP32 slab tests and
 * survivors that reach the exact f64 test.
 *   A  the round-1 scheme: union boxes of aligned runs of 8 leaf positions, then the instance boxes of surviving runs
 *   B  instance boxes CLIPPED to the leaf's kd cell: one "occupied" box per leaf, then runs of 8 (leaf-relative),
 *      then the clipped instance boxes
 * Input: a directory written by tools/leafstats/dump.py (blob.bin, camera.bin, params.bin, background.bin).
 *   build/leafstats <dir> [threads]
 */
#include <stdio.h>

#include <stdint.h>
#define LS_KINDS 4
#define LS_DEPTHS 4
typedef struct {
    uint64_t casts, leaf_visits, candidates;
    uint64_t a_tests, a_exact;
    uint64_t b_tests, b_exact, b_leaf_skipped;
    uint64_t c_tests;                    /* B with 4-ary levels (4, 16, 64) instead of runs of 8 */
    uint64_t empty_visits;               /* leaf visits where no clipped box survives */
    uint64_t blas_leaf_visits, blas_candidates, blas_a_exact, blas_b_exact, blas_b_skipped;
} Cnt;
static Cnt g_cnt[LS_KINDS][LS_DEPTHS];
static __thread int ls_kind, ls_depth;
#define ADD(field, v) __atomic_fetch_add(&g_cnt[ls_kind][ls_depth].field, (uint64_t)(v), __ATOMIC_RELAXED)
static inline void ls_cast(int kind, unsigned depth) {
    ls_kind = kind;
    ls_depth = (int)(depth < LS_DEPTHS ? depth : LS_DEPTHS - 1);
    ADD(casts, 1);
}
#define ORACLE_HOOK_SCENE_CAST(kind, depth) ls_cast((kind), (depth))
static void ls_tlas_leaf(const void* cx, uint32_t first, uint32_t count, const void* ray, const void* range);
#define ORACLE_HOOK_TLAS_LEAF(cx, first, count, ray, range) ls_tlas_leaf(cx, first, count, ray, range)
static void ls_blas_leaf(const void* cx, const void* tr, uint32_t first, uint32_t count, const void* ray, const void* range);
#define ORACLE_HOOK_BLAS_LEAF(cx, tr, first, count, ray, range) ls_blas_leaf(cx, tr, first, count, ray, range)

#include "../../oracle/oracle.c"

typedef struct { double lo[3], hi[3]; } Box;

static Box* g_inst_box;      /* per instance, world */
static Box* g_item_clip;     /* per tlas item position: instance box clipped to the leaf cell */
static Box* g_leaf_occ;      /* per tlas node (leaves only): union of the clipped boxes */
static Box* g_grp_a;         /* union of aligned runs of 8 global positions (scheme A) */
static uint32_t* g_leaf_of_first; /* node index by first item (lookup for the hook) */
static uint32_t* g_node_of_item0; /* tlas item position -> leaf node index (for positions that start a leaf) */
/* BLAS */
static Box* g_tri_box;       /* per triangle (global index), object space */
static Box* g_bitem_clip;    /* per blas item position (global): triangle box clipped to its leaf cell */
static Box* g_bleaf_occ;     /* per blas node (global index) */
static uint32_t* g_bnode_of_item0; /* global blas item position -> global node index, for leaf-first positions */

static Box box_empty(void) { Box b = {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; return b; }
static void box_grow(Box* b, const Box* o) {
    for (int c = 0; c < 3; ++c) { if (o->lo[c] < b->lo[c]) b->lo[c] = o->lo[c]; if (o->hi[c] > b->hi[c]) b->hi[c] = o->hi[c]; }
}
static Box box_clip(const Box* a, const Box* cell, double pad) {
    Box r;
    for (int c = 0; c < 3; ++c) {
        r.lo[c] = a->lo[c] > cell->lo[c] - pad ? a->lo[c] : cell->lo[c] - pad;
        r.hi[c] = a->hi[c] < cell->hi[c] + pad ? a->hi[c] : cell->hi[c] + pad;
    }
    return r;
}
static int box_is_empty(const Box* b) { return b->lo[0] > b->hi[0] || b->lo[1] > b->hi[1] || b->lo[2] > b->hi[2]; }

static int slab(const Box* b, const Ray* r, double s, double e) {
    if (box_is_empty(b)) return 0;
    double tn = -INFINITY, tf = INFINITY;
    const double o[3] = {r->origin.x, r->origin.y, r->origin.z}, d[3] = {r->direction.x, r->direction.y, r->direction.z};
    for (int c = 0; c < 3; ++c) {
        if (d[c] == 0.0) { if (o[c] < b->lo[c] || o[c] > b->hi[c]) return 0; continue; }
        double a = (b->lo[c] - o[c]) / d[c], bb = (b->hi[c] - o[c]) / d[c];
        if (a > bb) { double t = a; a = bb; bb = t; }
        if (a > tn) tn = a;
        if (bb < tf) tf = bb;
    }
    return !(tn > tf) && !(tf < s) && !(tn > e);
}

static void instance_world_box(const Scene* sc, uint32_t i, const Box* mesh_box, Box* out) {
    const PtInstance* in = &sc->instances[i];
    Box ob;
    if (in->prim == PT_PRIM_SPHERE) { for (int c = 0; c < 3; ++c) { ob.lo[c] = -1; ob.hi[c] = 1; } }
    else if (in->prim == PT_PRIM_TRIANGLE || in->prim == PT_PRIM_MESH || in->prim == PT_PRIM_KDMESH) ob = mesh_box[in->mesh];
    else { for (int c = 0; c < 3; ++c) { ob.lo[c] = -0.5; ob.hi[c] = 0.5; } if (in->prim == PT_PRIM_PLANE) ob.lo[1] = ob.hi[1] = 0.0; }
    for (int c = 0; c < 3; ++c) { ob.lo[c] -= 1e-4; ob.hi[c] += 1e-4; }
    const double* m = sc->instance_trans[i].trans;
    *out = box_empty();
    for (int corner = 0; corner < 8; ++corner) {
        V3 p = v3((corner & 1) ? ob.hi[0] : ob.lo[0], (corner & 2) ? ob.hi[1] : ob.lo[1], (corner & 4) ? ob.hi[2] : ob.lo[2]);
        V3 w = xf_point(m, p);
        Box pb = {{w.x, w.y, w.z}, {w.x, w.y, w.z}};
        box_grow(out, &pb);
    }
}

static void build_cells(const PtKdNode* nodes, uint32_t node, Box cell, void (*leaf_fn)(uint32_t node, const Box* cell, void* u), void* u) {
    const PtKdNode* n = &nodes[node];
    uint32_t axis = n->a & 3u;
    if (axis == 3u) { leaf_fn(node, &cell, u); return; }
    Box f = cell, b = cell;
    if (n->split > f.lo[axis]) f.lo[axis] = n->split;
    if (n->split < b.hi[axis]) b.hi[axis] = n->split;
    build_cells(nodes, n->a >> 2, f, leaf_fn, u);
    build_cells(nodes, n->b, b, leaf_fn, u);
}

static const Scene* g_sc;
static void tlas_leaf_fn(uint32_t node, const Box* cell, void* u) {
    (void)u;
    const PtKdNode* n = &g_sc->tlas_nodes[node];
    uint32_t first = n->a >> 2, count = n->b;
    Box occ = box_empty();
    for (uint32_t k = 0; k < count; ++k) {
        Box c = box_clip(&g_inst_box[g_sc->tlas_items[first + k]], cell, 1e-3);
        g_item_clip[first + k] = c;
        if (!box_is_empty(&c)) box_grow(&occ, &c);
    }
    g_leaf_occ[node] = occ;
    if (count) g_node_of_item0[first] = node;
}

typedef struct { const PtMesh* mesh; } BlasU;
static void blas_leaf_fn(uint32_t node, const Box* cell, void* u) {
    const PtMesh* mesh = ((BlasU*)u)->mesh;
    const PtKdNode* n = &g_sc->blas_nodes[mesh->node_first + node];
    uint32_t first = n->a >> 2, count = n->b;
    Box occ = box_empty();
    for (uint32_t k = 0; k < count; ++k) {
        uint32_t tri = mesh->tri_first + g_sc->blas_items[mesh->item_first + first + k];
        Box c = box_clip(&g_tri_box[tri], cell, 1e-6);
        g_bitem_clip[mesh->item_first + first + k] = c;
        if (!box_is_empty(&c)) box_grow(&occ, &c);
    }
    g_bleaf_occ[mesh->node_first + node] = occ;
    if (count) g_bnode_of_item0[mesh->item_first + first] = mesh->node_first + node;
}

static void ls_init(const Scene* sc) {
    g_sc = sc;
    const uint32_t nm = sc->h.n_meshes, ni = sc->h.n_instances, nt = sc->h.n_triangles;
    Box* mesh_box = calloc(nm ? nm : 1, sizeof(Box));
    g_tri_box = calloc(nt ? nt : 1, sizeof(Box));
    for (uint32_t t = 0; t < nt; ++t) {
        const double* v = (const double*)&sc->tri_pos[t];
        Box b = box_empty();
        for (int c = 0; c < 9; ++c) { if (v[c] < b.lo[c % 3]) b.lo[c % 3] = v[c]; if (v[c] > b.hi[c % 3]) b.hi[c % 3] = v[c]; }
        g_tri_box[t] = b;
    }
    for (uint32_t m = 0; m < nm; ++m) {
        mesh_box[m] = box_empty();
        for (uint32_t t = 0; t < sc->meshes[m].tri_count; ++t) box_grow(&mesh_box[m], &g_tri_box[sc->meshes[m].tri_first + t]);
    }
    g_inst_box = calloc(ni ? ni : 1, sizeof(Box));
    Box root = box_empty();
    for (uint32_t i = 0; i < ni; ++i) { instance_world_box(sc, i, mesh_box, &g_inst_box[i]); box_grow(&root, &g_inst_box[i]); }
    const uint32_t nitems = sc->h.n_tlas_items;
    g_item_clip = calloc(nitems ? nitems : 1, sizeof(Box));
    g_leaf_occ = calloc(sc->h.n_tlas_nodes, sizeof(Box));
    g_node_of_item0 = calloc(nitems ? nitems : 1, sizeof(uint32_t));
    build_cells(sc->tlas_nodes, 0, root, tlas_leaf_fn, NULL);
    const uint32_t ng = (nitems + 7) / 8;
    g_grp_a = calloc(ng ? ng : 1, sizeof(Box));
    for (uint32_t g = 0; g < ng; ++g) {
        g_grp_a[g] = box_empty();
        for (uint32_t k = g * 8; k < nitems && k < g * 8 + 8; ++k) box_grow(&g_grp_a[g], &g_inst_box[sc->tlas_items[k]]);
    }
    g_bitem_clip = calloc(sc->h.n_blas_items ? sc->h.n_blas_items : 1, sizeof(Box));
    g_bleaf_occ = calloc(sc->h.n_blas_nodes ? sc->h.n_blas_nodes : 1, sizeof(Box));
    g_bnode_of_item0 = calloc(sc->h.n_blas_items ? sc->h.n_blas_items : 1, sizeof(uint32_t));
    for (uint32_t m = 0; m < nm; ++m) {
        if (sc->meshes[m].kind != PT_MESH_KD) continue;
        BlasU u = {&sc->meshes[m]};
        build_cells(sc->blas_nodes + sc->meshes[m].node_first, 0, mesh_box[m], blas_leaf_fn, &u);
    }
    free(mesh_box);
    fprintf(stderr, "root box [%g %g %g] - [%g %g %g]\n", root.lo[0], root.lo[1], root.lo[2], root.hi[0], root.hi[1], root.hi[2]);
}

static void ls_tlas_leaf(const void* cxv, uint32_t first, uint32_t count, const void* rayv, const void* rangev) {
    (void)cxv;
    const Ray* ray = rayv;
    const Range* range = rangev;
    const double s = range->start, e = range->end;
    ADD(leaf_visits, 1);
    ADD(candidates, count);
    if (!count) return;
    const uint32_t end = first + count;
    /* A */
    uint64_t at = 0, ax = 0;
    for (uint32_t g = first >> 3; g <= (end - 1) >> 3; ++g) {
        uint32_t k0 = first > (g << 3) ? first : (g << 3), k1 = end < (g << 3) + 8 ? end : (g << 3) + 8;
        if (k1 - k0 > 2) { ++at; if (!slab(&g_grp_a[g], ray, s, e)) continue; }
        for (uint32_t k = k0; k < k1; ++k) { ++at; if (slab(&g_inst_box[g_sc->tlas_items[k]], ray, s, e)) ++ax; }
    }
    ADD(a_tests, at);
    ADD(a_exact, ax);
    /* B */
    uint64_t bt = 1, bx = 0, ct = 1;
    const uint32_t node = g_node_of_item0[first];
    if (!slab(&g_leaf_occ[node], ray, s, e)) { ADD(b_leaf_skipped, 1); ADD(b_tests, 1); ADD(c_tests, 1); ADD(empty_visits, 1); return; }
    for (uint32_t k0 = first; k0 < end; k0 += 8) {
        uint32_t k1 = k0 + 8 < end ? k0 + 8 : end;
        if (k1 - k0 > 2) {
            Box g = box_empty();
            for (uint32_t k = k0; k < k1; ++k) if (!box_is_empty(&g_item_clip[k])) box_grow(&g, &g_item_clip[k]);
            ++bt;
            if (!slab(&g, ray, s, e)) continue;
        }
        for (uint32_t k = k0; k < k1; ++k) { ++bt; if (slab(&g_item_clip[k], ray, s, e)) ++bx; }
    }
    /* C: 4-ary levels over leaf-relative positions: runs of 64 -> 16 -> 4 -> 1 */
    for (uint32_t k64 = first; k64 < end; k64 += 64) {
        uint32_t e64 = k64 + 64 < end ? k64 + 64 : end;
        if (count > 64) {
            Box g = box_empty();
            for (uint32_t k = k64; k < e64; ++k) if (!box_is_empty(&g_item_clip[k])) box_grow(&g, &g_item_clip[k]);
            ++ct;
            if (!slab(&g, ray, s, e)) continue;
        }
        for (uint32_t k16 = k64; k16 < e64; k16 += 16) {
            uint32_t e16 = k16 + 16 < e64 ? k16 + 16 : e64;
            Box g = box_empty();
            for (uint32_t k = k16; k < e16; ++k) if (!box_is_empty(&g_item_clip[k])) box_grow(&g, &g_item_clip[k]);
            ++ct;
            if (!slab(&g, ray, s, e)) continue;
            for (uint32_t k4 = k16; k4 < e16; k4 += 4) {
                uint32_t e4 = k4 + 4 < e16 ? k4 + 4 : e16;
                Box g4 = box_empty();
                for (uint32_t k = k4; k < e4; ++k) if (!box_is_empty(&g_item_clip[k])) box_grow(&g4, &g_item_clip[k]);
                ++ct;
                if (!slab(&g4, ray, s, e)) continue;
                ct += e4 - k4;
            }
        }
    }
    ADD(b_tests, bt);
    ADD(b_exact, bx);
    ADD(c_tests, ct);
    if (!bx) ADD(empty_visits, 1);
}

static void ls_blas_leaf(const void* cxv, const void* trv, uint32_t first, uint32_t count, const void* rayv, const void* rangev) {
    (void)cxv;
    const Tree* tr = trv;
    const Ray* ray = rayv;
    const Range* range = rangev;
    ADD(blas_leaf_visits, 1);
    ADD(blas_candidates, count);
    if (!count) return;
    const PtMesh* mesh = tr->mesh;
    uint64_t ax = 0, bx = 0;
    const uint32_t gfirst = mesh->item_first + first;
    for (uint32_t k = 0; k < count; ++k)
        if (slab(&g_tri_box[mesh->tri_first + tr->items[first + k]], ray, range->start, range->end)) ++ax;
    ADD(blas_a_exact, ax);
    if (!slab(&g_bleaf_occ[g_bnode_of_item0[gfirst]], ray, range->start, range->end)) { ADD(blas_b_skipped, 1); return; }
    for (uint32_t k = 0; k < count; ++k)
        if (slab(&g_bitem_clip[gfirst + k], ray, range->start, range->end)) ++bx;
    ADD(blas_b_exact, bx);
}

static void* read_file(const char* dir, const char* name, size_t* n) {
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(1); }
    fseek(f, 0, SEEK_END);
    *n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = malloc(*n ? *n : 1);
    if (fread(p, 1, *n, f) != *n) { perror("read"); exit(1); }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: leafstats <dump dir> [threads]\n"); return 2; }
    size_t nb, nc, np, ng;
    void* blob = read_file(argv[1], "blob.bin", &nb);
    PtCamera* cam = read_file(argv[1], "camera.bin", &nc);
    PtRenderParams* params = read_file(argv[1], "params.bin", &np);
    double* bg = read_file(argv[1], "background.bin", &ng);
    const int threads = argc > 2 ? atoi(argv[2]) : 8;
    Scene sc;
    if (scene_view(blob, nb, &sc)) { fprintf(stderr, "bad blob\n"); return 1; }
    ls_init(&sc);
    uint8_t* rgb = calloc((size_t)params->width * params->height, 3);
    OracleStats st;
    int rc = oracle_render(blob, nb, cam, params, bg, rgb, NULL, NULL, NULL, threads, &st);
    fprintf(stderr, "oracle rc %d\n", rc);
    static const char* kinds[LS_KINDS] = {"primary", "reflect", "refract", "shadow"};
    printf("%-8s %5s %10s %7s %7s | %8s %7s | %8s %7s %7s %7s | %7s || %7s %6s %6s %6s %6s\n", "kind", "depth", "casts", "leaf/c", "cand/c",
           "A tst/c", "A ex/c", "B tst/c", "B ex/c", "B skip", "empty", "C tst/c", "bl/c", "bcand", "A bex", "B bex", "Bskip");
    for (int k = 0; k < LS_KINDS; ++k)
        for (int d = 0; d < LS_DEPTHS; ++d) {
            const Cnt* c = &g_cnt[k][d];
            /* casts are not hooked per cast; derive from OracleStats for depth-less totals */
            if (!c->leaf_visits) continue;
            double n = (double)c->casts;
            if (n == 0) n = 1;
            printf("%-8s %5d %10llu %7.2f %7.1f | %8.1f %7.2f | %8.1f %7.2f %7.2f %7.2f | %7.1f || %7.2f %6.1f %6.2f %6.2f %6.2f\n", kinds[k], d,
                   (unsigned long long)c->casts, c->leaf_visits / n, c->candidates / n, c->a_tests / n, c->a_exact / n, c->b_tests / n,
                   c->b_exact / n, (double)c->b_leaf_skipped / c->leaf_visits, (double)c->empty_visits / c->leaf_visits, c->c_tests / n,
                   c->blas_leaf_visits / n, c->blas_candidates / n, c->blas_a_exact / n, c->blas_b_exact / n,
                   c->blas_leaf_visits ? (double)c->blas_b_skipped / c->blas_leaf_visits : 0.0);
        }
    printf("rays: primary %llu shadow %llu reflect %llu refract %llu; kd_splits %llu inst %llu tri %llu\n",
           (unsigned long long)st.rays_primary, (unsigned long long)st.rays_shadow, (unsigned long long)st.rays_reflect,
           (unsigned long long)st.rays_refract, (unsigned long long)st.kd_splits, (unsigned long long)st.instance_tests,
           (unsigned long long)st.triangle_tests);
    return 0;
}
