/*
 * oracle.c — CPU restatement of portrayer's per-pixel render loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (portrayer_b200/,
 * include/) links, imports or executes this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * do, and only as the checker / the CPU baseline.
 *
 * PARITY PINNING STATUS: the reference is pure Rust and cannot be compiled in
 * this image (no cargo/rustc, no vendored crates), so there is no oracle/_ref.
 * The oracle is pinned against every known-answer test the reference holds for
 * this path (tests/test_oracle_kats.py):
 *   src/math.rs:159-179            quadratic roots (count, order, 1e-6)
 *   src/kdtree/node.rs:219-352     two traversal KATs (winning material)
 *   src/kdtree/kdmesh.rs:99-166    Mesh == KDMesh, bit-exact colour, 100 000 rays on castle.obj
 * and against the reference's own published renders (tests/test_reference_renders.py,
 * the PNGs under render/ at SAMPLES=100 with OS randomness -> statistical, not bit, agreement).
 * Third-party arithmetic whose source is absent (vek 0.9.8 summation order,
 * roots 0.0.5 quadratic variant, libm) is restated from the published
 * algorithms: PARITY UNPINNED AT ULP LEVEL for those.
 *
 * Every function cites the reference file:line it follows.  Arithmetic is f64,
 * compiled with -ffp-contract=off (no FMA), sums left-to-right.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* analysis hooks: no-ops unless a tool that #includes this file defines them (tools/leafstats) */
#ifndef ORACLE_HOOK_SCENE_CAST
#define ORACLE_HOOK_SCENE_CAST(kind, depth) ((void)0)
#endif
#ifndef ORACLE_HOOK_TLAS_LEAF
#define ORACLE_HOOK_TLAS_LEAF(cx, first, count, ray, range) ((void)0)
#endif
#ifndef ORACLE_HOOK_BLAS_LEAF
#define ORACLE_HOOK_BLAS_LEAF(cx, tr, first, count, ray, range) ((void)0)
#endif

#define EPSILON PT_EPSILON /* src/math.rs:15 */
#define GAMMA PT_GAMMA     /* src/math.rs:20 */
#define PI 3.14159265358979323846264338327950288

typedef struct { double x, y, z; } V3;
typedef struct { double start, end; } Range; /* std::ops::Range<f64>, half-open */

static inline V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
static inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 scale(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 divs(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double magnitude(V3 a) { return sqrt(dot(a, a)); }
static inline V3 normalized(V3 a) { return divs(a, magnitude(a)); }
static inline int range_contains(const Range* r, double t) { return r->start <= t && t < r->end; }

/* Vec3Ext::transformed_point / transformed_direction, src/math.rs:44-52, on rows 0..2 of the Mat4 */
static inline V3 xf_point(const double* m, V3 p) {
    return v3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
static inline V3 xf_dir(const double* m, V3 d) {
    return v3(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z,
              m[8] * d.x + m[9] * d.y + m[10] * d.z);
}
/* direction through the TRANSPOSE of rows 0..2 (normal_trans = invtrans.transposed(), flat_scene.rs:105) */
static inline V3 xf_dir_transposed(const double* m, V3 d) {
    return v3(m[0] * d.x + m[4] * d.y + m[8] * d.z, m[1] * d.x + m[5] * d.y + m[9] * d.z,
              m[2] * d.x + m[6] * d.y + m[10] * d.z);
}

typedef struct { V3 origin, direction; } Ray; /* src/ray.rs:101-107 */
static inline V3 ray_at(const Ray* r, double t) { return add(r->origin, scale(r->direction, t)); } /* ray.rs:125-127 */
static inline Ray ray_transformed(const Ray* r, const double* m) { /* ray.rs:130-135 */
    Ray o = {xf_point(m, r->origin), xf_dir(m, r->direction)};
    return o;
}

/* src/ray.rs:10-36 plus the ids the hit-id buffer reports */
typedef struct {
    double ray_parameter;
    V3 hit_point;
    V3 normal;
    int has_tex_coord;
    double u, v;
    int has_nmt;
    double nmt[9]; /* normal_map_transform, row-major 3x3 */
    uint32_t sub;  /* triangle index / cube face / cylinder-cone part */
} Isect;

/* ------------------------------------------------------------------ scene view */
typedef struct {
    PtBlobHeader h;
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
    const PtTexture* textures;
    const uint8_t* texels;
} Scene;

static int scene_view(const void* blob, uint64_t bytes, Scene* s) {
    if (!blob || bytes < sizeof(PtBlobHeader)) return PT_ERR_INVALID;
    memcpy(&s->h, blob, sizeof s->h);
    if (s->h.magic != PT_BLOB_MAGIC || s->h.version != PT_BLOB_VERSION || s->h.total_bytes > bytes) return PT_ERR_INVALID;
    const unsigned char* b = (const unsigned char*)blob;
    s->tlas_nodes = (const PtKdNode*)(b + s->h.off_tlas_nodes);
    s->tlas_items = (const uint32_t*)(b + s->h.off_tlas_items);
    s->instances = (const PtInstance*)(b + s->h.off_instances);
    s->instance_trans = (const PtInstanceTrans*)(b + s->h.off_instance_trans);
    s->meshes = (const PtMesh*)(b + s->h.off_meshes);
    s->blas_nodes = (const PtKdNode*)(b + s->h.off_blas_nodes);
    s->blas_items = (const uint32_t*)(b + s->h.off_blas_items);
    s->tri_pos = (const PtTriPos*)(b + s->h.off_tri_pos);
    s->tri_normals = (const PtTriNormals*)(b + s->h.off_tri_normals);
    s->tri_uvs = (const PtTriUvs*)(b + s->h.off_tri_uvs);
    s->materials = (const PtMaterial*)(b + s->h.off_materials);
    s->lights = (const PtLight*)(b + s->h.off_lights);
    s->textures = (const PtTexture*)(b + s->h.off_textures);
    s->texels = b + s->h.off_texels;
    return PT_OK;
}

/* per-thread context: counters and the first "panic" */
typedef struct {
    const Scene* sc;
    uint32_t rng_mode;
    uint64_t seed;
    uint32_t max_depth;
    int linear_scene; /* PT_RENDER_LINEAR_TLAS: the scene root is the FlatScene itself (no k-d tree), see scene_ray_cast */
    OracleStats st;
    int error;
} Ctx;

/* ------------------------------------------------------------------ RNG (contract in portrayer_gpu.h) */
static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
static double draw(const Ctx* c, uint64_t pixel, uint64_t sample, uint64_t path, uint32_t dim) {
    if (c->rng_mode == PT_RNG_FIXED) return 0.5;
    uint64_t h = mix64(c->seed ^ (pixel * 0x9E3779B97F4A7C15ull + sample));
    h = mix64(h + ((path << 8) | dim));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0); /* rand 0.7 Standard f64: 53 bits * 2^-53 */
}

/* ------------------------------------------------------------------ roots / Quadratic */
/* roots 0.0.5 find_roots_quadratic (called from src/math.rs:107-114); roots ascending.
 * Restated from the crate's published algorithm (source absent offline). */
int oracle_solve_quadratic(double a2, double a1, double a0, double* out) {
    if (a2 == 0.0) { /* find_roots_linear */
        if (a1 == 0.0) {
            if (a0 == 0.0) { out[0] = 0.0; return 1; }
            return 0;
        }
        out[0] = -a0 / a1;
        return 1;
    }
    double discriminant = a1 * a1 - 4.0 * a2 * a0;
    if (discriminant < 0.0) return 0;
    double a2x2 = 2.0 * a2;
    if (discriminant == 0.0) { out[0] = -a1 / a2x2; return 1; }
    /* To improve precision, do not use the smallest divisor */
    double sq = sqrt(discriminant);
    double same_sign, diff_sign;
    if (a1 < 0.0) { same_sign = -a1 + sq; diff_sign = -a1 - sq; }
    else { same_sign = -a1 - sq; diff_sign = -a1 + sq; }
    double x1, x2;
    if (fabs(same_sign) > fabs(a2x2)) {
        double a0x2 = 2.0 * a0;
        if (fabs(diff_sign) > fabs(a2x2)) { x1 = a0x2 / same_sign; x2 = a0x2 / diff_sign; }
        else { x1 = a0x2 / same_sign; x2 = same_sign / a2x2; }
    } else { x1 = diff_sign / a2x2; x2 = same_sign / a2x2; }
    if (x1 < x2) { out[0] = x1; out[1] = x2; } else { out[0] = x2; out[1] = x1; }
    return 2;
}
/* Solutions::find_in_range, src/math.rs:94-96 */
static int quadratic_in_range(double a, double b, double c, const Range* r, double* t) {
    double roots[2];
    int n = oracle_solve_quadratic(a, b, c, roots);
    for (int i = 0; i < n; ++i)
        if (range_contains(r, roots[i])) { *t = roots[i]; return 1; }
    return 0;
}

/* ------------------------------------------------------------------ primitives */
static void mat3_from_cols(double* m, V3 c0, V3 c1, V3 c2) { /* vek Mat3::from_col_arrays */
    m[0] = c0.x; m[1] = c1.x; m[2] = c2.x;
    m[3] = c0.y; m[4] = c1.y; m[5] = c2.y;
    m[6] = c0.z; m[7] = c1.z; m[8] = c2.z;
}

/* InfinitePlane::ray_hit, src/primitive/infinite_plane.rs:46-79 */
static int infinite_plane_hit(V3 normal, V3 point, const Ray* ray, const Range* r, Isect* out) {
    double dot_dir_normal = dot(ray->direction, normal);
    double t = -dot(sub(ray->origin, point), normal) / dot_dir_normal;
    if (!range_contains(r, t)) return 0;
    out->ray_parameter = t;
    out->hit_point = ray_at(ray, t);
    out->normal = normal;
    out->has_tex_coord = 0;
    out->has_nmt = 0;
    out->sub = 0;
    return 1;
}

/* Sphere::ray_hit, src/primitive/sphere.rs:26-106 */
static int sphere_hit(const Ray* ray, const Range* r, Isect* out) {
    V3 origin = ray->origin, direction = ray->direction;
    double a = dot(direction, direction);
    double b = 2.0 * dot(origin, direction);
    double c = dot(origin, origin) - 1.0 * 1.0;
    double t;
    if (!quadratic_in_range(a, b, c, r, &t)) return 0;
    V3 hit_point = ray_at(ray, t);
    out->u = (PI + atan2(-hit_point.z, hit_point.x)) / (2.0 * PI);
    out->v = acos(hit_point.y) / PI;
    out->has_tex_coord = 1;
    V3 normal = hit_point;
    V3 to_top = normalized(sub(v3(0.0, 1.0, 0.0), hit_point));
    if (fabs(to_top.x) < EPSILON && fabs(to_top.z) < EPSILON) {
        mat3_from_cols(out->nmt, v3(1.0, 0.0, 0.0), normal, normal.y > 0.0 ? v3(0.0, 0.0, 1.0) : v3(0.0, 0.0, -1.0));
    } else {
        V3 horizontal_tangent = cross(to_top, normal);
        V3 vertical_tangent = cross(normal, horizontal_tangent);
        mat3_from_cols(out->nmt, horizontal_tangent, normal, vertical_tangent);
    }
    out->has_nmt = 1;
    out->ray_parameter = t;
    out->hit_point = hit_point;
    out->normal = normal;
    out->sub = 0;
    return 1;
}

/* Cube::contains, src/primitive/cube.rs:22-27 */
static int cube_contains(V3 p) {
    double radius = 0.5 + EPSILON;
    return -radius <= p.x && p.x <= radius && -radius <= p.y && p.y <= radius && -radius <= p.z && p.z <= radius;
}

/* FACES table, src/primitive/cube.rs:46-65: (plane point, plane normal, uv axis, uv offset) */
static const struct { V3 point, normal; double axis_u, axis_v, off_u, off_v; } CUBE_FACES[6] = {
    {{0.5, 0.0, 0.0}, {1.0, 0.0, 0.0}, -1.0, 1.0, 1.0 / 2.0, 1.0 / 3.0},   /* Right */
    {{-0.5, 0.0, 0.0}, {-1.0, 0.0, 0.0}, 1.0, 1.0, 0.0, 1.0 / 3.0},        /* Left */
    {{0.0, 0.5, 0.0}, {0.0, 1.0, 0.0}, 1.0, -1.0, 1.0 / 4.0, 0.0},         /* Top */
    {{0.0, -0.5, 0.0}, {0.0, -1.0, 0.0}, 1.0, 1.0, 1.0 / 4.0, 2.0 / 3.0},  /* Bottom */
    {{0.0, 0.0, 0.5}, {0.0, 0.0, 1.0}, 1.0, 1.0, 1.0 / 4.0, 1.0 / 3.0},    /* Near */
    {{0.0, 0.0, -0.5}, {0.0, 0.0, -1.0}, -1.0, 1.0, 3.0 / 4.0, 1.0 / 3.0}, /* Far */
};

/* Cube::ray_hit, src/primitive/cube.rs:38-141 */
static int cube_hit(const Ray* ray, const Range* init, Isect* out) {
    Range r = *init;
    int best = -1;
    Isect hit;
    for (int f = 0; f < 6; ++f) {
        Isect p_hit;
        if (infinite_plane_hit(CUBE_FACES[f].normal, CUBE_FACES[f].point, ray, &r, &p_hit)) {
            if (cube_contains(p_hit.hit_point)) {
                r.end = p_hit.ray_parameter;
                best = f;
                hit = p_hit;
            }
        }
    }
    if (best < 0) return 0;
    V3 n = CUBE_FACES[best].normal;
    V3 hit_p = hit.hit_point;
    double fu, fv;
    if (n.x != 0.0) { fu = hit_p.z; fv = hit_p.y; }
    else if (n.y != 0.0) { fu = hit_p.x; fv = hit_p.z; }
    else { fu = hit_p.x; fv = hit_p.y; }
    double norm_u = fu * CUBE_FACES[best].axis_u + 0.5;
    double norm_v = 0.5 - fv * CUBE_FACES[best].axis_v;
    hit.u = norm_u / 4.0 + CUBE_FACES[best].off_u;
    hit.v = norm_v / 3.0 + CUBE_FACES[best].off_v;
    hit.has_tex_coord = 1;
    V3 to_top = normalized(sub(v3(0.0, 1.0, 0.0), hit.hit_point));
    if (fabs(to_top.x) < EPSILON && fabs(to_top.z) < EPSILON) {
        mat3_from_cols(hit.nmt, v3(1.0, 0.0, 0.0), n, n.y > 0.0 ? v3(0.0, 0.0, 1.0) : v3(0.0, 0.0, -1.0));
    } else {
        V3 horizontal_tangent = cross(to_top, n);
        V3 vertical_tangent = cross(n, horizontal_tangent);
        mat3_from_cols(hit.nmt, horizontal_tangent, n, vertical_tangent);
    }
    hit.has_nmt = 1;
    hit.sub = (uint32_t)best;
    *out = hit;
    return 1;
}

/* Plane::ray_hit, src/primitive/plane.rs:26-53 */
static int plane_hit(const Ray* ray, const Range* r, Isect* out) {
    Isect hit;
    if (!infinite_plane_hit(v3(0.0, 1.0, 0.0), v3(0.0, 0.0, 0.0), ray, r, &hit)) return 0;
    double radius = 0.5 + EPSILON;
    V3 p = hit.hit_point;
    if (!(-radius <= p.x && p.x <= radius && -radius <= p.z && p.z <= radius)) return 0;
    hit.u = p.x + 0.5;
    hit.v = p.z + 0.5;
    hit.has_tex_coord = 1;
    mat3_from_cols(hit.nmt, v3(1.0, 0.0, 0.0), v3(0.0, 1.0, 0.0), v3(0.0, 0.0, 1.0));
    hit.has_nmt = 1;
    *out = hit;
    return 1;
}

/* cylinder ray_hit_body, src/primitive/cylinder.rs:28-74 */
static int cylinder_body(const Ray* ray, const Range* r, Isect* out) {
    V3 origin = ray->origin, direction = ray->direction;
    double a = direction.x * direction.x + direction.z * direction.z;
    double b = 2.0 * origin.x * direction.x + 2.0 * origin.z * direction.z;
    double c = origin.x * origin.x + origin.z * origin.z - 0.5 * 0.5;
    double t;
    if (!quadratic_in_range(a, b, c, r, &t)) return 0;
    V3 hit_point = ray_at(ray, t);
    if (hit_point.y > 0.5 || hit_point.y < -0.5) return 0;
    out->ray_parameter = t;
    out->hit_point = hit_point;
    out->normal = v3(hit_point.x, 0.0, hit_point.z);
    out->has_tex_coord = 0;
    out->has_nmt = 0;
    out->sub = 0;
    return 1;
}
/* cylinder ray_hit_cap, src/primitive/cylinder.rs:77-116 */
static int cylinder_cap(double height, const Ray* ray, const Range* r, Isect* out) {
    V3 origin = ray->origin, direction = ray->direction;
    double t = (height - origin.y) / direction.y;
    if (!range_contains(r, t)) return 0;
    V3 hit_point = ray_at(ray, t);
    if ((hit_point.x * hit_point.x + hit_point.z * hit_point.z) > 0.5 * 0.5) return 0;
    out->ray_parameter = t;
    out->hit_point = hit_point;
    out->normal = v3(0.0, height / fabs(height), 0.0);
    out->has_tex_coord = 0;
    out->has_nmt = 0;
    return 1;
}
/* Cylinder::ray_hit, src/primitive/cylinder.rs:118-153 */
static int cylinder_hit(const Ray* ray, const Range* init, Isect* out) {
    Range r = *init;
    int found = 0;
    Isect hit;
    if (cylinder_body(ray, &r, &hit)) { r.end = hit.ray_parameter; *out = hit; out->sub = 0; found = 1; }
    if (cylinder_cap(0.5, ray, &r, &hit)) { r.end = hit.ray_parameter; *out = hit; out->sub = 1; found = 1; }
    if (cylinder_cap(-0.5, ray, &r, &hit)) { *out = hit; out->sub = 2; found = 1; }
    return found;
}

/* cone ray_hit_body, src/primitive/cone.rs:28-113 */
static int cone_body(const Ray* ray, const Range* r, Isect* out) {
    V3 origin = ray->origin, direction = ray->direction;
    const double HEIGHT = 1.0, RADIUS = 0.5;
    double h_sqr = HEIGHT * HEIGHT;
    double r_sqr = RADIUS * RADIUS;
    double a = 4.0 * direction.y * direction.y * r_sqr - 4.0 * h_sqr * (direction.x * direction.x + direction.z * direction.z);
    double b = -8.0 * h_sqr * (direction.x * origin.x + direction.z * origin.z) -
               4.0 * r_sqr * (direction.y * HEIGHT - 2.0 * direction.y * origin.y);
    double c = -4.0 * h_sqr * (origin.x * origin.x + origin.z * origin.z) +
               r_sqr * (h_sqr - 4.0 * HEIGHT * origin.y + 4.0 * origin.y * origin.y);
    double t;
    if (!quadratic_in_range(a, b, c, r, &t)) return 0;
    V3 hit_point = ray_at(ray, t);
    if (hit_point.y > 0.5 || hit_point.y < -0.5) return 0;
    V3 tip = v3(0.0, 0.5, 0.0);
    V3 tangent1 = sub(tip, hit_point);
    V3 opposite = v3(-hit_point.x, hit_point.y, -hit_point.z);
    V3 across = sub(opposite, hit_point);
    V3 tangent2 = cross(tangent1, across);
    V3 normal = cross(tangent1, tangent2);
    out->ray_parameter = t;
    out->hit_point = hit_point;
    out->normal = normal;
    out->has_tex_coord = 0;
    out->has_nmt = 0;
    out->sub = 0;
    return 1;
}
/* cone ray_hit_cap, src/primitive/cone.rs:116-157 */
static int cone_cap(const Ray* ray, const Range* r, Isect* out) {
    V3 origin = ray->origin, direction = ray->direction;
    double height = -0.5;
    double t = (height - origin.y) / direction.y;
    if (!range_contains(r, t)) return 0;
    V3 hit_point = ray_at(ray, t);
    if ((hit_point.x * hit_point.x + hit_point.z * hit_point.z) > 0.5 * 0.5) return 0;
    out->ray_parameter = t;
    out->hit_point = hit_point;
    out->normal = v3(0.0, -1.0, 0.0);
    out->has_tex_coord = 0;
    out->has_nmt = 0;
    out->sub = 1;
    return 1;
}
/* Cone::ray_hit, src/primitive/cone.rs:159-186 */
static int cone_hit(const Ray* ray, const Range* init, Isect* out) {
    Range r = *init;
    int found = 0;
    Isect hit;
    if (cone_body(ray, &r, &hit)) { r.end = hit.ray_parameter; *out = hit; found = 1; }
    if (cone_cap(ray, &r, &hit)) { *out = hit; found = 1; }
    return found;
}

/* Triangle::ray_hit, src/primitive/triangle.rs:38-147 */
static int triangle_hit(Ctx* cx, const PtMesh* mesh, uint32_t tri, const Ray* ray, const Range* range, Isect* out) {
    cx->st.triangle_tests++;
    const PtTriPos* tp = &cx->sc->tri_pos[mesh->tri_first + tri];
    V3 A = v3(tp->a[0], tp->a[1], tp->a[2]), B = v3(tp->b[0], tp->b[1], tp->b[2]), C = v3(tp->c[0], tp->c[1], tp->c[2]);
    V3 ab = sub(A, B), ac = sub(A, C), dir = ray->direction, ao = sub(A, ray->origin);
    double a = ab.x, b = ab.y, c = ab.z;
    double d = ac.x, e = ac.y, f = ac.z;
    double g = dir.x, h = dir.y, i = dir.z;
    double j = ao.x, k = ao.y, l = ao.z;

    double ei_hf = e * i - h * f;
    double gf_di = g * f - d * i;
    double dh_eg = d * h - e * g;
    double m = a * ei_hf + b * gf_di + c * dh_eg;

    double ak_jb = a * k - j * b;
    double jc_al = j * c - a * l;
    double bl_ck = b * l - c * k;

    double t = -(f * ak_jb + e * jc_al + d * bl_ck) / m;
    if (!range_contains(range, t)) return 0;
    double gamma = (i * ak_jb + h * jc_al + g * bl_ck) / m;
    if (gamma < 0.0 || gamma > 1.0) return 0;
    double beta = (j * ei_hf + k * gf_di + l * dh_eg) / m;
    if (beta < 0.0 || beta > 1.0 - gamma) return 0;

    V3 normal;
    if (mesh->flags & PT_MESH_FLAG_NORMALS) {
        const PtTriNormals* tn = &cx->sc->tri_normals[mesh->nrm_first + tri];
        double alpha = 1.0 - beta - gamma;
        V3 na = v3(tn->na[0], tn->na[1], tn->na[2]), nb = v3(tn->nb[0], tn->nb[1], tn->nb[2]), nc = v3(tn->nc[0], tn->nc[1], tn->nc[2]);
        normal = add(add(scale(na, alpha), scale(nb, beta)), scale(nc, gamma));
    } else {
        normal = cross(sub(B, A), sub(C, A));
    }
    out->has_tex_coord = 0;
    out->has_nmt = 0;
    if (mesh->flags & PT_MESH_FLAG_UVS) {
        const PtTriUvs* tu = &cx->sc->tri_uvs[mesh->uv_first + tri];
        double alpha = 1.0 - beta - gamma;
        double uu = tu->uva[0] * alpha + tu->uvb[0] * beta + tu->uvc[0] * gamma;
        double vv = tu->uva[1] * alpha + tu->uvb[1] * beta + tu->uvc[1] * gamma;
        out->u = uu;
        out->v = 1.0 - vv;
        out->has_tex_coord = 1;

        V3 edge1 = sub(B, A), edge2 = sub(C, A);
        double d1u = tu->uvb[0] - tu->uva[0], d1v = tu->uvb[1] - tu->uva[1];
        double d2u = tu->uvc[0] - tu->uva[0], d2v = tu->uvc[1] - tu->uva[1];
        V3 tangent = v3(d2v * edge1.x - d1v * edge2.x, d2v * edge1.y - d1v * edge2.y, d2v * edge1.z - d1v * edge2.z);
        V3 bitangent = v3(-d2u * edge1.x + d1u * edge2.x, -d2u * edge1.y + d1u * edge2.y, -d2u * edge1.z + d1u * edge2.z);
        double coeff = d1u * d2v - d2u * d1v;
        tangent = normalized(divs(tangent, coeff));
        bitangent = normalized(divs(bitangent, coeff));
        V3 nn = normalized(normal);
        mat3_from_cols(out->nmt, tangent, nn, bitangent);
        out->has_nmt = 1;
    }
    out->ray_parameter = t;
    out->hit_point = ray_at(ray, t);
    out->normal = normal;
    out->sub = tri;
    return 1;
}

/* BoundingBox::test_hit, src/bounding_box.rs:104-116 */
static int bbox_test_hit(Ctx* cx, const PtMesh* mesh, const Ray* ray, const Range* r) {
    cx->st.bbox_gates++;
    Ray local_ray = ray_transformed(ray, mesh->bbox_invtrans);
    if (cube_contains(ray_at(&local_ray, r->start))) return 1;
    Isect tmp;
    return cube_hit(&local_ray, r, &tmp);
}

/* Mesh::ray_hit, src/primitive/mesh.rs:145-168 */
static int mesh_hit(Ctx* cx, const PtMesh* mesh, const Ray* ray, const Range* init, Isect* out) {
    if (!bbox_test_hit(cx, mesh, ray, init)) return 0;
    Range r = *init;
    int found = 0;
    for (uint32_t t = 0; t < mesh->tri_count; ++t) {
        Isect hit;
        if (triangle_hit(cx, mesh, t, ray, &r, &hit)) {
            r.end = hit.ray_parameter;
            *out = hit;
            found = 1;
        }
    }
    return found;
}

/* ------------------------------------------------------------------ kd traversal */
static int instance_ray_cast(Ctx* cx, uint32_t inst, const Ray* ray, Range* range, Isect* out);

/* what a leaf does with its item list: RayCast for [T] (ray.rs:87-99) over
 * instances, or RayHit for [T] (ray.rs:50-63) over triangles, both folded with
 * a shrinking range.  On Some the range end is the hit's t. */
typedef struct {
    int is_blas;
    const PtMesh* mesh; /* blas */
    const PtKdNode* nodes;
    const uint32_t* items;
} Tree;

static int leaf_cast(Ctx* cx, const Tree* tr, uint32_t first, uint32_t count, const Ray* ray, Range* range, Isect* out,
                     uint32_t* out_inst) {
    int found = 0;
    if (tr->is_blas) {
        ORACLE_HOOK_BLAS_LEAF(cx, tr, first, count, ray, range);
        Range r = *range; /* [T]::ray_hit clones the range, ray.rs:52 */
        for (uint32_t k = 0; k < count; ++k) {
            Isect hit;
            if (triangle_hit(cx, tr->mesh, tr->items[first + k], ray, &r, &hit)) {
                r.end = hit.ray_parameter;
                *out = hit;
                found = 1;
            }
        }
        if (found) range->end = out->ray_parameter; /* node.rs:40-46 */
    } else {
        ORACLE_HOOK_TLAS_LEAF(cx, first, count, ray, range);
        for (uint32_t k = 0; k < count; ++k) {
            Isect hit;
            uint32_t inst = tr->items[first + k];
            if (instance_ray_cast(cx, inst, ray, range, &hit)) {
                range->end = hit.ray_parameter; /* ray.rs:93 (already set by the instance, flat_scene.rs:92) */
                *out = hit;
                *out_inst = inst;
                found = 1;
            }
        }
    }
    return found;
}

/* KDTreeNode::ray_cast_impl, src/kdtree/node.rs:66-203 (recursive, like the reference) */
static int kd_cast(Ctx* cx, const Tree* tr, uint32_t node, const Ray* ray, Range* t_range, double extent, Isect* out,
                   uint32_t* out_inst) {
    const PtKdNode* n = &tr->nodes[node];
    uint32_t axis = n->a & 3u;
    if (axis == 3u) return leaf_cast(cx, tr, n->a >> 2, n->b, ray, t_range, out, out_inst);
    cx->st.kd_splits++;
    uint32_t front = n->a >> 2, back = n->b;
    /* node.rs:119-127 */
    double t_max = t_range->start + extent;
    t_max = range_contains(t_range, t_max) ? t_max : t_range->end - EPSILON;
    double t_min = t_range->start + EPSILON;
    V3 ray_start = ray_at(ray, t_min);
    V3 ray_end = ray_at(ray, t_max);
    /* which_side with an axis-unit normal and a point on that axis only:
     * (p - point).dot(normal) >= 0, infinite_plane.rs:27-35 */
    V3 normal = v3(axis == 0 ? 1.0 : 0.0, axis == 1 ? 1.0 : 0.0, axis == 2 ? 1.0 : 0.0);
    V3 point = scale(normal, n->split);
    int start_front = dot(sub(ray_start, point), normal) >= 0.0;
    int end_front = dot(sub(ray_end, point), normal) >= 0.0;
    if (start_front == end_front)
        return kd_cast(cx, tr, start_front ? front : back, ray, t_range, extent, out, out_inst); /* node.rs:134-137 */

    /* ray_hit_axis_aligned_plane, node.rs:90-110 */
    V3 np = mul(normal, point), no = mul(normal, ray->origin), nd = mul(normal, ray->direction);
    double plane_value = np.x + np.y + np.z; /* (a * b).sum() */
    double ray_origin = no.x + no.y + no.z;
    double ray_direction = nd.x + nd.y + nd.z;
    double plane_t = (plane_value - ray_origin) / ray_direction;
    if (!range_contains(t_range, plane_t)) {
        cx->error = PT_ERR_KD_PLANE_MISS; /* .expect("bug: ray should definitely hit infinite plane") */
        return 0;
    }
    uint32_t near = start_front ? front : back, far = start_front ? back : front;
    Range near_range = {t_range->start, plane_t};
    if (kd_cast(cx, tr, near, ray, &near_range, extent, out, out_inst)) { /* node.rs:150-155 */
        *t_range = near_range;
        return 1;
    }
    Range far_range = {plane_t, t_range->end};
    if (kd_cast(cx, tr, far, ray, &far_range, extent, out, out_inst)) { /* node.rs:158-163 */
        *t_range = far_range;
        return 1;
    }
    return 0;
}

/* KDMesh::ray_hit, src/kdtree/kdmesh.rs:62-74 -> KDTreeNode::ray_hit, node.rs:33-51 */
static int kdmesh_hit(Ctx* cx, const PtMesh* mesh, const Ray* ray, const Range* range, Isect* out) {
    if (!bbox_test_hit(cx, mesh, ray, range)) return 0;
    Tree tr = {1, mesh, cx->sc->blas_nodes + mesh->node_first, cx->sc->blas_items + mesh->item_first};
    Range r = *range; /* node.rs:38 */
    uint32_t unused = 0;
    return kd_cast(cx, &tr, 0, ray, &r, mesh->extent, out, &unused);
}

/* Primitive::ray_hit dispatch, src/primitive.rs:55-61 */
static int primitive_hit(Ctx* cx, const PtInstance* in, const Ray* ray, const Range* r, Isect* out) {
    switch (in->prim) {
        case PT_PRIM_SPHERE: return sphere_hit(ray, r, out);
        case PT_PRIM_CUBE: return cube_hit(ray, r, out);
        case PT_PRIM_PLANE: return plane_hit(ray, r, out);
        case PT_PRIM_CYLINDER: return cylinder_hit(ray, r, out);
        case PT_PRIM_CONE: return cone_hit(ray, r, out);
        case PT_PRIM_TRIANGLE: return triangle_hit(cx, &cx->sc->meshes[in->mesh], 0, ray, r, out);
        case PT_PRIM_MESH: return mesh_hit(cx, &cx->sc->meshes[in->mesh], ray, r, out);
        case PT_PRIM_KDMESH: return kdmesh_hit(cx, &cx->sc->meshes[in->mesh], ray, r, out);
        default: return 0;
    }
}

/* FlatSceneNode::ray_cast, src/flat_scene.rs:71-99 */
static int instance_ray_cast(Ctx* cx, uint32_t inst, const Ray* ray, Range* range, Isect* out) {
    cx->st.instance_tests++;
    const PtInstance* in = &cx->sc->instances[inst];
    Ray local_ray = ray_transformed(ray, in->invtrans);
    Isect hit;
    if (!primitive_hit(cx, in, &local_ray, range, &hit)) return 0;
    hit.hit_point = xf_point(cx->sc->instance_trans[inst].trans, hit.hit_point);
    hit.normal = xf_dir_transposed(in->invtrans, hit.normal);
    range->end = hit.ray_parameter;
    *out = hit;
    return 1;
}

/* scene.root.ray_cast: KDTreeNode<FlatSceneNode>::ray_cast, node.rs:27-31.
 * Cross-check mode: Scene<R> is generic over its RayCast root (scene.rs, ray.rs:139-148); with the
 * FlatScene as the root the cast is `[FlatSceneNode]::ray_cast`, ray.rs:87-99 over flat_scene.rs:71-99 — every
 * instance in list order sharing one shrinking range, no tree.  Same nearest hit as the tree walk wherever the tree
 * loses nothing (the walk's probe / EPSILON quirks can), first-listed instance wins an exact tie in both. */
static int scene_ray_cast(Ctx* cx, const Ray* ray, Range* range, Isect* out, uint32_t* out_inst) {
    if (cx->linear_scene) {
        int found = 0;
        for (uint32_t inst = 0; inst < cx->sc->h.n_instances; ++inst) {
            Isect hit;
            if (instance_ray_cast(cx, inst, ray, range, &hit)) { /* sets range->end, ray.rs:91-95 */
                *out = hit;
                *out_inst = inst;
                found = 1;
            }
        }
        return found;
    }
    Tree tr = {0, NULL, cx->sc->tlas_nodes, cx->sc->tlas_items};
    return kd_cast(cx, &tr, 0, ray, range, cx->sc->h.tlas_extent, out, out_inst);
}

/* ------------------------------------------------------------------ textures */
/* Rust `f64 as i64`: truncate toward zero, saturate, NaN -> 0 */
static int64_t f64_as_i64(double v) {
    if (v != v) return 0;
    if (v >= 9223372036854775807.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)v;
}
static int64_t rem_euclid(int64_t value, int64_t rhs) { /* texture.rs:107-119 */
    int64_t r = value % rhs;
    if (r < 0) return rhs < 0 ? r - rhs : r + rhs;
    return r;
}
/* RgbImageBuffer::at, src/texture.rs:104-141 */
static void texture_at(Ctx* cx, int32_t tex, double u, double v, double* rgb) {
    cx->st.texel_lookups++;
    const PtTexture* t = &cx->sc->textures[tex];
    int64_t width = t->width, height = t->height;
    int64_t x = f64_as_i64(u * (double)(width - 1));
    int64_t y = f64_as_i64(v * (double)(height - 1));
    uint32_t xi = (uint32_t)rem_euclid(x, width);
    uint32_t yi = (uint32_t)rem_euclid(y, height);
    const uint8_t* p = cx->sc->texels + t->offset + ((uint64_t)yi * t->width + xi) * 3;
    rgb[0] = (double)p[0] / 255.0;
    rgb[1] = (double)p[1] / 255.0;
    rgb[2] = (double)p[2] / 255.0;
}

/* ------------------------------------------------------------------ shading */
typedef struct { uint64_t pixel, sample; } PathKey;

static void ray_color(Ctx* cx, const Ray* ray, const double* background, uint32_t depth, PathKey key, uint64_t path,
                      int ray_kind, double* out_rgb, uint32_t* hit_id, double* hit_t);

/* refracted_direction, src/material.rs:27-48 */
static int refracted_direction(V3 ray_dir, V3 normal, double refraction_index, V3* out) {
    double eta = refraction_index;
    double eta_outside = 1.00;
    double ray_dot_norm = dot(ray_dir, normal);
    double under_sqrt = 1.0 - eta_outside * eta_outside * (1.0 - ray_dot_norm * ray_dot_norm) / (eta * eta);
    if (under_sqrt < 0.0) return 0;
    V3 refracted_dir_1 = divs(scale(sub(ray_dir, scale(normal, ray_dot_norm)), eta_outside), eta);
    V3 refracted_dir_2 = scale(normal, sqrt(under_sqrt));
    *out = sub(refracted_dir_1, refracted_dir_2);
    return 1;
}

/* Material::hit_color, src/material.rs:91-320 */
static void hit_color(Ctx* cx, const PtMaterial* mat, const double* background, V3 ray_dir, const Isect* hit,
                      uint32_t depth, PathKey key, uint64_t path, double* out) {
    if (depth > cx->max_depth) { /* material.rs:102-104 */
        out[0] = background[0]; out[1] = background[1]; out[2] = background[2];
        return;
    }
    cx->st.shaded_hits++;
    const Scene* sc = cx->sc;
    V3 view = neg(ray_dir);
    V3 hit_point = hit->hit_point;

    /* uv' = uv_trans * (u, v, 1), material.rs:114-117 */
    double tu = 0.0, tv = 0.0;
    if (hit->has_tex_coord) {
        const double* m = mat->uv_trans;
        tu = m[0] * hit->u + m[1] * hit->v + m[2] * 1.0;
        tv = m[3] * hit->u + m[4] * hit->v + m[5] * 1.0;
    }

    V3 normal;
    if (mat->normals < 0) {
        normal = normalized(hit->normal);
    } else if (hit->has_tex_coord && hit->has_nmt) {
        /* NormalMap::normal_at, texture.rs:192-221 */
        double c[3];
        texture_at(cx, mat->normals, tu, tv, c);
        V3 norm = v3(2.0 * c[0] - 1.0, 2.0 * c[1] - 1.0, -(2.0 * c[2] - 1.0));
        V3 tex_norm = v3(1.0 * norm.x + 0.0 * norm.y + 0.0 * norm.z, 0.0 * norm.x + 0.0 * norm.y + -1.0 * norm.z,
                         0.0 * norm.x + -1.0 * norm.y + 0.0 * norm.z);
        V3 tn = normalized(tex_norm);
        const double* m = hit->nmt;
        normal = v3(m[0] * tn.x + m[1] * tn.y + m[2] * tn.z, m[3] * tn.x + m[4] * tn.y + m[5] * tn.z,
                    m[6] * tn.x + m[7] * tn.y + m[8] * tn.z);
    } else {
        cx->error = PT_ERR_NO_TEXCOORD_NORMALMAP; /* material.rs:133 */
        out[0] = out[1] = out[2] = 0.0;
        return;
    }

    double kd[3];
    if (mat->texture < 0) {
        kd[0] = mat->diffuse[0]; kd[1] = mat->diffuse[1]; kd[2] = mat->diffuse[2];
    } else if (hit->has_tex_coord) {
        /* ImageTexture::at, texture.rs:162-168 */
        texture_at(cx, mat->texture, tu, tv, kd);
        kd[0] = pow(kd[0], GAMMA); kd[1] = pow(kd[1], GAMMA); kd[2] = pow(kd[2], GAMMA);
    } else {
        cx->error = PT_ERR_NO_TEXCOORD_TEXTURE; /* material.rs:141 */
        out[0] = out[1] = out[2] = 0.0;
        return;
    }

    double color[3] = {sc->h.ambient[0] * kd[0], sc->h.ambient[1] * kd[1], sc->h.ambient[2] * kd[2]}; /* material.rs:148 */
    const uint32_t n_lights = sc->h.n_lights;
    for (uint32_t l = 0; l < n_lights; ++l) { /* material.rs:149-212 */
        const PtLight* light = &sc->lights[l];
        V3 lpos = v3(light->position[0], light->position[1], light->position[2]);
        V3 area_a = v3(light->area_a[0], light->area_a[1], light->area_a[2]);
        V3 area_b = v3(light->area_b[0], light->area_b[1], light->area_b[2]);
        int empty = (area_a.x == 0.0 && area_a.y == 0.0 && area_a.z == 0.0) ||
                    (area_b.x == 0.0 && area_b.y == 0.0 && area_b.z == 0.0); /* light.rs:51-53 */
        V3 light_pos = lpos;
        if (!empty) { /* light.rs:62-70, 88-90 */
            double a_coord = 2.0 * draw(cx, key.pixel, key.sample, path, 2 + 2 * l) - 1.0;
            double b_coord = 2.0 * draw(cx, key.pixel, key.sample, path, 3 + 2 * l) - 1.0;
            light_pos = add(lpos, add(scale(area_a, a_coord), scale(area_b, b_coord)));
        }
        V3 hit_to_light = sub(light_pos, hit_point);
        double light_dist = magnitude(hit_to_light);
        V3 light_dir = divs(hit_to_light, light_dist);
        double attenuation = light->falloff[0] + light->falloff[1] * light_dist + light->falloff[2] * light_dist * light_dist;

        Ray shadow_ray = {hit_point, light_dir};
        Range shadow_range = {EPSILON, INFINITY};
        Isect sh;
        uint32_t sh_inst;
        cx->st.rays_shadow++;
        const OracleStats before = cx->st;
        ORACLE_HOOK_SCENE_CAST(3, depth);
        const int shadowed = scene_ray_cast(cx, &shadow_ray, &shadow_range, &sh, &sh_inst); /* material.rs:174-179 */
        cx->st.shadow_kd_splits += cx->st.kd_splits - before.kd_splits;
        cx->st.shadow_instance_tests += cx->st.instance_tests - before.instance_tests;
        cx->st.shadow_triangle_tests += cx->st.triangle_tests - before.triangle_tests;
        cx->st.shadow_bbox_gates += cx->st.bbox_gates - before.bbox_gates;
        if (!shadowed) {
            double normal_light = fmax(dot(normal, light_dir), 0.0);
            double diffuse[3] = {kd[0] * light->color[0] * normal_light, kd[1] * light->color[1] * normal_light,
                                 kd[2] * light->color[2] * normal_light};
            double specular[3] = {0.0, 0.0, 0.0};
            if (mat->specular[0] > EPSILON || mat->specular[1] > EPSILON || mat->specular[2] > EPSILON) {
                V3 half = normalized(add(view, light_dir));
                double normal_half_shiny = pow(fmax(dot(normal, half), 0.0), 4.0 * mat->shininess);
                specular[0] = mat->specular[0] * light->color[0] * normal_half_shiny;
                specular[1] = mat->specular[1] * light->color[1] * normal_half_shiny;
                specular[2] = mat->specular[2] * light->color[2] * normal_half_shiny;
            }
            color[0] += (diffuse[0] + specular[0]) / attenuation;
            color[1] += (diffuse[1] + specular[1]) / attenuation;
            color[2] += (diffuse[2] + specular[2]) / attenuation;
        }
    }

    if (mat->reflectivity > 0.0) { /* material.rs:216-317 */
        V3 reflect_dir = sub(ray_dir, scale(scale(normal, 2.0), dot(ray_dir, normal)));
        if (mat->glossy_side_length > 0.0) { /* material.rs:220-239 */
            V3 offset_vector = (fabs(reflect_dir.x) < EPSILON && fabs(reflect_dir.y) < EPSILON)
                                   ? add(reflect_dir, v3(0.0, 0.1, 0.0))
                                   : add(reflect_dir, v3(0.0, 0.0, 0.1));
            V3 u_basis = cross(reflect_dir, offset_vector);
            V3 v_basis = cross(reflect_dir, u_basis);
            double g = mat->glossy_side_length;
            double u_coord = -g / 2.0 + draw(cx, key.pixel, key.sample, path, 2 + 2 * n_lights) * g;
            double v_coord = -g / 2.0 + draw(cx, key.pixel, key.sample, path, 3 + 2 * n_lights) * g;
            reflect_dir = add(reflect_dir, add(scale(u_basis, u_coord), scale(v_basis, v_coord)));
        }
        Ray reflected_ray = {hit_point, reflect_dir};
        double reflected_color[3];
        ray_color(cx, &reflected_ray, background, depth + 1, key, path << 1, 1, reflected_color, NULL, NULL);

        if (mat->refraction_index > 0.0) {
            V3 refract_dir;
            double cos_incident = 0.0;
            int have = 0;
            if (dot(ray_dir, normal) < 0.0) {
                if (!refracted_direction(ray_dir, normal, mat->refraction_index, &refract_dir)) {
                    cx->error = PT_ERR_TIR_INSIDE; /* material.rs:258 */
                } else {
                    cos_incident = dot(neg(ray_dir), normal);
                    have = 1;
                }
            } else if (refracted_direction(ray_dir, neg(normal), 1.0 / mat->refraction_index, &refract_dir)) {
                cos_incident = dot(refract_dir, normal);
                have = 1;
            } else {
                /* Total internal reflection, material.rs:277-284 */
                color[0] += mat->reflectivity * reflected_color[0];
                color[1] += mat->reflectivity * reflected_color[1];
                color[2] += mat->reflectivity * reflected_color[2];
            }
            if (have) {
                double r0 = (mat->refraction_index - 1.0) * (mat->refraction_index - 1.0);
                r0 = r0 / ((mat->refraction_index + 1.0) * (mat->refraction_index + 1.0));
                double base = 1.0 - cos_incident;
                /* powi(5): llvm.powi -> compiler-rt __powidf2 (square-and-multiply) */
                double p5;
                {
                    double acc = base, r = base; /* n = 5 = 0b101 */
                    acc = acc * acc;             /* ^2 */
                    acc = acc * acc;             /* ^4 */
                    r = r * acc;                 /* ^5 */
                    p5 = r;
                }
                double reflectivity = r0 + (1.0 - r0) * p5;
                double transmittance = 1.0 - reflectivity;
                Ray refracted_ray = {hit_point, refract_dir};
                double refracted_color[3];
                ray_color(cx, &refracted_ray, background, depth + 1, key, (path << 1) | 1, 2, refracted_color, NULL, NULL);
                for (int ch = 0; ch < 3; ++ch) {
                    double total_color = reflectivity * reflected_color[ch] + transmittance * refracted_color[ch];
                    color[ch] += mat->reflectivity * total_color;
                }
            }
        } else {
            color[0] += mat->reflectivity * reflected_color[0];
            color[1] += mat->reflectivity * reflected_color[1];
            color[2] += mat->reflectivity * reflected_color[2];
        }
    }
    out[0] = color[0]; out[1] = color[1]; out[2] = color[2];
}

/* Ray::color, src/ray.rs:139-148. ray_kind: 0 primary, 1 reflected, 2 refracted (stats only) */
static void ray_color(Ctx* cx, const Ray* ray, const double* background, uint32_t depth, PathKey key, uint64_t path,
                      int ray_kind, double* out_rgb, uint32_t* hit_id, double* hit_t) {
    if (depth > cx->max_depth) {
        /* The reference casts this ray (ray.rs:141) but the result is bg either
         * way (material.rs:102-104 / ray.rs:146). Counted, not traced. */
        cx->st.rays_depth_cut++;
        out_rgb[0] = background[0]; out_rgb[1] = background[1]; out_rgb[2] = background[2];
        return;
    }
    if (ray_kind == 0) cx->st.rays_primary++;
    else if (ray_kind == 1) cx->st.rays_reflect++;
    else cx->st.rays_refract++;
    Range t_range = {EPSILON, INFINITY};
    Isect hit;
    uint32_t inst = 0xFFFFFFFFu;
    ORACLE_HOOK_SCENE_CAST(ray_kind, depth);
    if (scene_ray_cast(cx, ray, &t_range, &hit, &inst)) {
        if (hit_id) { hit_id[0] = inst; hit_id[1] = hit.sub; }
        if (hit_t) *hit_t = hit.ray_parameter;
        const PtMaterial* mat = &cx->sc->materials[cx->sc->instances[inst].material];
        hit_color(cx, mat, background, ray->direction, &hit, depth, key, path, out_rgb);
    } else {
        if (hit_id) { hit_id[0] = 0xFFFFFFFFu; hit_id[1] = 0; }
        if (hit_t) *hit_t = INFINITY;
        out_rgb[0] = background[0]; out_rgb[1] = background[1]; out_rgb[2] = background[2];
    }
}

/* Camera::ray_at, src/camera.rs:48-84 */
static Ray camera_ray_at(const PtCamera* cam, double x, double y) {
    double pixel_ndc_y = y / cam->height;
    double pixel_view_y = (1.0 - 2.0 * pixel_ndc_y) * cam->fov_factor;
    double pixel_ndc_x = x / cam->width;
    double pixel_view_x = (2.0 * pixel_ndc_x - 1.0) * cam->aspect_ratio * cam->fov_factor;
    V3 pixel_view = v3(pixel_view_x, pixel_view_y, -1.0);
    V3 pixel_world = xf_point(cam->view_to_world, pixel_view);
    V3 eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
    Ray r = {eye, normalized(sub(pixel_world, eye))};
    return r;
}

static void background_at(const PtRenderParams* p, const double* bg, uint32_t x, uint32_t y, double* out) {
    const double* src = p->bg_mode == PT_BG_PER_PIXEL ? bg + ((uint64_t)y * p->width + x) * 3
                        : p->bg_mode == PT_BG_PER_ROW ? bg + (uint64_t)y * 3
                                                      : bg;
    out[0] = src[0]; out[1] = src[1]; out[2] = src[2];
}

static uint8_t to_u8(double c) { /* Rust `as u8`: truncate, saturate, NaN -> 0. render.rs:143-147 */
    double v = c * 255.0;
    if (v != v) return 0;
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}
static double clamp01(double v) { /* vek Clamp::clamp01: partial_min(partial_max(v, 0), 1) */
    double lo = v >= 0.0 ? v : 0.0;
    return lo <= 1.0 ? lo : 1.0;
}

/* render_single_pixel, src/render.rs:22-51 */
static void render_single_pixel(Ctx* cx, const PtCamera* cam, const PtRenderParams* p, const double* bg, uint32_t x,
                                uint32_t y, double* color_out, uint32_t* hit_id, double* hit_t) {
    double background_color[3];
    background_at(p, bg, x, y, background_color);
    double total[3] = {0.0, 0.0, 0.0};
    PathKey key = {(uint64_t)y * p->width + x, 0};
    for (uint32_t s = 0; s < p->samples; ++s) {
        key.sample = s;
        double jx = draw(cx, key.pixel, s, 1, 0), jy = draw(cx, key.pixel, s, 1, 1);
        Ray ray = camera_ray_at(cam, (double)x + jx, (double)y + jy);
        double c[3];
        ray_color(cx, &ray, background_color, 0, key, 1, 0, c, s == 0 ? hit_id : NULL, s == 0 ? hit_t : NULL);
        total[0] = total[0] + c[0]; total[1] = total[1] + c[1]; total[2] = total[2] + c[2];
    }
    for (int ch = 0; ch < 3; ++ch) {
        double c = total[ch] / (double)p->samples;
        c = pow(c, 1.0 / GAMMA);
        color_out[ch] = clamp01(c);
    }
}

/* ------------------------------------------------------------------ drivers */
typedef struct {
    Ctx cx;
    const PtCamera* cam;
    const PtRenderParams* p;
    const double* bg;
    uint8_t* rgb;
    uint32_t* hit_id;
    double* hit_t;
    double* color;
    int tid, n_threads;
    uint32_t row_stride, row_offset; /* only rows y = row_offset + k * row_stride are rendered (benchmark sampling) */
    uint32_t* next_row;              /* shared cursor: rows are handed out dynamically, like rayon's work stealing */
} RenderJob;

static int pixel_owned(const PtRenderParams* p, uint32_t x, uint32_t y) {
    if (x < p->x1 || x > p->x2 || y < p->y1 || y > p->y2) return 0; /* render.rs:136-138 */
    if (p->world > 1) {
        uint32_t tw = p->tile_w ? p->tile_w : 32, th = p->tile_h ? p->tile_h : 32;
        uint32_t tiles_x = (p->width + tw - 1) / tw;
        uint64_t tile = (uint64_t)(y / th) * tiles_x + x / tw;
        if (tile % p->world != p->rank) return 0;
    }
    return 1;
}

#define ORACLE_SEGMENT 32u
static void* render_worker(void* arg) {
    RenderJob* j = (RenderJob*)arg;
    const PtRenderParams* p = j->p;
    /* the pixel loop, src/render.rs:127-150: the sampled rows of the slice are handed out dynamically in segments of
     * ORACLE_SEGMENT pixels (rayon's work stealing keeps every thread busy to the end of a frame; so does this) */
    const uint32_t seg_per_row = (p->width + ORACLE_SEGMENT - 1) / ORACLE_SEGMENT;
    for (;;) {
        const uint32_t unit = __atomic_fetch_add(j->next_row, 1u, __ATOMIC_RELAXED);
        const uint32_t k = unit / seg_per_row, seg = unit % seg_per_row;
        const uint64_t yy = (uint64_t)p->y1 + j->row_offset + (uint64_t)k * j->row_stride;
        if (yy > p->y2 || yy >= p->height) break;
        const uint32_t y = (uint32_t)yy;
        const uint32_t x_end = (seg + 1) * ORACLE_SEGMENT < p->width ? (seg + 1) * ORACLE_SEGMENT : p->width;
        for (uint32_t x = seg * ORACLE_SEGMENT; x < x_end; ++x) {
            if (!pixel_owned(p, x, y)) continue;
            uint64_t i = (uint64_t)y * p->width + x;
            double c[3];
            uint32_t id[2] = {0xFFFFFFFFu, 0};
            double ht = INFINITY;
            render_single_pixel(&j->cx, j->cam, p, j->bg, x, y, c, id, &ht);
            if (j->rgb) { j->rgb[i * 3] = to_u8(c[0]); j->rgb[i * 3 + 1] = to_u8(c[1]); j->rgb[i * 3 + 2] = to_u8(c[2]); }
            if (j->hit_id) { j->hit_id[i * 2] = id[0]; j->hit_id[i * 2 + 1] = id[1]; }
            if (j->hit_t) j->hit_t[i] = ht;
            if (j->color) { j->color[i * 3] = c[0]; j->color[i * 3 + 1] = c[1]; j->color[i * 3 + 2] = c[2]; }
        }
    }
    return NULL;
}

static void stats_add(OracleStats* a, const OracleStats* b) {
    a->rays_primary += b->rays_primary; a->rays_shadow += b->rays_shadow;
    a->rays_reflect += b->rays_reflect; a->rays_refract += b->rays_refract;
    a->rays_depth_cut += b->rays_depth_cut;
    a->kd_splits += b->kd_splits; a->instance_tests += b->instance_tests;
    a->triangle_tests += b->triangle_tests; a->bbox_gates += b->bbox_gates;
    a->shaded_hits += b->shaded_hits; a->texel_lookups += b->texel_lookups;
    a->shadow_kd_splits += b->shadow_kd_splits; a->shadow_instance_tests += b->shadow_instance_tests;
    a->shadow_triangle_tests += b->shadow_triangle_tests; a->shadow_bbox_gates += b->shadow_bbox_gates;
}

int oracle_render(const void* blob, uint64_t bytes, const PtCamera* cam, const PtRenderParams* params,
                  const double* background, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out,
                  double* color_out, int n_threads, OracleStats* stats) {
    return oracle_render_rows(blob, bytes, cam, params, background, rgb_inout, hit_id_out, hit_t_out, color_out, n_threads, 1, 0, stats);
}

int oracle_render_rows(const void* blob, uint64_t bytes, const PtCamera* cam, const PtRenderParams* params,
                       const double* background, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out,
                       double* color_out, int n_threads, uint32_t row_stride, uint32_t row_offset, OracleStats* stats) {
    Scene sc;
    int rc = scene_view(blob, bytes, &sc);
    if (rc != PT_OK) return rc;
    if (!cam || !params || !background || params->samples == 0 || params->width == 0 || params->height == 0) return PT_ERR_INVALID;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    RenderJob* jobs = (RenderJob*)calloc((size_t)n_threads, sizeof(RenderJob));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    uint32_t next_row = 0;
    if (row_stride == 0) row_stride = 1;
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].row_stride = row_stride; jobs[t].row_offset = row_offset; jobs[t].next_row = &next_row;
        jobs[t].cx.sc = &sc;
        jobs[t].cx.rng_mode = params->rng_mode;
        jobs[t].cx.seed = params->seed;
        jobs[t].cx.max_depth = params->max_depth ? params->max_depth : PT_MAX_RECURSION_DEPTH;
        jobs[t].cx.linear_scene = (params->flags & PT_RENDER_LINEAR_TLAS) != 0;
        jobs[t].cam = cam; jobs[t].p = params; jobs[t].bg = background;
        jobs[t].rgb = rgb_inout; jobs[t].hit_id = hit_id_out; jobs[t].hit_t = hit_t_out; jobs[t].color = color_out;
        jobs[t].tid = t; jobs[t].n_threads = n_threads;
        if (n_threads > 1) pthread_create(&th[t], NULL, render_worker, &jobs[t]);
    }
    if (n_threads == 1) render_worker(&jobs[0]);
    OracleStats total;
    memset(&total, 0, sizeof total);
    int error = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        stats_add(&total, &jobs[t].cx.st);
        if (!error && jobs[t].cx.error) error = jobs[t].cx.error;
    }
    if (stats) *stats = total;
    free(jobs);
    free(th);
    return error;
}

typedef struct {
    Ctx cx;
    uint64_t n, begin, end;
    const double *origins, *dirs, *bg;
    double* color;
    uint32_t* hit_id;
    double* hit_t;
} TraceJob;

static void* trace_worker(void* arg) {
    TraceJob* j = (TraceJob*)arg;
    for (uint64_t i = j->begin; i < j->end; ++i) {
        Ray ray = {v3(j->origins[i * 3], j->origins[i * 3 + 1], j->origins[i * 3 + 2]),
                   v3(j->dirs[i * 3], j->dirs[i * 3 + 1], j->dirs[i * 3 + 2])};
        PathKey key = {i, 0};
        double c[3];
        uint32_t id[2] = {0xFFFFFFFFu, 0};
        double ht = INFINITY;
        ray_color(&j->cx, &ray, j->bg, 0, key, 1, 0, c, id, &ht);
        if (j->color) { j->color[i * 3] = c[0]; j->color[i * 3 + 1] = c[1]; j->color[i * 3 + 2] = c[2]; }
        if (j->hit_id) { j->hit_id[i * 2] = id[0]; j->hit_id[i * 2 + 1] = id[1]; }
        if (j->hit_t) j->hit_t[i] = ht;
    }
    return NULL;
}

/* ray.color(&scene, background, 0) for explicit rays (what kdmesh.rs:155-163 does per ray).
 * The "pixel" of ray i in the RNG key is i, sample 0. */
int oracle_trace_rays(const void* blob, uint64_t bytes, uint64_t n, const double* origins, const double* dirs,
                      const double* background3, uint32_t rng_mode, uint64_t seed, uint32_t max_depth, double* color_out,
                      uint32_t* hit_id_out, double* hit_t_out, int n_threads, OracleStats* stats) {
    Scene sc;
    int rc = scene_view(blob, bytes, &sc);
    if (rc != PT_OK) return rc;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    TraceJob* jobs = (TraceJob*)calloc((size_t)n_threads, sizeof(TraceJob));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t per = (n + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].cx.sc = &sc;
        jobs[t].cx.rng_mode = rng_mode;
        jobs[t].cx.seed = seed;
        jobs[t].cx.max_depth = max_depth ? max_depth : PT_MAX_RECURSION_DEPTH;
        jobs[t].n = n;
        jobs[t].begin = per * (uint64_t)t < n ? per * (uint64_t)t : n;
        jobs[t].end = per * (uint64_t)(t + 1) < n ? per * (uint64_t)(t + 1) : n;
        jobs[t].origins = origins; jobs[t].dirs = dirs; jobs[t].bg = background3;
        jobs[t].color = color_out; jobs[t].hit_id = hit_id_out; jobs[t].hit_t = hit_t_out;
        if (n_threads > 1) pthread_create(&th[t], NULL, trace_worker, &jobs[t]);
    }
    if (n_threads == 1) trace_worker(&jobs[0]);
    OracleStats total;
    memset(&total, 0, sizeof total);
    int error = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (n_threads > 1) pthread_join(th[t], NULL);
        stats_add(&total, &jobs[t].cx.st);
        if (!error && jobs[t].cx.error) error = jobs[t].cx.error;
    }
    if (stats) *stats = total;
    free(jobs);
    free(th);
    return error;
}

/* Camera::ray_at for a list of (x, y) — used by the tests to build the ray
 * sets of kdmesh.rs:155-160 without re-deriving camera maths in Python. */
void oracle_camera_rays(const PtCamera* cam, uint64_t n, const double* xy, double* origins, double* dirs) {
    for (uint64_t i = 0; i < n; ++i) {
        Ray r = camera_ray_at(cam, xy[i * 2], xy[i * 2 + 1]);
        origins[i * 3] = r.origin.x; origins[i * 3 + 1] = r.origin.y; origins[i * 3 + 2] = r.origin.z;
        dirs[i * 3] = r.direction.x; dirs[i * 3 + 1] = r.direction.y; dirs[i * 3 + 2] = r.direction.z;
    }
}
