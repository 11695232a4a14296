// Ray casting on the device: the reference's kd-tree walk with all of its
// quirks (SURVEY §8 quirk list 12-14, Appendix A.1/A.2), per-instance object
// space intersection, and the `t`-only forms of every primitive's ray_hit.
// Hit attributes (normal, uv, normal-map basis) are NOT computed here: they are
// pure functions of (instance, sub id, t) and are rebuilt once in the shade
// kernel for the final hit (Appendix A.3).
//
//   KDTreeNode::ray_cast_impl   src/kdtree/node.rs:66-203
//   [T]::ray_cast / ray_hit     src/ray.rs:50-63,87-99
//   FlatSceneNode::ray_cast     src/flat_scene.rs:71-99
//   KDMesh::ray_hit             src/kdtree/kdmesh.rs:62-74
//   Mesh::ray_hit               src/primitive/mesh.rs:145-168
//   BoundingBox::test_hit       src/bounding_box.rs:104-116
//   Sphere/Cube/Plane/Cylinder/Cone/Triangle/InfinitePlane::ray_hit   src/primitive/*.rs
//   Quadratic::solve            src/math.rs:107-114 -> roots::find_roots_quadratic
#pragma once
#include "device_scene.cuh"

// tuning switches (A/B'd on the box with tools/build_variants.sh + tools/ab_bench.sh)
#ifndef PT_ANY_CHEAP_FIRST
#define PT_ANY_CHEAP_FIRST 1
#endif

namespace ptd {

struct WorkCounters {
    // the REFERENCE's work for the rays traced (what the oracle counts, culled or not)
    uint32_t kd_splits = 0, instance_tests = 0, triangle_tests = 0, bbox_gates = 0;
    uint32_t prim_flops = 0;  // sum over instance tests of the primitive's own f64 op count (SURVEY 8d P_type)
    // what the device EXECUTED (the physical roofline of bench.py): FP32 slab tests, exact f64 instance / triangle
    // tests and bbox gates that survived the culls, and the primitives' own f64 op counts of those instance tests
    uint32_t x_box = 0, x_inst = 0, x_tri = 0, x_gate = 0, x_prim_flops = 0;
};
// f64 add/mul/div/sqrt count of one analytic ray_hit: sphere 30, cube 150, plane 25, cylinder 70, cone 90;
// mesh types are accounted through their triangle tests and bbox gates
PT_D uint32_t prim_flop_count(uint32_t prim) {
    switch (prim) {
        case PT_PRIM_SPHERE: return 30u;
        case PT_PRIM_CUBE: return 150u;
        case PT_PRIM_PLANE: return 25u;
        case PT_PRIM_CYLINDER: return 70u;
        case PT_PRIM_CONE: return 90u;
        default: return 0u;
    }
}

struct Hit {
    double t;
    uint32_t inst;
    uint32_t sub;
};

// ------------------------------------------------------------------ quadratic
// roots 0.0.5 find_roots_quadratic: roots ascending, "do not use the smallest divisor"
PT_D int solve_quadratic(double a2, double a1, double a0, double& x_lo, double& x_hi) {
    if (a2 == 0.0) {
        if (a1 == 0.0) {
            if (a0 == 0.0) { x_lo = 0.0; return 1; }
            return 0;
        }
        x_lo = -a0 / a1;
        return 1;
    }
    const double discriminant = a1 * a1 - 4.0 * a2 * a0;
    if (discriminant < 0.0) return 0;
    const double a2x2 = 2.0 * a2;
    if (discriminant == 0.0) { x_lo = -a1 / a2x2; return 1; }
    const double sq = sqrt(discriminant);
    double same_sign, diff_sign;
    if (a1 < 0.0) { same_sign = -a1 + sq; diff_sign = -a1 - sq; }
    else { same_sign = -a1 - sq; diff_sign = -a1 + sq; }
    double x1, x2;
    if (fabs(same_sign) > fabs(a2x2)) {
        const double a0x2 = 2.0 * a0;
        if (fabs(diff_sign) > fabs(a2x2)) { x1 = a0x2 / same_sign; x2 = a0x2 / diff_sign; }
        else { x1 = a0x2 / same_sign; x2 = same_sign / a2x2; }
    } else { x1 = diff_sign / a2x2; x2 = same_sign / a2x2; }
    if (x1 < x2) { x_lo = x1; x_hi = x2; } else { x_lo = x2; x_hi = x1; }
    return 2;
}
// Solutions::find_in_range: the smallest root inside [s, e)
PT_D bool quadratic_in_range(double a, double b, double c, double s, double e, double& t) {
    double r0 = 0.0, r1 = 0.0;
    const int n = solve_quadratic(a, b, c, r0, r1);
    if (n >= 1 && in_range(s, e, r0)) { t = r0; return true; }
    if (n == 2 && in_range(s, e, r1)) { t = r1; return true; }
    return false;
}

// ------------------------------------------------------------------ analytic primitives (t and sub id only)
PT_D bool cube_contains(V3 p) {  // cube.rs:22-27
    const double radius = 0.5 + kEps;
    return -radius <= p.x && p.x <= radius && -radius <= p.y && p.y <= radius && -radius <= p.z && p.z <= radius;
}

// t of InfinitePlane::ray_hit for cube face f (infinite_plane.rs:46-79 with the
// axis-unit normals and on-axis points of cube.rs:46-65; the dot products reduce
// to one component, the other products are exact zeros)
PT_D double cube_face_t(int f, V3 o, V3 d) {
    switch (f) {
        case 0: return -(o.x - 0.5) / d.x;    // Right  n=+x p=+0.5
        case 1: return (o.x + 0.5) / (-d.x);  // Left   n=-x p=-0.5
        case 2: return -(o.y - 0.5) / d.y;    // Top
        case 3: return (o.y + 0.5) / (-d.y);  // Bottom
        case 4: return -(o.z - 0.5) / d.z;    // Near
        default: return (o.z + 0.5) / (-d.z); // Far
    }
}

// Cube::ray_hit fold over the 6 faces, cube.rs:67-82. ANY: stop at the first accepted face.
template <bool ANY>
PT_D bool cube_t(V3 o, V3 d, double s, double e, double& t_out, uint32_t& face_out) {
    bool found = false;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        const double t = cube_face_t(f, o, d);
        if (in_range(s, e, t) && cube_contains(ray_at(o, d, t))) {
            e = t;
            t_out = t;
            face_out = (uint32_t)f;
            found = true;
            if (ANY) return true;
        }
    }
    return found;
}

PT_D bool sphere_t(V3 o, V3 d, double s, double e, double& t) {  // sphere.rs:49-56
    const double a = dot(d, d);
    const double b = 2.0 * dot(o, d);
    const double c = dot(o, o) - 1.0;
    return quadratic_in_range(a, b, c, s, e, t);
}

PT_D bool plane_t(V3 o, V3 d, double s, double e, double& t_out) {  // plane.rs:35-52, infinite_plane.rs:63-69
    const double t = -o.y / d.y;
    if (!in_range(s, e, t)) return false;
    const V3 p = ray_at(o, d, t);
    const double radius = 0.5 + kEps;
    if (!(-radius <= p.x && p.x <= radius && -radius <= p.z && p.z <= radius)) return false;
    t_out = t;
    return true;
}

PT_D bool cyl_cap_t(double height, V3 o, V3 d, double s, double e, double& t_out) {  // cylinder.rs:77-116, cone.rs:116-157
    const double t = (height - o.y) / d.y;
    if (!in_range(s, e, t)) return false;
    const V3 p = ray_at(o, d, t);
    if ((p.x * p.x + p.z * p.z) > 0.5 * 0.5) return false;
    t_out = t;
    return true;
}

template <bool ANY>
PT_D bool cylinder_t(V3 o, V3 d, double s, double e, double& t_out, uint32_t& part) {  // cylinder.rs:28-74,118-153
    bool found = false;
    {
        const double a = d.x * d.x + d.z * d.z;
        const double b = 2.0 * o.x * d.x + 2.0 * o.z * d.z;
        const double c = o.x * o.x + o.z * o.z - 0.5 * 0.5;
        double t;
        if (quadratic_in_range(a, b, c, s, e, t)) {
            const V3 p = ray_at(o, d, t);
            if (!(p.y > 0.5 || p.y < -0.5)) { e = t; t_out = t; part = 0; found = true; if (ANY) return true; }
        }
    }
    double t;
    if (cyl_cap_t(0.5, o, d, s, e, t)) { e = t; t_out = t; part = 1; found = true; if (ANY) return true; }
    if (cyl_cap_t(-0.5, o, d, s, e, t)) { t_out = t; part = 2; found = true; }
    return found;
}

template <bool ANY>
PT_D bool cone_t(V3 o, V3 d, double s, double e, double& t_out, uint32_t& part) {  // cone.rs:28-113,159-186
    bool found = false;
    {
        const double HEIGHT = 1.0, RADIUS = 0.5;
        const double h_sqr = HEIGHT * HEIGHT;
        const double r_sqr = RADIUS * RADIUS;
        const double a = 4.0 * d.y * d.y * r_sqr - 4.0 * h_sqr * (d.x * d.x + d.z * d.z);
        const double b = -8.0 * h_sqr * (d.x * o.x + d.z * o.z) - 4.0 * r_sqr * (d.y * HEIGHT - 2.0 * d.y * o.y);
        const double c = -4.0 * h_sqr * (o.x * o.x + o.z * o.z) + r_sqr * (h_sqr - 4.0 * HEIGHT * o.y + 4.0 * o.y * o.y);
        double t;
        if (quadratic_in_range(a, b, c, s, e, t)) {
            const V3 p = ray_at(o, d, t);
            if (!(p.y > 0.5 || p.y < -0.5)) { e = t; t_out = t; part = 0; found = true; if (ANY) return true; }
        }
    }
    double t;
    if (cyl_cap_t(-0.5, o, d, s, e, t)) { t_out = t; part = 1; found = true; }
    return found;
}

// Triangle::ray_hit up to the barycentric tests (triangle.rs:38-80)
struct TriBary {
    double beta, gamma;
};
// `tie`: also accept t == e_ (only mesh_fold asks: it visits the triangles out of index order and settles exact ties itself)
PT_D bool triangle_t(const PtTriPos* __restrict__ tp, V3 o, V3 dir, double s, double e_, double& t_out, TriBary* bary, bool tie = false) {
    const double* v = reinterpret_cast<const double*>(tp);
    const V3 A = v3(__ldg(v + 0), __ldg(v + 1), __ldg(v + 2));
    const V3 B = v3(__ldg(v + 3), __ldg(v + 4), __ldg(v + 5));
    const V3 C = v3(__ldg(v + 6), __ldg(v + 7), __ldg(v + 8));
    const V3 ab = A - B, ac = A - C, ao = A - o;
    const double a = ab.x, b = ab.y, c = ab.z;
    const double d = ac.x, e = ac.y, f = ac.z;
    const double g = dir.x, h = dir.y, i = dir.z;
    const double j = ao.x, k = ao.y, l = ao.z;

    const double ei_hf = e * i - h * f;
    const double gf_di = g * f - d * i;
    const double dh_eg = d * h - e * g;
    const double m = a * ei_hf + b * gf_di + c * dh_eg;

    const double ak_jb = a * k - j * b;
    const double jc_al = j * c - a * l;
    const double bl_ck = b * l - c * k;

    const double t = -(f * ak_jb + e * jc_al + d * bl_ck) / m;
    if (!in_range(s, e_, t) && !(tie && t == e_ && s <= t)) return false;
    const double gamma = (i * ak_jb + h * jc_al + g * bl_ck) / m;
    if (gamma < 0.0 || gamma > 1.0) return false;
    const double beta = (j * ei_hf + k * gf_di + l * dh_eg) / m;
    if (beta < 0.0 || beta > 1.0 - gamma) return false;
    t_out = t;
    if (bary) { bary->beta = beta; bary->gamma = gamma; }
    return true;
}

// BoundingBox::test_hit, bounding_box.rs:104-116: is_some() only
PT_D bool bbox_gate(const PtMesh* __restrict__ mesh, V3 o, V3 d, double s, double e) {
    double m[12];
    load_doubles12(mesh->bbox_invtrans, m);
    const V3 lo = xf_point(m, o), ld = xf_dir(m, d);
    if (cube_contains(ray_at(lo, ld, s))) return true;
    double t;
    uint32_t face;
    return cube_t<true>(lo, ld, s, e, t, face);
}

// ------------------------------------------------------------------ conservative FP32 cull
// Before a candidate is tested exactly (f64, object space) a padded FP32 box that contains every hit it could return
// is slab-tested against the ray.  The test may only answer "certainly no hit inside [s, e)": every rounding source
// is covered by padding that is orders of magnitude larger than the FP32 error (the boxes are rounded outward and
// padded at build time, the ray origin by 4e-6 * |o|, the slab parameters by 2e-5 relative), so a candidate the f64
// test would accept is never dropped and the result stays bit-identical to the un-culled walk.  The cull runs on the
// FP32 pipe (explicit FMAs: the library is built -fmad=false for the f64 path, and nothing here needs to round like the
// reference); the f64 pipe only sees the survivors.
//
// kClipPadT: boxes CLIPPED to a k-d leaf's cell (leaf_cull.cu) rest on "a hit the leaf accepts lies inside the leaf's
// cell".  The walk above decides sides from two probe points, ray.at(s + EPSILON) and ray.at(t_max) (node.rs:119-131):
// between them the ray is on the side it was sent to, but a hit in the first or last EPSILON of an ancestor's range —
// or, when a range is shorter than 2 EPSILON, up to 2 EPSILON from the plane's crossing — can lie on the other side,
// by at most 2 EPSILON of ray PARAMETER whatever the direction's length.  So the slab parameters of a clipped box get
// an ABSOLUTE pad as well (each box side 2 EPSILON: 4e-5 between them).  Beyond t_max = s + extent the ray is outside
// the tree's bounds for every ray that passes `probe_covers` below; the others use the unclipped box set.
constexpr float kClipPadT = 4e-5f;

struct RayF {
    float ix, iy, iz;           // 1 / direction (inf when a component is 0)
    float cx_lo, cy_lo, cz_lo;  // -(origin + pad) / direction: box minima * i + c_lo = slab parameter
    float cx_hi, cy_hi, cz_hi;  // -(origin - pad) / direction: box maxima
};
PT_D RayF make_rayf(V3 o, V3 d) {
    RayF r;
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    const float pad = 4e-6f * fmaxf(fabsf(ox), fmaxf(fabsf(oy), fabsf(oz)));
    r.ix = 1.0f / (float)d.x; r.iy = 1.0f / (float)d.y; r.iz = 1.0f / (float)d.z;
    r.cx_lo = -((ox + pad) * r.ix); r.cy_lo = -((oy + pad) * r.iy); r.cz_lo = -((oz + pad) * r.iz);
    r.cx_hi = -((ox - pad) * r.ix); r.cy_hi = -((oy - pad) * r.iy); r.cz_hi = -((oz - pad) * r.iz);
    return r;
}
// the ray's range as the slab tests see it: [s, e) widened by 1e-4 relative, the far end also by kClipPadT
struct RangeF {
    float s, e;
};
PT_D RangeF make_rangef(double s, double e) { return RangeF{(float)s * 0.9999f, (float)e * 1.0001f + kClipPadT}; }  // s, e > 0

// padded slab parameters of the ray inside the box; fminf / fmaxf drop a NaN operand (0 * inf on a slab boundary):
// the slab then does not constrain
PT_D void box_interval(const float4* __restrict__ bb, const RayF& r, float& tn, float& tf) {
    const float4 lo = __ldg(bb), hi = __ldg(bb + 1);
    const float ax = __fmaf_rn(lo.x, r.ix, r.cx_lo), bx = __fmaf_rn(hi.x, r.ix, r.cx_hi);
    const float ay = __fmaf_rn(lo.y, r.iy, r.cy_lo), by = __fmaf_rn(hi.y, r.iy, r.cy_hi);
    const float az = __fmaf_rn(lo.z, r.iz, r.cz_lo), bz = __fmaf_rn(hi.z, r.iz, r.cz_hi);
    tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    tn = __fmaf_rn(-2e-5f, fabsf(tn), tn);
    tf = __fmaf_rn(2e-5f, fabsf(tf), tf) + kClipPadT;
}
// false = the ray certainly has no hit inside the box within the range
PT_D bool box_may_hit(const float4* __restrict__ bb, const RayF& r, const RangeF& rg) {
    float tn, tf;
    box_interval(bb, r, tn, tf);
    return !(tn > tf) && !(tf < rg.s) && !(tn > rg.e);
}
// the same for a box that may be EMPTY (lo > hi: a subtree without a single candidate): nothing to hit
PT_D bool box_may_hit_nonempty(const float4* __restrict__ bb, const RayF& r, const RangeF& rg) {
    const float4 lo = __ldg(bb), hi = __ldg(bb + 1);
    if (lo.x > hi.x) return false;
    const float ax = __fmaf_rn(lo.x, r.ix, r.cx_lo), bx = __fmaf_rn(hi.x, r.ix, r.cx_hi);
    const float ay = __fmaf_rn(lo.y, r.iy, r.cy_lo), by = __fmaf_rn(hi.y, r.iy, r.cy_hi);
    const float az = __fmaf_rn(lo.z, r.iz, r.cz_lo), bz = __fmaf_rn(hi.z, r.iz, r.cz_hi);
    float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    tn = __fmaf_rn(-2e-5f, fabsf(tn), tn);
    tf = __fmaf_rn(2e-5f, fabsf(tf), tf) + kClipPadT;
    return !(tn > tf) && !(tf < rg.s) && !(tn > rg.e);
}
// The probe segment of the walk ends at s + extent (node.rs:119): does the ray leave `box` — a box around everything
// the tree holds — before that?  Then no hit lies beyond the probe and the clipped boxes are valid for the whole walk.
// A ray that misses the box altogether cannot hit anything: any box set will do.
PT_D bool probe_covers(const float4* __restrict__ box, const RayF& r, double s, double extent) {
    float tn, tf;
    box_interval(box, r, tn, tf);
    return (tn > tf) || (tf < (float)((s + extent) * 0.999));  // NaN / inf anywhere: false
}

// ------------------------------------------------------------------ kd walk
// One entry per pending far child: the child's 16-byte record itself (it was fetched together with the near child, so a
// pop costs ONE local-memory round trip instead of index -> node) and the start of its range.  The END of a pending
// range is not stored: a push happens with the current range [s, e) and leaves [s, tp) current and [tp, e) pending, so
// the current `e` is always the `s` of the entry on top of the stack (the walk's initial `e` when the stack is empty),
// and a popped entry's end is the `s` of the entry below it.
struct KdStack {
    uint4 node[PT_MAX_KD_STACK];
    double s[PT_MAX_KD_STACK];
};

// Iterative form of ray_cast_impl (node.rs:66-203). `leaf(rank, gbase, first, count, s, e)` returns true when the
// leaf's fold produced a hit inside [s, e) — (rank, gbase) locate the leaf's cull boxes, leaf_cull.cu —; the first
// leaf that does ends the walk.
// "while-while" form: every lane first descends to its next leaf (a short loop of split steps), and only then do the
// lanes of the warp run their leaves together.  With one loop whose body is "a split step OR a whole leaf", a lane that
// is still descending advances ONE split per leaf any other lane of its warp processes — and a leaf can hold a whole
// KDMesh walk.
// `od`: the ray as an ARRAY (origin x, y, z, direction x, y, z) that the split step indexes by the node's axis: one
// local-memory load per operand (an L1 hit) where selecting among three register pairs costs a branch and six moves
// per operand, and the ray does not occupy twelve registers for the length of the walk.
#ifndef PT_KD_RAY_ARRAY
#define PT_KD_RAY_ARRAY 1
#endif
// PRUNE: before the walk enters a child it slab-tests the child's NODE BOX — the union of the occupied boxes of every
// leaf below it (leaf_cull.cu) — against the range the child would be walked with; a subtree the ray certainly misses
// there is not entered (the far child is not even pushed).  Such a subtree holds no leaf that can return a hit, so
// hits, hit ids and `t` are exactly those of the full walk; what is lost is the reference's WORK below (the counting
// kernels walk in full) and a kd-plane panic (PT_DEVERR_KD_PLANE) the reference would have hit down there.
template <bool PRUNE, class LeafFn>
PT_D bool kd_walk(const PtKdNode* __restrict__ nodes, double extent, const double (&od)[6], double s, const double e0, KdStack& stack,
                  LeafFn& leaf, uint32_t& err, uint32_t& n_splits, const float4* __restrict__ node_boxes, const RayF& rf) {
    int sp = 0;
    double e = e0;
    const uint4* __restrict__ nodes4 = reinterpret_cast<const uint4*>(nodes);
    uint4 w = __ldg(nodes4);
    if (PRUNE && (w.z & 3u) != 3u && !box_may_hit_nonempty(node_boxes, rf, make_rangef(s, e))) return false;
    for (;;) {
        bool dead = false;  // this subtree yields nothing: the reference would have panicked in it, or (PRUNE) the ray misses its box
        while ((w.z & 3u) != 3u) {
            const uint32_t axis = w.z & 3u;
            const double split = __hiloint2double((int)w.y, (int)w.x);
            ++n_splits;
            const uint32_t front = w.z >> 2, back = w.w;
            // both children are fetched before the side tests: the loads' latency overlaps the f64 dependency chain
            // below instead of following it (the node fetch was the top long-scoreboard stall of the walk)
            const uint4 wf = __ldg(nodes4 + front), wb = __ldg(nodes4 + back);
            // node.rs:119-127
            double t_max = s + extent;
            if (!in_range(s, e, t_max)) t_max = e - kEps;
            const double t_min = s + kEps;
#if PT_KD_RAY_ARRAY
            const double oa = od[axis], da = od[3 + axis];
#else
            const double oa = axis == 0 ? od[0] : (axis == 1 ? od[1] : od[2]);
            const double da = axis == 0 ? od[3] : (axis == 1 ? od[4] : od[5]);
#endif
            const double p0 = oa + da * t_min;
            const double p1 = oa + da * t_max;
            const bool f0 = (p0 - split) >= 0.0;  // which_side, infinite_plane.rs:27-35
            const bool f1 = (p1 - split) >= 0.0;
            if (f0 != f1) {
                const double tp = (split - oa) / da;  // ray_hit_axis_aligned_plane, node.rs:90-110
                if (!in_range(s, e, tp)) {
                    err |= PT_DEVERR_KD_PLANE;  // .expect("bug: ray should definitely hit infinite plane")
                    dead = true;
                    break;
                }
                // near child on [s, tp); if it misses, far child on [tp, e). node.rs:150-166
                const uint4 wfar = f0 ? wb : wf;
                // (a far LEAF is pushed as it is: its fold starts with the same test of the same box)
                if (!PRUNE || (wfar.z & 3u) == 3u || box_may_hit_nonempty(node_boxes + 2 * (size_t)(f0 ? back : front), rf, make_rangef(tp, e))) {
                    stack.node[sp] = wfar;
                    stack.s[sp] = tp;
                    ++sp;
                } else {
                    // not pushed: the walk's "current end = start of the entry on top" bookkeeping needs the entry all
                    // the same when the near side comes back empty — push a marker that pops to nothing
                    stack.node[sp] = make_uint4(0u, 0u, 3u, 0u);  // an empty leaf
                    stack.s[sp] = tp;
                    ++sp;
                }
                e = tp;
            }
            w = f0 ? wf : wb;  // node.rs:134-137
            if (PRUNE && (w.z & 3u) != 3u && !box_may_hit_nonempty(node_boxes + 2 * (size_t)(f0 ? front : back), rf, make_rangef(s, e))) {
                dead = true;
                break;
            }
        }
        if (!dead && w.w != 0u && leaf(w.x, w.y, w.z >> 2, w.w, s, e)) return true;
        if (sp == 0) return false;
        --sp;
        w = stack.node[sp];
        s = stack.s[sp];
        e = sp ? stack.s[sp - 1] : e0;
    }
}

// One leaf of a k-d tree: the fold over its candidate list with a shrinking range (ray.rs:50-63 / :87-99), restricted
// to the candidates whose boxes the ray may hit inside the range.  Three levels of boxes (leaf_cull.cu): the leaf's
// occupied box, the union box of every run of 8 list positions, the candidates' own boxes.  Per 32 positions the lanes
// first slab-test boxes into a survivor mask — a tight, branch-light loop — and then run the exact tests of the
// survivors in list order, each culled again against the range as it has shrunk since: a one-phase loop (test a box,
// then maybe the primitive) makes the whole warp wait whenever ANY lane has a survivor.
//   exact(k, s, e): the reference's test of list position k over [s, e); on a hit it records it, sets e = t and returns true.
template <bool ANY, class Exact>
PT_D bool leaf_fold(const LeafCull& lc, uint32_t set, uint32_t rank, uint32_t gbase, uint32_t count, const RayF& rf, double s, double& e,
                    Exact& exact, uint32_t& n_box) {
    RangeF rg = make_rangef(s, e);
    ++n_box;
    if (!box_may_hit(lc.occ + set + 2 * (size_t)rank, rf, rg)) return false;
    const float4* __restrict__ grp = lc.grp + set + 2 * (size_t)gbase;
    const float4* __restrict__ item = lc.item + set + 16 * (size_t)gbase;
    const uint32_t n_groups = (count + 7u) >> 3;
    bool found = false;
    for (uint32_t gc = 0; gc < n_groups; gc += 4u) {  // 4 runs of 8 = one 32-bit survivor mask
        uint32_t mask = 0u;
        const uint32_t g_stop = min(gc + 4u, n_groups);
        for (uint32_t g = gc; g < g_stop; ++g) {
            if (n_groups > 1u) {  // a single run's box is the occupied box
                ++n_box;
                if (!box_may_hit(grp + 2 * (size_t)g, rf, rg)) continue;
            }
            n_box += 8u;
            const float4* __restrict__ ib = item + 16 * (size_t)g;
            uint32_t m8 = 0u;
#pragma unroll
            for (uint32_t j = 0; j < 8u; ++j)
                if (box_may_hit(ib + 2 * j, rf, rg)) m8 |= 1u << j;
            mask |= m8 << ((g - gc) << 3);
        }
        const uint32_t left = count - (gc << 3);
        if (left < 32u) mask &= (1u << left) - 1u;  // the padding slots of the last run
#if PT_ANY_CHEAP_FIRST
        if (ANY) {
            // "is there a hit" does not depend on the order of the tests: the analytic candidates go first, a mesh
            // candidate (a whole KDMesh walk / Mesh fold) only when none of them stopped the ray
            uint32_t later = 0u;
            while (mask) {
                const uint32_t bit = mask & (0u - mask);
                mask ^= bit;
                const uint32_t k = (gc << 3) + (uint32_t)__ffs((int)bit) - 1u;
                if (exact.expensive(k)) { later |= bit; continue; }
                if (exact(k, s, e)) return true;
            }
            mask = later;
        }
#endif
        while (mask) {
            const uint32_t k = (gc << 3) + (uint32_t)__ffs((int)mask) - 1u;
            mask &= mask - 1u;
            if (found) {  // the range has shrunk since phase 1
                ++n_box;
                if (!box_may_hit(item + 2 * (size_t)k, rf, rg)) continue;
            }
            if (exact(k, s, e)) {
                if (ANY) return true;
                found = true;
                rg = make_rangef(s, e);
            }
        }
    }
    return found;
}

// leaf of a KDMesh tree: RayHit for [Triangle] with a shrinking clone of the range (ray.rs:50-63)
template <bool ANY>
struct BlasLeaf {
    const LeafCull& lc;
    uint32_t set;
    const uint32_t* __restrict__ items;
    const PtTriPos* __restrict__ tris;
    V3 o, d;
    RayF rf;
    double t;
    uint32_t tri;
    uint32_t n_tests;
    uint32_t first;
    uint32_t x_box, x_tri;  // executed slab / triangle tests
    PT_D bool expensive(uint32_t) const { return false; }
    PT_D bool operator()(uint32_t k, double s, double& e) {
        const uint32_t idx = __ldg(items + first + k);
        double tt;
        ++x_tri;
        if (!triangle_t(tris + idx, o, d, s, e, tt, nullptr)) return false;
        e = tt;
        t = tt;
        tri = idx;
        if (ANY) n_tests += k + 1u;
        return true;
    }
    PT_D bool operator()(uint32_t rank, uint32_t gbase, uint32_t first_, uint32_t count, double s, double e) {
        first = first_;
        if (!ANY) n_tests += count;
        const uint32_t before = n_tests;
        const bool found = leaf_fold<ANY>(lc, set, rank, gbase, count, rf, s, e, *this, x_box);
        if (ANY && !found) n_tests = before + count;
        return found;
    }
};

// Mesh::ray_hit's fold over EVERY triangle of the mesh in index order with a shrinking range (mesh.rs:157-167,
// ray.rs:50-63).  A triangle is accepted when its own test passes with t inside [s, e_now) and then e_now = t, so what
// the fold returns is the triangle with the smallest t, and among exactly equal t the one with the smallest index: the
// lexicographic minimum of (t, index) over the triangles whose test passes inside [s, e) — a quantity that does not
// depend on the order of evaluation.  It is evaluated here in MORTON order of the triangle centroids (DScene::fold_order,
// fold_order.cu), where consecutive positions are neighbours in space and the union boxes of aligned runs of 4, 16,
// 64 ... 4096 positions (DScene::fold_aabb[1..]) form an implicit 4-ary hierarchy: a run whose padded FP32 box the ray
// certainly misses inside [s, e_now] is skipped whole (coarsest aligned level first), a surviving triangle gets the exact
// f64 test, ties on t are settled by index.  All boxes are rounded outward and padded like the instance boxes, and
// the cull range is closed at e_now so that an exact tie is never culled.  ANY mode returns at the first accepted
// triangle ("is there a hit" does not depend on the order either).  The count of triangle tests is the reference's:
// all of them for a closest-hit fold.
// A 5 804-triangle fold costs a few dozen FP32 box tests and a handful of f64 triangle tests instead of 5 804 f64 tests.
template <bool ANY>
PT_D bool mesh_fold(const DScene& sc, uint32_t tri_first, uint32_t tri_count, V3 o, V3 d, double s, double e, double& t_out,
                    uint32_t& sub, uint32_t& n_tests, uint32_t& x_box, uint32_t& x_tri) {
    const RayF rf = make_rayf(o, d);
    RangeF rg = make_rangef(s, e);
    const uint32_t* __restrict__ order = sc.fold_order;
    const float4* __restrict__ bb0 = sc.fold_aabb[0];
    const uint32_t end = tri_first + tri_count;
    const int levels = (int)sc.fold_levels;
    bool found = false;
    uint32_t best = 0xFFFFFFFFu;
    uint32_t k = tri_first;
    while (k < end) {
        // the coarsest level this position starts a run of, then finer ones: one box test per visited node
        int l = k ? min(levels, (__ffs((int)k) - 1) >> 1) : levels;
        bool skipped = false;
        for (; l >= 1; --l) {
            ++x_box;
            if (!box_may_hit(sc.fold_aabb[l] + 2 * (size_t)(k >> (2 * l)), rf, rg)) {
                k = min(k + (1u << (2 * l)), end);
                skipped = true;
                break;
            }
        }
        if (skipped) continue;
        double tt;
        ++x_box;
        if (box_may_hit(bb0 + 2 * (size_t)k, rf, rg)) {
            const uint32_t idx = __ldg(order + k);
            ++x_tri;
            if (triangle_t(sc.tri_pos + idx, o, d, s, e, tt, nullptr, found) && (tt < e || idx < best)) {
                if (ANY) { n_tests += 1u; t_out = tt; sub = idx - tri_first; return true; }
                e = tt;
                rg = make_rangef(s, e);
                best = idx;
                found = true;
            }
        }
        ++k;
    }
    n_tests += tri_count;
    if (found) { t_out = e; sub = best - tri_first; }
    return found;
}

// Mesh / KDMesh::ray_hit in object space: the bounding-box gate (mesh.rs:153, kdmesh.rs:67), then the fold over every
// triangle (mesh.rs:157-167) or the walk of the mesh's own k-d tree on a clone of the range (node.rs:33-51).
// `world_exit`: upper bound of the ray parameter at which the ray leaves the instance's box (KDMesh only: decides
// whether the KDMesh walk may use the clipped triangle boxes, see probe_covers).
template <bool ANY, bool PRUNE>
PT_D bool mesh_kinds_t(const DScene& sc, uint32_t prim, const PtMesh* mesh, V3 o, V3 d, double s, double e, double& t, uint32_t& sub,
                       KdStack& blas_stack, uint32_t& err, WorkCounters& wc, float world_exit) {
    const uint32_t tri_first = __ldg(&mesh->tri_first);
    ++wc.x_gate;
    if (!bbox_gate(mesh, o, d, s, e)) return false;  // (the reference's count: by the caller)
    if (prim == PT_PRIM_MESH) {
        uint32_t n_tests = 0;
        const bool found = mesh_fold<ANY>(sc, tri_first, __ldg(&mesh->tri_count), o, d, s, e, t, sub, n_tests, wc.x_box, wc.x_tri);
        wc.triangle_tests += n_tests;
        return found;
    }
    const uint32_t item_first = __ldg(&mesh->item_first);
    const double extent = __ldg(&mesh->extent);
    // the walk's first probe ends at s + extent (object-space extent on the world-ray parameter): the
    // clipped triangle boxes hold only if the ray has left the mesh by then
    const uint32_t set = world_exit < (float)((s + extent) * 0.999) ? 0u : sc.bl_cull.set_stride;
    BlasLeaf<ANY> leaf{sc.bl_cull, set, sc.blas_items + item_first, sc.tri_pos + tri_first, o, d, make_rayf(o, d), 0.0, 0, 0, 0, 0, 0};
    const double od[6] = {o.x, o.y, o.z, d.x, d.y, d.z};
    const uint32_t node_first = __ldg(&mesh->node_first);
    const bool hit = kd_walk<PRUNE>(sc.blas_nodes + node_first, extent, od, s, e, blas_stack, leaf, err, wc.kd_splits,
                                    sc.bl_cull.node + set + 2 * (size_t)node_first, leaf.rf);
    wc.triangle_tests += leaf.n_tests;
    wc.x_box += leaf.x_box;
    wc.x_tri += leaf.x_tri;
    if (hit) { t = leaf.t; sub = leaf.tri; }
    return hit;
}

// Primitive::ray_hit in object space (primitive.rs:55-61), t + sub id only.
template <bool ANY, bool COUNT, bool PRUNE>
PT_D bool primitive_t(const DScene& sc, uint32_t prim, uint32_t mesh_id, V3 o, V3 d, double s, double e, double& t,
                      uint32_t& sub, KdStack& blas_stack, uint32_t& err, WorkCounters& wc, float world_exit) {
    sub = 0;
    switch (prim) {
        case PT_PRIM_SPHERE: return sphere_t(o, d, s, e, t);
        case PT_PRIM_CUBE: return cube_t<ANY>(o, d, s, e, t, sub);
        case PT_PRIM_PLANE: return plane_t(o, d, s, e, t);
        case PT_PRIM_CYLINDER: return cylinder_t<ANY>(o, d, s, e, t, sub);
        case PT_PRIM_CONE: return cone_t<ANY>(o, d, s, e, t, sub);
        default: break;
    }
    const PtMesh* mesh = sc.meshes + mesh_id;
    if (prim == PT_PRIM_TRIANGLE) { ++wc.x_tri; return triangle_t(sc.tri_pos + __ldg(&mesh->tri_first), o, d, s, e, t, nullptr); }
    return mesh_kinds_t<ANY, PRUNE>(sc, prim, mesh, o, d, s, e, t, sub, blas_stack, err, wc, world_exit);
}

// FlatSceneNode::ray_cast of instance `inst` (flat_scene.rs:71-99): the ray in object space, direction NOT renormalised
// (flat_scene.rs:75, ray.rs:130-135), then the primitive's own test over [s, e).  `pm` = (prim, mesh) of the record.
template <bool ANY, bool COUNT, bool PRUNE>
PT_D bool instance_t(const DScene& sc, uint32_t inst, uint2 pm, V3 o, V3 d, const RayF& rf, double s, double e, double& t,
                     uint32_t& sub, KdStack& blas_stack, uint32_t& err, WorkCounters& wc) {
    const PtInstance* rec = sc.instances + inst;
    double m[12];
    load_doubles12(rec->invtrans, m);
    const V3 lo = xf_point(m, o), ld = xf_dir(m, d);
    float world_exit = INFINITY;
    if (pm.x == PT_PRIM_KDMESH) {
        float tn;
        box_interval(sc.inst_aabb + 2 * (size_t)inst, rf, tn, world_exit);
    }
    ++wc.x_inst;
    wc.x_prim_flops += prim_flop_count(pm.x);
    return primitive_t<ANY, COUNT, PRUNE>(sc, pm.x, pm.y, lo, ld, s, e, t, sub, blas_stack, err, wc, world_exit);
}

// leaf of the scene tree: RayCast for [FlatSceneNode] sharing one shrinking range (ray.rs:87-99)
template <bool ANY, bool COUNT, bool PRUNE>
struct TlasLeaf {
    const DScene& sc;
    uint32_t set;
    V3 o, d;
    RayF rf;
    KdStack& blas_stack;
    Hit& hit;
    uint32_t& err;
    WorkCounters& wc;
    uint32_t first;
    PT_D bool expensive(uint32_t k) const {  // a Mesh / KDMesh candidate
        const uint32_t prim = __ldg(&sc.instances[__ldg(sc.tlas_items + first + k)].prim);
        return prim == PT_PRIM_MESH || prim == PT_PRIM_KDMESH;
    }
    PT_D bool operator()(uint32_t k, double s, double& e) {  // list position k
        const uint32_t inst = __ldg(sc.tlas_items + first + k);
        const uint2 pm = __ldg(reinterpret_cast<const uint2*>(&sc.instances[inst].prim));
        double t;
        uint32_t sub;
        if (!instance_t<ANY, COUNT, PRUNE>(sc, inst, pm, o, d, rf, s, e, t, sub, blas_stack, err, wc)) return false;
        e = t;  // flat_scene.rs:92
        hit.t = t;
        hit.inst = inst;
        hit.sub = sub;
        return true;
    }
    PT_D bool operator()(uint32_t rank, uint32_t gbase, uint32_t first_, uint32_t count, double s, double e) {
        first = first_;
        if (COUNT) {
            // the reference's work for every candidate, culled or not
            for (uint32_t k = first; k < first + count; ++k) {
                const uint32_t prim = __ldg(&sc.instances[__ldg(sc.tlas_items + k)].prim);
                ++wc.instance_tests;
                wc.prim_flops += prim_flop_count(prim);
                wc.bbox_gates += (prim == PT_PRIM_MESH || prim == PT_PRIM_KDMESH) ? 1u : 0u;  // mesh.rs:153, kdmesh.rs:67
                wc.triangle_tests += prim == PT_PRIM_TRIANGLE ? 1u : 0u;
            }
        }
        return leaf_fold<ANY>(sc.tl_cull, set, rank, gbase, count, rf, s, e, *this, wc.x_box);
    }
};

// scene.root.ray_cast(ray, [EPSILON, inf)) — ray.rs:140-141, material.rs:174-179
template <bool ANY, bool COUNT, bool PRUNE>
PT_D bool scene_cast(const DScene& sc, V3 o, V3 d, Hit& hit, KdStack& tlas_stack, KdStack& blas_stack, uint32_t& err,
                     WorkCounters& wc) {
    const RayF rf = make_rayf(o, d);
    const uint32_t set = probe_covers(sc.tl_root, rf, kEps, sc.tlas_extent) ? 0u : sc.tl_cull.set_stride;
    TlasLeaf<ANY, COUNT, PRUNE> leaf{sc, set, o, d, rf, blas_stack, hit, err, wc, 0};
    const double od[6] = {o.x, o.y, o.z, d.x, d.y, d.z};
    return kd_walk<PRUNE>(sc.tlas_nodes, sc.tlas_extent, od, kEps, (double)INFINITY, tlas_stack, leaf, err, wc.kd_splits,
                          sc.tl_cull.node + set, rf);
}

// PT_RENDER_LINEAR_TLAS: the scene WITHOUT its k-d tree — FlatScene as the root, i.e. `[FlatSceneNode]::ray_cast`
// (ray.rs:87-99 over flat_scene.rs:71-99): every instance in list order with one shrinking range [EPSILON, inf).
// The semantic cross-check of the tree walk: no FP32 cull, no tree, O(instances) per ray.
template <bool ANY, bool COUNT>
PT_D bool scene_cast_linear(const DScene& sc, V3 o, V3 d, Hit& hit, KdStack& blas_stack, uint32_t& err, WorkCounters& wc) {
    const RayF rf = make_rayf(o, d);
    const double s = kEps;
    double e = (double)INFINITY;
    bool found = false;
    for (uint32_t inst = 0; inst < sc.n_instances; ++inst) {
        const uint2 pm = __ldg(reinterpret_cast<const uint2*>(&sc.instances[inst].prim));
        if (COUNT) {
            ++wc.instance_tests;
            wc.prim_flops += prim_flop_count(pm.x);
            wc.bbox_gates += (pm.x == PT_PRIM_MESH || pm.x == PT_PRIM_KDMESH) ? 1u : 0u;
            wc.triangle_tests += pm.x == PT_PRIM_TRIANGLE ? 1u : 0u;
        }
        double t;
        uint32_t sub;
        if (instance_t<ANY, COUNT, false>(sc, inst, pm, o, d, rf, s, e, t, sub, blas_stack, err, wc)) {
            e = t;
            hit.t = t; hit.inst = inst; hit.sub = sub;
            found = true;
            if (ANY) return true;
        }
    }
    return found;
}

}  // namespace ptd
