// f64 vector helpers + the deterministic RNG of the boundary contract.
// Compiled with -fmad=false: every a*b+c below is a separate DMUL and DADD,
// rounded exactly like the reference's (FMA-free) f64 arithmetic.  Sums are
// left-to-right, matching the operation order documented in DESIGN.md.
#pragma once
#include <cstdint>

#include "portrayer_gpu.h"

#define PT_HD __host__ __device__ __forceinline__
#define PT_D __device__ __forceinline__

namespace ptd {

constexpr double kEps = PT_EPSILON;
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct V3 {
    double x, y, z;
};
PT_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
PT_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PT_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PT_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
PT_HD V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
PT_HD V3 operator/(V3 a, double s) { return V3{a.x / s, a.y / s, a.z / s}; }
PT_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
PT_HD double magnitude(V3 a) { return sqrt(dot(a, a)); }
PT_HD V3 normalized(V3 a) { return a / magnitude(a); }
// Ray::at, src/ray.rs:125-127: origin + direction * t
PT_HD V3 ray_at(V3 o, V3 d, double t) { return V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t}; }
// half-open Range<f64>::contains
PT_HD bool in_range(double s, double e, double t) { return s <= t && t < e; }

// Vec3Ext::transformed_point / transformed_direction (src/math.rs:44-52) on a row-major 3x4
PT_HD V3 xf_point(const double* m, V3 p) {
    return V3{m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}
PT_HD V3 xf_dir(const double* m, V3 d) {
    return V3{m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z,
              m[8] * d.x + m[9] * d.y + m[10] * d.z};
}
// normal_trans = invtrans.transposed() (flat_scene.rs:105): direction through the transpose
PT_HD V3 xf_dir_transposed(const double* m, V3 d) {
    return V3{m[0] * d.x + m[4] * d.y + m[8] * d.z, m[1] * d.x + m[5] * d.y + m[9] * d.z,
              m[2] * d.x + m[6] * d.y + m[10] * d.z};
}

// ---- RNG: see the contract in include/portrayer_gpu.h (PtRenderParams / PT_RNG_*)
PT_HD uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
PT_HD double draw(uint32_t rng_mode, uint64_t seed, uint64_t pixel, uint64_t sample, uint64_t path, uint32_t dim) {
    if (rng_mode == PT_RNG_FIXED) return 0.5;
    uint64_t h = mix64(seed ^ (pixel * 0x9E3779B97F4A7C15ull + sample));
    h = mix64(h + ((path << 8) | dim));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

// 128-byte-friendly loads of scene records through the read-only path
PT_D void load_doubles12(const double* __restrict__ src, double* dst) {
    const double2* p = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double2 v = __ldg(p + i);
        dst[2 * i] = v.x;
        dst[2 * i + 1] = v.y;
    }
}

}  // namespace ptd
