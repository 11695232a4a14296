// Shading on the device: rebuild the RayIntersection of the final hit, then
// Material::hit_color's local part (ambient + unshadowed lights) and the child
// rays it spawns.  Texture and normal-map lookups are done in software with the
// reference's exact nearest-texel / wrap / truncation rules — hardware texture
// filtering would not reproduce them.
//
//   hit attributes   sphere.rs:58-96, cube.rs:83-139, plane.rs:39-48, cylinder.rs:63-66,107-109,
//                    cone.rs:99-104,148, triangle.rs:82-138, flat_scene.rs:85-88
//   Material::hit_color   src/material.rs:91-320, refracted_direction :27-48
//   RgbImageBuffer::at    src/texture.rs:104-141, ImageTexture::at :162-168, NormalMap::normal_at :192-221
//   Light / Parallelogram src/light.rs:31-33,51-53,62-70,88-90
#pragma once
#include "device_scene.cuh"
#include "traverse.cuh"

namespace ptd {

struct SurfaceHit {
    V3 hit_point;  // world
    V3 normal;     // world, not normalised
    bool has_uv, has_nmt;
    double u, v;
    double nmt[9];  // row-major 3x3, object space (quirk 3: never rotated into world space)
};

PT_D void mat3_from_cols(double* m, V3 c0, V3 c1, V3 c2) {
    m[0] = c0.x; m[1] = c1.x; m[2] = c2.x;
    m[3] = c0.y; m[4] = c1.y; m[5] = c2.y;
    m[6] = c0.z; m[7] = c1.z; m[8] = c2.z;
}

// the "to_top" basis shared by sphere.rs:80-96 and cube.rs:119-136
PT_D void to_top_basis(V3 hit_point, V3 normal, double* nmt) {
    const V3 to_top = normalized(v3(0.0, 1.0, 0.0) - hit_point);
    if (fabs(to_top.x) < kEps && fabs(to_top.z) < kEps) {
        mat3_from_cols(nmt, v3(1.0, 0.0, 0.0), normal, normal.y > 0.0 ? v3(0.0, 0.0, 1.0) : v3(0.0, 0.0, -1.0));
    } else {
        const V3 horizontal_tangent = cross(to_top, normal);
        const V3 vertical_tangent = cross(normal, horizontal_tangent);
        mat3_from_cols(nmt, horizontal_tangent, normal, vertical_tangent);
    }
}

// world-space hit point of a node: trans * (o_local + d_local * t)  (flat_scene.rs:87), NOT o_world + d_world * t
PT_D V3 world_hit_point(const DScene& sc, uint32_t inst, V3 o, V3 d, double t, V3* local_o, V3* local_d, V3* local_p) {
    double m[12];
    load_doubles12(sc.instances[inst].invtrans, m);
    const V3 lo = xf_point(m, o), ld = xf_dir(m, d);
    const V3 p = ray_at(lo, ld, t);
    if (local_o) { *local_o = lo; *local_d = ld; *local_p = p; }
    double tr[12];
    load_doubles12(sc.instance_trans[inst].trans, tr);
    return xf_point(tr, p);
}

// Rebuild the winning candidate's RayIntersection. need_uv: only when the material samples a texture or normal map.
PT_D void reconstruct_hit(const DScene& sc, uint32_t inst, uint32_t sub, V3 o, V3 d, double t, bool need_uv, SurfaceHit& out) {
    V3 lo, ld, p;
    out.hit_point = world_hit_point(sc, inst, o, d, t, &lo, &ld, &p);
    const PtInstance* rec = sc.instances + inst;
    const uint32_t prim = __ldg(&rec->prim);
    V3 n = v3(0.0, 0.0, 0.0);
    out.has_uv = false;
    out.has_nmt = false;
    switch (prim) {
        case PT_PRIM_SPHERE: {
            n = p;
            if (need_uv) {
                out.u = (kPi + atan2(-p.z, p.x)) / (2.0 * kPi);
                out.v = acos(p.y) / kPi;
                out.has_uv = true;
                to_top_basis(p, n, out.nmt);
                out.has_nmt = true;
            }
            break;
        }
        case PT_PRIM_CUBE: {
            // FACES, cube.rs:46-65
            const double nx[6] = {1.0, -1.0, 0.0, 0.0, 0.0, 0.0};
            const double ny[6] = {0.0, 0.0, 1.0, -1.0, 0.0, 0.0};
            const double nz[6] = {0.0, 0.0, 0.0, 0.0, 1.0, -1.0};
            n = v3(nx[sub], ny[sub], nz[sub]);
            if (need_uv) {
                const double axis_u[6] = {-1.0, 1.0, 1.0, 1.0, 1.0, -1.0};
                const double axis_v[6] = {1.0, 1.0, -1.0, 1.0, 1.0, 1.0};
                const double off_u[6] = {1.0 / 2.0, 0.0, 1.0 / 4.0, 1.0 / 4.0, 1.0 / 4.0, 3.0 / 4.0};
                const double off_v[6] = {1.0 / 3.0, 1.0 / 3.0, 0.0, 2.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0};
                double fu, fv;
                if (sub < 2) { fu = p.z; fv = p.y; }
                else if (sub < 4) { fu = p.x; fv = p.z; }
                else { fu = p.x; fv = p.y; }
                const double norm_u = fu * axis_u[sub] + 0.5;
                const double norm_v = 0.5 - fv * axis_v[sub];
                out.u = norm_u / 4.0 + off_u[sub];
                out.v = norm_v / 3.0 + off_v[sub];
                out.has_uv = true;
                to_top_basis(p, n, out.nmt);
                out.has_nmt = true;
            }
            break;
        }
        case PT_PRIM_PLANE: {
            n = v3(0.0, 1.0, 0.0);
            if (need_uv) {
                out.u = p.x + 0.5;
                out.v = p.z + 0.5;
                out.has_uv = true;
                mat3_from_cols(out.nmt, v3(1.0, 0.0, 0.0), v3(0.0, 1.0, 0.0), v3(0.0, 0.0, 1.0));
                out.has_nmt = true;
            }
            break;
        }
        case PT_PRIM_CYLINDER: {
            if (sub == 0) n = v3(p.x, 0.0, p.z);
            else n = v3(0.0, sub == 1 ? 1.0 : -1.0, 0.0);  // height / height.abs()
            break;
        }
        case PT_PRIM_CONE: {
            if (sub == 0) {
                const V3 tip = v3(0.0, 0.5, 0.0);
                const V3 tangent1 = tip - p;
                const V3 opposite = v3(-p.x, p.y, -p.z);
                const V3 across = opposite - p;
                const V3 tangent2 = cross(tangent1, across);
                n = cross(tangent1, tangent2);
            } else {
                n = v3(0.0, -1.0, 0.0);
            }
            break;
        }
        default: {  // TRIANGLE / MESH / KDMESH: triangle.rs:82-138
            const PtMesh* mesh = sc.meshes + __ldg(&rec->mesh);
            const uint32_t flags = __ldg(&mesh->flags);
            const uint32_t tri = __ldg(&mesh->tri_first) + sub;
            const double* v = reinterpret_cast<const double*>(sc.tri_pos + tri);
            const V3 A = v3(__ldg(v + 0), __ldg(v + 1), __ldg(v + 2));
            const V3 B = v3(__ldg(v + 3), __ldg(v + 4), __ldg(v + 5));
            const V3 C = v3(__ldg(v + 6), __ldg(v + 7), __ldg(v + 8));
            TriBary bary{0.0, 0.0};
            if (flags & (PT_MESH_FLAG_NORMALS | PT_MESH_FLAG_UVS)) {
                // same expressions as the traversal's test -> same beta, gamma
                double tt;
                triangle_t(sc.tri_pos + tri, lo, ld, -(double)INFINITY, (double)INFINITY, tt, &bary);
            }
            if (flags & PT_MESH_FLAG_NORMALS) {
                const double* nn = reinterpret_cast<const double*>(sc.tri_normals + __ldg(&mesh->nrm_first) + sub);
                const V3 na = v3(__ldg(nn + 0), __ldg(nn + 1), __ldg(nn + 2));
                const V3 nb = v3(__ldg(nn + 3), __ldg(nn + 4), __ldg(nn + 5));
                const V3 nc = v3(__ldg(nn + 6), __ldg(nn + 7), __ldg(nn + 8));
                const double alpha = 1.0 - bary.beta - bary.gamma;
                n = na * alpha + nb * bary.beta + nc * bary.gamma;
            } else {
                n = cross(B - A, C - A);
            }
            if ((flags & PT_MESH_FLAG_UVS) && need_uv) {
                const double* uv = reinterpret_cast<const double*>(sc.tri_uvs + __ldg(&mesh->uv_first) + sub);
                const double uau = __ldg(uv + 0), uav = __ldg(uv + 1), ubu = __ldg(uv + 2), ubv = __ldg(uv + 3),
                             ucu = __ldg(uv + 4), ucv = __ldg(uv + 5);
                const double alpha = 1.0 - bary.beta - bary.gamma;
                const double uu = uau * alpha + ubu * bary.beta + ucu * bary.gamma;
                const double vv = uav * alpha + ubv * bary.beta + ucv * bary.gamma;
                out.u = uu;
                out.v = 1.0 - vv;
                out.has_uv = true;
                const V3 edge1 = B - A, edge2 = C - A;
                const double d1u = ubu - uau, d1v = ubv - uav, d2u = ucu - uau, d2v = ucv - uav;
                V3 tangent = v3(d2v * edge1.x - d1v * edge2.x, d2v * edge1.y - d1v * edge2.y, d2v * edge1.z - d1v * edge2.z);
                V3 bitangent = v3(-d2u * edge1.x + d1u * edge2.x, -d2u * edge1.y + d1u * edge2.y, -d2u * edge1.z + d1u * edge2.z);
                const double coeff = d1u * d2v - d2u * d1v;
                tangent = normalized(tangent / coeff);
                bitangent = normalized(bitangent / coeff);
                mat3_from_cols(out.nmt, tangent, normalized(n), bitangent);
                out.has_nmt = true;
            }
            break;
        }
    }
    // hit.normal.transformed_direction(normal_trans), flat_scene.rs:88
    double m[12];
    load_doubles12(rec->invtrans, m);
    out.normal = xf_dir_transposed(m, n);
}

// Rust `f64 as i64`: truncate toward zero, saturate, NaN -> 0 (cvt.rzi.s64.f64 would give 0x8000... for NaN)
PT_D long long f64_as_i64(double v) {
    if (v != v) return 0;
    if (v >= 9223372036854775807.0) return 0x7FFFFFFFFFFFFFFFll;
    if (v <= -9223372036854775808.0) return (long long)0x8000000000000000ull;
    return (long long)v;
}
PT_D long long rem_euclid(long long value, long long rhs) {  // texture.rs:107-119
    const long long r = value % rhs;
    if (r < 0) return rhs < 0 ? r - rhs : r + rhs;
    return r;
}
// RgbImageBuffer::at, texture.rs:104-141
PT_D void texture_at(const DScene& sc, int tex, double u, double v, double* rgb) {
    const TextureDev* t = sc.textures + tex;
    const uint4 rec = __ldg(reinterpret_cast<const uint4*>(t));  // width, height, texel pointer
    const uint2 wh = make_uint2(rec.x, rec.y);
    const uint8_t* texels = reinterpret_cast<const uint8_t*>(((unsigned long long)rec.w << 32) | rec.z);
    const long long width = wh.x, height = wh.y;
    const long long x = f64_as_i64(u * (double)(width - 1));
    const long long y = f64_as_i64(v * (double)(height - 1));
    const unsigned xi = (unsigned)rem_euclid(x, width);
    const unsigned yi = (unsigned)rem_euclid(y, height);
    const uint8_t* p = texels + ((unsigned long long)yi * wh.x + xi) * 3ull;
    rgb[0] = (double)__ldg(p) / 255.0;
    rgb[1] = (double)__ldg(p + 1) / 255.0;
    rgb[2] = (double)__ldg(p + 2) / 255.0;
}

// ImageTexture::at (texture.rs:162-168): the same texel, each channel through pow(c, 2.2).  A channel has 256
// possible values, so the power comes from a table the device filled with its own pow() (gamma_lut_kernel).
PT_D void texture_at_gamma(const DScene& sc, int tex, double u, double v, double* rgb) {
    const TextureDev* t = sc.textures + tex;
    const uint4 rec = __ldg(reinterpret_cast<const uint4*>(t));
    const uint8_t* texels = reinterpret_cast<const uint8_t*>(((unsigned long long)rec.w << 32) | rec.z);
    const long long width = rec.x, height = rec.y;
    const long long x = f64_as_i64(u * (double)(width - 1));
    const long long y = f64_as_i64(v * (double)(height - 1));
    const unsigned xi = (unsigned)rem_euclid(x, width);
    const unsigned yi = (unsigned)rem_euclid(y, height);
    const uint8_t* p = texels + ((unsigned long long)yi * rec.x + xi) * 3ull;
    rgb[0] = __ldg(sc.gamma_lut + __ldg(p));
    rgb[1] = __ldg(sc.gamma_lut + __ldg(p + 1));
    rgb[2] = __ldg(sc.gamma_lut + __ldg(p + 2));
}

// refracted_direction, material.rs:27-48
PT_D bool refracted_direction(V3 ray_dir, V3 normal, double refraction_index, V3& out) {
    const double eta = refraction_index;
    const double eta_outside = 1.00;
    const double ray_dot_norm = dot(ray_dir, normal);
    const double under_sqrt = 1.0 - eta_outside * eta_outside * (1.0 - ray_dot_norm * ray_dot_norm) / (eta * eta);
    if (under_sqrt < 0.0) return false;
    const V3 refracted_dir_1 = ((ray_dir - normal * ray_dot_norm) * eta_outside) / eta;
    const V3 refracted_dir_2 = normal * sqrt(under_sqrt);
    out = refracted_dir_1 - refracted_dir_2;
    return true;
}

// light position for a shaded node: point light, or a sample of the parallelogram (light.rs:62-70,88-90)
PT_D V3 light_sample_position(const PtLight* __restrict__ light, uint32_t l, uint32_t rng_mode, uint64_t seed, uint64_t pixel,
                              uint64_t sample, uint64_t path) {
    const double* q = reinterpret_cast<const double*>(light);
    const V3 position = v3(__ldg(q + 0), __ldg(q + 1), __ldg(q + 2));
    const V3 a = v3(__ldg(q + 9), __ldg(q + 10), __ldg(q + 11));
    const V3 b = v3(__ldg(q + 12), __ldg(q + 13), __ldg(q + 14));
    const bool empty = (a.x == 0.0 && a.y == 0.0 && a.z == 0.0) || (b.x == 0.0 && b.y == 0.0 && b.z == 0.0);
    if (empty) return position;
    const double a_coord = 2.0 * draw(rng_mode, seed, pixel, sample, path, 2 + 2 * l) - 1.0;
    const double b_coord = 2.0 * draw(rng_mode, seed, pixel, sample, path, 3 + 2 * l) - 1.0;
    return position + (a * a_coord + b * b_coord);
}

}  // namespace ptd
