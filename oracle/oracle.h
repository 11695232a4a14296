/* oracle.h — entry points of the CPU oracle (TEST INFRASTRUCTURE ONLY; see oracle.c). */
#ifndef PORTRAYER_ORACLE_H
#define PORTRAYER_ORACLE_H
#include <stdint.h>

#include "portrayer_gpu.h" /* the boundary's data format: the oracle consumes the same blob as the device */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleStats {
    uint64_t rays_primary, rays_shadow, rays_reflect, rays_refract, rays_depth_cut;
    uint64_t kd_splits, instance_tests, triangle_tests, bbox_gates;
    uint64_t shaded_hits, texel_lookups;
    /* the share of the four work counters above spent inside shadow-ray casts (material.rs:174-179); the rest belongs
     * to primary / reflected / refracted rays — what the device's closest-hit kernel must reproduce exactly */
    uint64_t shadow_kd_splits, shadow_instance_tests, shadow_triangle_tests, shadow_bbox_gates;
} OracleStats;

/* The pixel loop of src/render.rs:127-150 on n_threads host threads.
 * color_out (nullable): W*H*3 post-gamma, clamped, pre-quantisation f64. Returns 0 or a PtError. */
int oracle_render(const void* blob, uint64_t bytes, const PtCamera* cam, const PtRenderParams* params,
                  const double* background, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out,
                  double* color_out, int n_threads, OracleStats* stats);

/* The same over a strided sample of the slice's rows: only rows y1 + row_offset + k * row_stride are rendered (what
 * bench.py times as the host-CPU baseline of a frame too big to render whole). */
int oracle_render_rows(const void* blob, uint64_t bytes, const PtCamera* cam, const PtRenderParams* params,
                       const double* background, uint8_t* rgb_inout, uint32_t* hit_id_out, double* hit_t_out,
                       double* color_out, int n_threads, uint32_t row_stride, uint32_t row_offset, OracleStats* stats);

/* Ray::color(scene, background, 0) (src/ray.rs:139-148) for explicit rays. */
int oracle_trace_rays(const void* blob, uint64_t bytes, uint64_t n, const double* origins, const double* dirs,
                      const double* background3, uint32_t rng_mode, uint64_t seed, uint32_t max_depth, double* color_out,
                      uint32_t* hit_id_out, double* hit_t_out, int n_threads, OracleStats* stats);

/* Camera::ray_at (src/camera.rs:48-84) for n (x, y) pairs. */
void oracle_camera_rays(const PtCamera* cam, uint64_t n, const double* xy, double* origins, double* dirs);

/* Quadratic::solve (src/math.rs:107-114): roots ascending into out[2], returns the count. */
int oracle_solve_quadratic(double a, double b, double c, double* out);

#ifdef __cplusplus
}
#endif
#endif
