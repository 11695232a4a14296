"""Shared helpers of the parity tests: run the same scene through the CPU oracle
and through the CUDA path (via the C ABI), and compare by the bar of
BASELINE.json: 8-bit RGB within +-1 LSB on >= 99.9 % of pixels, primitive hit
ids bit-exact except grazing-edge pixels with |dt| below DT_EPSILON."""
from __future__ import annotations

import numpy as np

import portrayer_b200 as pt
from oracle import binding as oracle
from portrayer_b200 import _ffi
from portrayer_b200.render import _background_arg, make_params

RGB_TOLERANCE_LSB = 1
RGB_MIN_FRACTION = 0.999
DT_EPSILON = 1e-7  # relative: |t_gpu - t_cpu| <= DT_EPSILON * max(1, |t|) at a hit-id mismatch counts as a grazing edge


def render_oracle(scene: pt.Scene, samples=1, rng="fixed", seed=1, size=None, threads=None, row_stride=1, row_offset=0, **kw):
    """row_stride > 1: the oracle renders only rows row_offset + k * row_stride (a strided sample of a frame that is too
    big to render whole on the CPU); the other rows of the result are untouched (hit ids kNone, t = inf, rgb 0)."""
    w, h = size or (scene.width, scene.height)
    bg, bg_mode = _background_arg(scene, w, h)
    params = make_params(w, h, samples, rng, seed, bg_mode=bg_mode, **kw)
    return oracle.render(scene.blob, scene.camera(w, h), params, bg, threads=threads, row_stride=row_stride, row_offset=row_offset)


def compare_rows(gpu_img: pt.Image, ref: "oracle.OracleResult", rows, label: str = "") -> dict:
    """compare() restricted to the image rows the oracle rendered"""
    class View:
        pass
    g, r = View(), View()
    g.buffer, g.hit_id, g.hit_t = gpu_img.buffer[rows], gpu_img.hit_id[rows], gpu_img.hit_t[rows]
    r.rgb, r.hit_id, r.hit_t = ref.rgb[rows], ref.hit_id[rows], ref.hit_t[rows]
    return compare(g, r, label)


def render_gpu(scene: pt.Scene, samples=1, rng="fixed", seed=1, size=None, dscene=None, **kw):
    w, h = size or (scene.width, scene.height)
    img = pt.Image(w, h)
    stats = img.render(scene, samples=samples, rng=rng, seed=seed, want_hit_ids=True, dscene=dscene, **kw)
    return img, stats


def compare(gpu_img: pt.Image, ref: "oracle.OracleResult", label: str = "") -> dict:
    a = gpu_img.buffer.astype(np.int32)
    b = ref.rgb.astype(np.int32)
    diff = np.abs(a - b).max(axis=2)
    within = float((diff <= RGB_TOLERANCE_LSB).mean())
    exact = float((diff == 0).mean())
    id_mismatch = np.any(gpu_img.hit_id != ref.hit_id, axis=2)
    n_mis = int(id_mismatch.sum())
    grazing = 0
    if n_mis:
        tg, tc = gpu_img.hit_t[id_mismatch], ref.hit_t[id_mismatch]
        fin = np.isfinite(tg) & np.isfinite(tc)
        ok = np.zeros(n_mis, bool)
        ok[fin] = np.abs(tg[fin] - tc[fin]) <= DT_EPSILON * np.maximum(1.0, np.abs(tc[fin]))
        grazing = int(ok.sum())
    t_same = np.array_equal(gpu_img.hit_t, ref.hit_t)
    return {"label": label, "rgb_within_1lsb": within, "rgb_exact": exact, "max_lsb_diff": int(diff.max()),
            "hit_id_mismatches": n_mis, "hit_id_grazing": grazing, "hit_t_bit_identical": t_same}


def assert_parity(report: dict):
    assert report["rgb_within_1lsb"] >= RGB_MIN_FRACTION, report
    assert report["hit_id_mismatches"] == report["hit_id_grazing"], report


def mesh_tie_rays():
    """Axis-aligned rays over the edge-mesh-ties scene (host/examples/kats.cpp) and, by brute force, the triangle the
    reference's index-order fold must return for each: the FIRST listed triangle of the z = -4 layer containing the
    point (all of them return t == 4.0 exactly), else the filler cell behind it (t == 8.0), else a miss."""
    layer = [((8, 8), (12, 8), (8, 12)), ((0, 0), (16, 0), (0, 16)), ((0, 0), (4, 0), (0, 4)), ((2, 2), (6, 2), (2, 6)),
             ((0, 0), (16, 0), (0, 16)), ((1, 1), (3, 1), (1, 3)), ((20, 0), (24, 0), (20, 4)), ((20, 0), (24, 0), (20, 4))]
    xs = np.arange(0.25, 26.0, 0.5)
    ys = np.arange(0.25, 22.0, 0.5)
    gx, gy = np.meshgrid(xs, ys)
    px, py = gx.ravel(), gy.ravel()
    origins = np.stack([px, py, np.zeros_like(px)], axis=1)
    dirs = np.tile(np.array([0.0, 0.0, -1.0]), (len(px), 1))
    expect_sub = np.full(len(px), -1, np.int64)
    expect_t = np.full(len(px), np.inf)
    for i, (x, y) in enumerate(zip(px, py)):
        for k, ((ax, ay), (bx, by), (cx, cy)) in enumerate(layer):  # right triangles with legs along +x / +y from a
            leg = bx - ax
            if x >= ax and y >= ay and (x - ax) + (y - ay) <= leg:
                expect_sub[i], expect_t[i] = k, 4.0
                break
        else:
            cx_, cy_ = int(np.floor(x)), int(np.floor(y))
            if 0 <= cx_ < 20 and 0 <= cy_ < 20 and (x - cx_) + (y - cy_) <= 1.0:
                expect_sub[i], expect_t[i] = 8 + cy_ * 20 + cx_, 8.0
    return origins, dirs, expect_sub, expect_t
