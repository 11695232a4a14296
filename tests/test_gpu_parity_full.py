"""Parity of the CUDA path against the CPU oracle at the sizes BASELINE.json's configs STATE (-m gpu).

tests/test_gpu_parity.py runs the same scenes reduced so that the oracle finishes in a second or two; here the device
renders the whole frame at the stated size and sample count and the oracle renders what the host cores manage in a few
seconds: the whole frame where that is cheap (configs[2]), otherwise a STRIDED sample of its rows (RNG draws are keyed
by the global pixel index, so a row of the full frame can be rendered alone).  Everything is compared by the bar of
BASELINE.json (parity.py), and the ray counts of the compared rows must match exactly."""
import numpy as np
import pytest

import parity
import portrayer_b200 as pt
from conftest import has_reference_assets
from portrayer_b200 import _ffi

pytestmark = pytest.mark.gpu


def _strided(scene, label, size, samples, stride, offset=0, flags=0, img=None):
    """whole frame on the device, every stride-th row on the oracle"""
    w, h = size
    if img is None:
        img, _ = parity.render_gpu(scene, samples=samples, rng="hash", size=size, flags=flags)
    ref = parity.render_oracle(scene, samples=samples, rng="hash", size=size, row_stride=stride, row_offset=offset)
    assert ref.rc in (0, _ffi.PT_ERR_KD_PLANE_MISS if flags & _ffi.PT_RENDER_TOLERATE_KD_PLANE else 0)
    rows = np.arange(offset, h, stride)
    rep = parity.compare_rows(img, ref, rows, label)
    rep["rows"] = len(rows)
    print(rep)
    parity.assert_parity(rep)
    # the same rows alone on the device (slice by slice would be slow: one render per row band of 1): ray counts of
    # the strided sample through the device's own slice render of the first and last sampled row
    for y in (int(rows[0]), int(rows[-1])):
        row_img, st = parity.render_gpu(scene, samples=samples, rng="hash", size=size, slice_=(0, y, w - 1, y), flags=flags)
        assert np.array_equal(row_img.buffer[y], img.buffer[y])  # slice render == the row of the whole frame
        one = parity.render_oracle(scene, samples=samples, rng="hash", size=size, slice_=(0, y, w - 1, y))
        assert (st.rays_primary, st.rays_shadow, st.rays_reflect, st.rays_refract, st.rays_depth_cut) == \
               (one.stats.rays_primary, one.stats.rays_shadow, one.stats.rays_reflect, one.stats.rays_refract, one.stats.rays_depth_cut)
    return img, rep


# configs[2]: examples/big-scene at its native 1980 x 1020, whole frame on both sides, default tree depth and a deep tree
@pytest.mark.parametrize("kd_depth", [10, 18])
def test_big_scene_native(gpu_ready, kd_depth):
    scene = pt.Scene.big_scene(10, kd_depth=kd_depth)
    assert (scene.width, scene.height) == (1980, 1020)
    img, stats = parity.render_gpu(scene, samples=1, rng="fixed")
    ref = parity.render_oracle(scene, samples=1, rng="fixed")
    assert ref.rc == 0
    rep = parity.compare(img, ref, f"big-scene native kd{kd_depth}")
    print(rep)
    parity.assert_parity(rep)
    assert rep["hit_t_bit_identical"]
    assert stats.rays == ref.stats.rays


# configs[2], synthetic half at the top of the stated range: 10^6 random instances (scene tree of depth 19) and one
# KDMesh of 10^6 random triangles (depth 19), 1980 x 1020, every 4th row on the oracle
@pytest.mark.parametrize("kind", ["instances", "triangles"])
def test_synthetic_1e6(gpu_ready, kind):
    n, depth = 1_000_000, 19
    scene = pt.Scene.synthetic_instances(n, kd_depth=depth) if kind == "instances" else pt.Scene.synthetic_triangles(n, kd_mesh_depth=depth)
    assert max(scene.header.tlas_depth, scene.header.blas_max_depth) == depth
    _, rep = _strided(scene, f"synthetic {kind} 1e6", (1980, 1020), 1, stride=4, offset=1)
    assert rep["hit_t_bit_identical"]


# configs[3]: water-glass + glossy-reflection + soft-shadows at 910 x 512, SAMPLES=16 (hashed jitter, glossy / area
# light / dielectric RNG dimensions); every 8th row on the oracle (soft-shadows folds two 5 804-triangle linear Meshes
# per candidate ray on the CPU)
@pytest.mark.parametrize("name,stride", [("glossy-reflection", 4), ("soft-shadows", 16), ("water-glass", 4)])
def test_configs3_as_stated(gpu_ready, name, stride):
    if name == "water-glass" and not has_reference_assets():
        pytest.skip("reference textures not synced (tools/sync_assets.py)")
    scene = pt.Scene.example(name)
    _strided(scene, f"{name} 910x512x16", (910, 512), 16, stride=stride, offset=stride // 2)


# configs[4]: graphics-castle at 3840 x 2160, SAMPLES=64 — the whole frame on the device (2.2e9 rays), 9 rows of it on
# the oracle; then the multi-GPU ownership rule at full size: ranks 0, 3 and 7 of world = 8 render their interleaved
# tiles and must reproduce the single-rank frame bit for bit on the pixels they own (and touch nothing else).
@pytest.mark.skipif(not has_reference_assets(), reason="reference assets not synced")
def test_configs4_as_stated(gpu_ready):
    scene = pt.Scene.example("graphics-castle")
    size, samples = (3840, 2160), 64
    flags = _ffi.PT_RENDER_TOLERATE_KD_PLANE  # one ray of this frame trips the reference's kd-plane panic (test_castle_kd_plane_panic_device)
    dscene = pt.DeviceScene(scene.blob)
    try:
        full = pt.Image(*size)
        full.render(scene, samples=samples, rng="hash", want_hit_ids=True, flags=flags, dscene=dscene)
        w, h = size
        stride, offset = 240, 100  # rows 100, 340, ..., 2020: away from the panic pixel's row 1742
        ref = parity.render_oracle(scene, samples=samples, rng="hash", size=size, row_stride=stride, row_offset=offset)
        assert ref.rc == 0
        rows = np.arange(offset, h, stride)
        rep = parity.compare_rows(full, ref, rows, "graphics-castle 3840x2160x64")
        rep["rows"] = len(rows)
        print(rep)
        parity.assert_parity(rep)
        for rank in (0, 3, 7):
            part = pt.Image(*size)
            part.buffer[:] = 7
            part.render(scene, samples=samples, rng="hash", flags=flags, dscene=dscene, rank=rank, world=8)
            ys, xs = np.mgrid[0:h, 0:w]
            tiles_x = (w + 31) // 32
            owned = ((ys // 32) * tiles_x + (xs // 32)) % 8 == rank
            assert np.array_equal(part.buffer[owned], full.buffer[owned])
            assert np.all(part.buffer[~owned] == 7)
    finally:
        dscene.close()
