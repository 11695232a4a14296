"""The oracle against the reference's OWN output: its published renders (render/*.png upstream), frozen as
box-filtered fixtures by tools/make_reference_goldens.py.  Upstream rendered them at SAMPLES=100 with OS-seeded
jitter, so agreement is statistical — mean absolute error and PSNR of the down-filtered images — not bit-exact.
This is what pins the oracle's scene programs, transform conventions, shading and recursion end to end; the
bit-level pins are the reference's known-answer tests (test_oracle_kats.py, test_host_kats.py).  CPU only."""
import os

import numpy as np
import pytest
from PIL import Image

import parity
import portrayer_b200 as pt
from conftest import has_reference_assets

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_renders")
FACTOR = 7
TEXTURED = {"normal-mapping", "normal-mapping-left", "normal-mapping-right", "water-glass", "transmission-refraction",
            "robot-alarm-clock", "robot-alarm-clock-cyan", "robot-alarm-clock-dark-blue", "robot-alarm-clock-red"}
# (example, samples, max mean-abs-error in LSB, min PSNR dB).  big-scene's random scene needs the re-implemented
# rand-0.7 StdRng (host/rand07.hpp) to reproduce upstream's object placement: agreement there pins it.
CASES = [
    ("primitives", 4, 0.8, 45.0),
    ("primitives-simple", 4, 0.8, 45.0),
    ("smooth-shading", 2, 0.8, 44.0),
    ("glossy-reflection", 8, 0.8, 45.0),
    ("soft-shadows", 8, 0.8, 45.0),
    ("normal-mapping", 4, 1.0, 44.0),
    ("normal-mapping-left", 4, 1.0, 44.0),
    ("normal-mapping-right", 4, 1.0, 44.0),
    # water-glass: the water cylinder's bottom cap is COPLANAR with the table top (y = 0.2, examples/water-glass.rs:90-
    # 113), so which surface a refracted ray meets first there is decided at ULP level by vek's matrix inverse /
    # product order (sources absent: "parity unpinned").  Upstream's image shows z-fighting
    # rings in that patch (about 1.7 % of the image); everywhere else agreement is as tight as the other scenes.
    ("water-glass", 8, 1.2, 33.0),
    ("big-scene", 1, 0.8, 44.0),
    # mirrors (reflectivity 1.0 / 0.9), with_children, three-angle rotated_xzy
    ("entering-the-mirror-dimension", 4, 0.8, 44.0),
    # dielectric glass pane + water cube around textured KDMesh fish, normal-mapped cubes
    ("transmission-refraction", 4, 0.8, 46.0),
    # area light + glossy metal / table at 2 samples against upstream's 100: noisier than the others by construction
    ("robot-alarm-clock", 2, 1.4, 42.0),
    # the three other colour variants upstream published (the example's commented-out diffuse lines, :98-100)
    ("robot-alarm-clock-cyan", 2, 1.45, 42.0),
    ("robot-alarm-clock-dark-blue", 2, 1.4, 42.0),
    ("robot-alarm-clock-red", 2, 1.4, 42.0),
]


def _box_down(img, f):
    h, w = (img.shape[0] // f) * f, (img.shape[1] // f) * f
    return img[:h, :w].astype(np.float64).reshape(h // f, f, w // f, f, 3).mean(axis=(1, 3))


@pytest.mark.parametrize("name,samples,max_mae,min_psnr", CASES)
def test_oracle_matches_published_render(name, samples, max_mae, min_psnr):
    if name in TEXTURED and not has_reference_assets():
        pytest.skip("reference textures not synced (tools/sync_assets.py)")
    golden = np.asarray(Image.open(os.path.join(GOLDEN, f"{name}.png")).convert("RGB")).astype(np.float64)
    scene = pt.Scene.example(name)
    res = parity.render_oracle(scene, samples=samples, rng="hash", seed=11)
    assert res.rc == 0
    ours = _box_down(res.rgb, FACTOR)
    assert ours.shape == golden.shape, (ours.shape, golden.shape)
    err = np.abs(ours - golden)
    mae = float(err.mean())
    mse = float(((ours - golden) ** 2).mean())
    psnr = 10.0 * np.log10(255.0 ** 2 / max(mse, 1e-12))
    print(f"{name}: MAE {mae:.3f} LSB, PSNR {psnr:.2f} dB, p99 {np.percentile(err, 99):.1f}")
    assert mae <= max_mae and psnr >= min_psnr, (name, mae, psnr)


# The CUDA path against the reference's published renders, at the reference's own sample count (SAMPLES defaults to
# 100, render.rs:107-113): the same statistical comparison as above, but with the sampling noise of OUR side gone too,
# so the bounds are tighter than the oracle's low-sample ones.  -m gpu.
GPU_CASES = [
    ("primitives", 0.45, 47.0), ("primitives-simple", 0.45, 47.0), ("smooth-shading", 0.5, 46.0), ("glossy-reflection", 0.5, 47.0),
    ("soft-shadows", 0.5, 47.0), ("normal-mapping", 0.7, 45.0), ("normal-mapping-left", 0.7, 45.0), ("normal-mapping-right", 0.7, 45.0),
    ("water-glass", 1.2, 33.0),  # coplanar cap / table patch, see above
    ("big-scene", 0.5, 46.0), ("entering-the-mirror-dimension", 0.5, 46.0), ("transmission-refraction", 0.6, 46.0),
    ("robot-alarm-clock", 0.8, 44.0), ("robot-alarm-clock-cyan", 0.85, 44.0), ("robot-alarm-clock-dark-blue", 0.8, 44.0),
    ("robot-alarm-clock-red", 0.8, 44.0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,max_mae,min_psnr", GPU_CASES)
def test_device_matches_published_render_at_100_samples(gpu_ready, name, max_mae, min_psnr):
    if name in TEXTURED and not has_reference_assets():
        pytest.skip("reference textures not synced (tools/sync_assets.py)")
    golden = np.asarray(Image.open(os.path.join(GOLDEN, f"{name}.png")).convert("RGB")).astype(np.float64)
    scene = pt.Scene.example(name)
    img = pt.Image(scene.width, scene.height)
    # Over the 1.4e9 rays of transmission-refraction at 100 samples ONE ray trips the reference's own
    # "bug: ray should definitely hit infinite plane" expect (kdtree/node.rs:147,178; the README calls the kd-tree
    # "occasionally buggy"); whether upstream's run met such a ray depends on its OS-seeded jitter.  The frame is
    # finished (that ray treats the split as a miss) and the event stays in the stats.
    from portrayer_b200 import _ffi
    st = img.render(scene, samples=100, rng="hash", seed=11, flags=_ffi.PT_RENDER_TOLERATE_KD_PLANE)
    assert st.device_error_bits == 0 or name == "transmission-refraction", hex(st.device_error_bits)
    ours = _box_down(img.buffer, FACTOR)
    assert ours.shape == golden.shape, (ours.shape, golden.shape)
    err = np.abs(ours - golden)
    mae = float(err.mean())
    psnr = 10.0 * np.log10(255.0 ** 2 / max(float(((ours - golden) ** 2).mean()), 1e-12))
    print(f"{name}: MAE {mae:.3f} LSB, PSNR {psnr:.2f} dB, p99 {np.percentile(err, 99):.1f}")
    assert mae <= max_mae and psnr >= min_psnr, (name, mae, psnr)
