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
            "robot-alarm-clock"}
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
