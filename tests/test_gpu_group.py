"""Single-process multi-GPU: pt_init_devices + the unchanged one-call render (-m gpu).

On a one-GPU box the group has one member (the API and the replicated-scene code paths still run); on a multi-GPU box
(`gpurun --gpus N`) the same tests fan the tiles over every device.  Both tests run in child processes: re-configuring
the library invalidates every handle the other tests of this session hold."""
import os
import subprocess
import sys

import pytest

from conftest import REPO, has_reference_assets

pytestmark = pytest.mark.gpu


def _device_count() -> int:
    import portrayer_b200._ffi as ffi

    return ffi.gpu.pt_device_count()


# tests/cpp/test_group.cpp: the host mirror's Image::render (src/render.rs:216-223), unchanged, on one device and on the
# group: identical bytes, identical ray counts, slices honoured
def test_image_render_through_a_device_group(gpu_ready):
    exe = os.path.join(REPO, "tests", "cpp", "test_group")
    assert os.path.exists(exe), "tests/cpp/test_group not built (make hosttest)"
    n = min(_device_count(), 8)
    name = "graphics-castle" if has_reference_assets() else "primitives"
    env = dict(os.environ, PORTRAYER_ASSETS=os.path.join(REPO, "assets"))
    # the standalone C++ driver has no image decoder (that stays `image` / Pillow on the host side): decode once here
    pre = subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0, {REPO!r}); import portrayer_b200 as pt; pt.Scene.example({name!r})"],
                         capture_output=True, text=True, env=dict(env, PORTRAYER_WRITE_DECODED="1"), timeout=600)
    assert pre.returncode == 0, pre.stderr[-2000:]
    out = subprocess.run([exe, str(n), name, "2"], capture_output=True, text=True, env=env, timeout=600)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"all group checks passed ({n} device" in out.stdout


_CHILD = r"""
import sys
import numpy as np
sys.path.insert(0, {repo!r}); sys.path.insert(0, {tests!r})
import parity
import portrayer_b200 as pt
from portrayer_b200 import _ffi
n = min(_ffi.gpu.pt_device_count(), 8)
assert pt.init_devices(list(range(n))) == n
for name, samples, rng in (("primitives", 1, "fixed"), ("glossy-reflection", 2, "hash")):
    scene = pt.Scene.example(name)
    size = (scene.width // 2, scene.height // 2)
    img, stats = parity.render_gpu(scene, samples=samples, rng=rng, size=size)   # pt_scene_upload + pt_render: the group path
    ref = parity.render_oracle(scene, samples=samples, rng=rng, size=size)
    rep = parity.compare(img, ref, name)
    print(rep)
    parity.assert_parity(rep)
    assert rep["hit_t_bit_identical"]
    assert (stats.rays_primary, stats.rays_shadow, stats.rays_reflect, stats.rays_refract) == \
           (ref.stats.rays_primary, ref.stats.rays_shadow, ref.stats.rays_reflect, ref.stats.rays_refract)
    # the caller's own rank / world nests inside the group: rank 1 of 3 writes exactly its tiles
    part = pt.Image(*size); part.buffer[:] = 7
    part.render(scene, samples=samples, rng=rng, rank=1, world=3, tile=16)
    ys, xs = np.mgrid[0:size[1], 0:size[0]]
    owned = ((ys // 16) * ((size[0] + 15) // 16) + xs // 16) % 3 == 1
    assert np.array_equal(part.buffer[owned], img.buffer[owned]) and np.all(part.buffer[~owned] == 7)
print("GROUP-OK", n)
"""


# the Python layer over the same path, checked against the oracle (hit ids gathered from every member)
def test_group_render_matches_the_oracle(gpu_ready):
    code = _CHILD.format(repo=REPO, tests=os.path.join(REPO, "tests"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    print(out.stdout, out.stderr[-2000:])
    assert out.returncode == 0, out.stdout + out.stderr[-3000:]
    assert "GROUP-OK" in out.stdout
