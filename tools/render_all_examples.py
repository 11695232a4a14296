#!/usr/bin/env python
"""Render every one of the reference's 28 example programs through the C ABI at its native size and the
reference's default SAMPLES=100, time it, and (optionally) save the PNGs:

    python tools/render_all_examples.py [--samples 100] [--out gpurun_out/examples] [--save]

Prints one line per example: image size, instances, rays, device ms, end-to-end ms (upload + render + read-back),
Mrays/s — the table kept as profiles/rNN_examples.txt."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import portrayer_b200 as pt  # noqa: E402
from portrayer_b200 import _ffi  # noqa: E402
from test_host_kats import REFERENCE_EXAMPLES  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=100)
ap.add_argument("--out", default="gpurun_out/examples")
ap.add_argument("--save", action="store_true")
args = ap.parse_args()
_ffi.check(_ffi.gpu.pt_init(0))
if args.save:
    os.makedirs(args.out, exist_ok=True)
names = list(REFERENCE_EXAMPLES) + ["normal-mapping-left", "normal-mapping-right"]  # normal-mapping.rs renders three images
print(f"# every example of the reference at native size, SAMPLES={args.samples}, hashed jitter, one B200, through pt_scene_upload + pt_render")
print(f"{'example':32s} {'size':>10s} {'inst':>6s} {'rays':>12s} {'device ms':>10s} {'e2e ms':>9s} {'Mrays/s':>9s}  note")
tot_dev = tot_e2e = tot_rays = 0.0
for name in names:
    t0 = time.perf_counter()
    scene = pt.Scene.example(name)
    t_prep = time.perf_counter() - t0
    img = pt.Image(scene.width, scene.height)
    flags = _ffi.PT_RENDER_TOLERATE_KD_PLANE  # finish the frame where the reference's kd-plane expect would panic; reported below
    img.render(scene, samples=min(args.samples, 2), rng="hash", seed=3, flags=flags)  # warm-up: allocations, textures, graph
    t0 = time.perf_counter()
    st = img.render(scene, samples=args.samples, rng="hash", seed=3, flags=flags)
    e2e = (time.perf_counter() - t0) * 1e3
    rays = st.rays_primary + st.rays_shadow + st.rays_reflect + st.rays_refract
    note = f"host scene prep {t_prep * 1e3:.0f} ms"
    if st.device_error_bits:
        note += f"; reference panic bits 0x{st.device_error_bits:x} tolerated"
    print(f"{name:32s} {scene.width:5d}x{scene.height:<4d} {scene.header.n_instances:6d} {rays:12d} {st.device_ms:10.2f} {e2e:9.2f} {rays / st.device_ms / 1e3:9.1f}  {note}")
    tot_dev += st.device_ms; tot_e2e += e2e; tot_rays += rays
    if args.save:
        from PIL import Image
        Image.fromarray(img.buffer).save(os.path.join(args.out, name + ".png"))
print(f"{'TOTAL':32s} {'':>10s} {'':>6s} {int(tot_rays):12d} {tot_dev:10.2f} {tot_e2e:9.2f} {tot_rays / tot_dev / 1e3:9.1f}")
