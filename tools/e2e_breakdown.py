#!/usr/bin/env python
"""Where the end-to-end time of one frame goes: pt_scene_upload vs pt_render (wall), and inside pt_render the
device / copy times PtStats reports.  python tools/e2e_breakdown.py [example ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (pinned host memory)

import portrayer_b200 as pt  # noqa: E402
from portrayer_b200 import _ffi  # noqa: E402
from portrayer_b200.render import _background_arg, make_params  # noqa: E402

names = sys.argv[1:] or ["primitives", "texture-mapping", "normal-mapping"]
_ffi.check(_ffi.gpu.pt_init(0))
for name in names:
    sc = pt.Scene.example(name)
    w, h = sc.width, sc.height
    bg, bg_mode = _background_arg(sc, w, h)
    p = make_params(w, h, int(os.environ.get("SAMPLES", "1")), "hash", 1, bg_mode=bg_mode)
    blob = torch.from_numpy(sc.blob.copy()).pin_memory().numpy()
    bgp = torch.from_numpy(np.ascontiguousarray(bg)).pin_memory().numpy()
    rgb = torch.zeros((h, w, 3), dtype=torch.uint8).pin_memory().numpy()
    cam = sc.camera()
    up, rd, dev, h2d, d2h = [], [], [], [], []
    for it in range(12):
        t0 = time.perf_counter()
        ds = pt.DeviceScene(blob)
        t1 = time.perf_counter()
        st = ds.render(cam, p, bgp, rgb)
        t2 = time.perf_counter()
        ds.close()
        if it >= 2:
            up.append(t1 - t0); rd.append(t2 - t1); dev.append(st.device_ms); h2d.append(st.h2d_ms); d2h.append(st.d2h_ms)
    f = lambda v: f"{1e3 * float(np.median(v)):.3f}"
    print(f"{name:22s} upload {f(up)} ms  render {f(rd)} ms  (device {np.median(dev):.3f} ms, h2d {np.median(h2d):.3f}, d2h {np.median(d2h):.3f})"
          f"  blob {blob.nbytes} B, uploaded {ds.uploaded_bytes if False else 0}")
