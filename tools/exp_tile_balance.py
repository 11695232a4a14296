#!/usr/bin/env python
"""What an N-GPU render of the default workload would take per rank, measured on ONE GPU: rank r of world N renders its
interleaved tiles alone; max over r = the N-GPU frame time, mean = perfect balance, sum vs the N = 1 frame = the price of
scattering a batch's tiles.   python tools/exp_tile_balance.py [world] [samples] [tile sizes ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import portrayer_b200 as pt
from portrayer_b200 import _ffi
from portrayer_b200.render import _background_arg, make_params

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tiles = [int(t) for t in sys.argv[3:]] or [32, 16, 64]
wl = bench.WORKLOADS["castle"]
scene = bench.build_scenes(wl)[0]
w, h = scene.width, scene.height
bg, bg_mode = _background_arg(scene, w, h)
ds = pt.DeviceScene(scene.blob)
flags = _ffi.PT_RENDER_TOLERATE_KD_PLANE


def frame_ms(rank, n, tile):
    p = make_params(w, h, samples, "hash", 1, bg_mode=bg_mode, rank=rank, world=n, tile=tile, flags=flags)
    fr = pt.Frame(ds, scene.camera(w, h), p)
    fr.set_background(np.ascontiguousarray(bg))
    fr.render()
    st = fr.render()
    fr.close()
    return float(st.device_ms), st.rays


one, _ = frame_ms(0, 1, 32)
print(f"N=1: {one:.1f} ms")
for tile in tiles:
    ms, rays = zip(*[frame_ms(r, world, tile) for r in range(world)])
    print(f"tile {tile}: per rank ms {[round(m, 1) for m in ms]}  max {max(ms):.1f}  mean {np.mean(ms):.1f}  max/mean {max(ms) / np.mean(ms):.3f}  "
          f"sum/N1 {sum(ms) / one:.3f}  -> efficiency at N={world}: {one / world / max(ms):.3f}   ray share spread {max(rays) / np.mean(rays):.3f}", flush=True)
