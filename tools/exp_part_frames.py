#!/usr/bin/env python
"""Experiment: one frame rendered as K tile-interleaved PART frames (the multi-GPU ownership rule, rank r of K) on K CUDA
streams of ONE GPU, so that one part's kernel tails are filled by another part's kernels.
    python tools/exp_part_frames.py [workload] [samples] [K ...] [--batch PATHS]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import portrayer_b200 as pt
from portrayer_b200 import _ffi
from portrayer_b200.render import _background_arg, make_params

args = [a for a in sys.argv[1:] if not a.startswith("--")]
batch = 0
for a in sys.argv[1:]:
    if a.startswith("--batch="):
        batch = int(a.split("=")[1])
wl = bench.WORKLOADS[args[0] if args else "castle-hd"]
samples = int(args[1]) if len(args) > 1 else wl["samples"]
ks = [int(k) for k in args[2:]] or [1, 2, 3]
scene = bench.build_scenes(wl)[0]
w, h = scene.width, scene.height
bg, bg_mode = _background_arg(scene, w, h)
dscene = pt.DeviceScene(scene.blob)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flags = _ffi.PT_RENDER_TOLERATE_KD_PLANE
for K in ks:
    frames = []
    for r in range(K):
        kw = dict(max_batch_paths=batch) if batch else {}
        p = make_params(w, h, samples, "hash", 1, bg_mode=bg_mode, rank=r, world=K, flags=flags, **kw)
        fr = pt.Frame(dscene, scene.camera(w, h), p)
        fr.set_background(np.ascontiguousarray(bg))
        frames.append(fr)
    streams = [torch.cuda.Stream() for _ in range(K)]
    ms = []
    for it in range(5):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for fr, s in zip(frames, streams):
            fr.enqueue(stream=s.cuda_stream)
        for s in streams:
            ev = torch.cuda.Event(); ev.record(s); torch.cuda.current_stream().wait_event(ev)
        rays = sum(fr.finish().rays for fr in frames)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    best = sorted(ms[2:])[len(ms[2:]) // 2]
    print(f"K={K} batch={batch or 'default'} ms {[round(m, 2) for m in ms]} median {best:.2f} -> {rays / best / 1e3:.1f} Mrays/s", flush=True)
    for fr in frames:
        fr.close()
