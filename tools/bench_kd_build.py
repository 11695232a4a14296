#!/usr/bin/env python
"""Measure the device k-d tree build (pt_kd_build) beside the host build it replaces.

    python tools/bench_kd_build.py [--n 1000000] [--kind instances|triangles] [--reps 5]

Prints one JSON line: device time (CUDA events inside the library, best / median of --reps), the host mirror's
time for the same tree (C++ restatement of KDLeaf::partitioned, single thread like the reference), end-to-end time
through the C ABI with host buffers (bounds H2D + build + records D2H), the trees' equality, and the build's HBM
roofline: algorithmic bytes (members x passes, counted by the library) / device time vs MEASURED_PEAKS.json.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--kind", choices=["instances", "triangles"], default="instances")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-host", action="store_true")
    args = ap.parse_args()

    import portrayer_b200 as pt
    from portrayer_b200 import _ffi, kdbuild

    _ffi.check(_ffi.gpu.pt_init(-1))
    depth = math.ceil(math.log2(args.n / 3))
    t0 = time.perf_counter()
    if args.kind == "instances":
        scene = pt.Scene.synthetic_instances(args.n, kd_depth=0)  # flat scene tree: only the bounds are wanted here
        bounds = scene.item_bounds()
    else:
        scene = pt.Scene.synthetic_triangles(args.n, kd_mesh_depth=0)
        bounds = kdbuild.triangle_bounds(scene.section("tri_pos", np.float64, 9))
    scene_s = time.perf_counter() - t0
    cfg = kdbuild.config(depth)

    # the step before: FlatScene::from on the device (pt_flatten) vs the host mirror, same scene graph
    flat_line = None
    if args.kind == "instances":
        from portrayer_b200 import flatten

        hier = flatten.hierarchy_of(scene)
        flatten.FlatScene.build(hier).close()  # warm-up
        f_ms, f_e2e = [], []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            fl = flatten.FlatScene.build(hier)
            f_e2e.append((time.perf_counter() - t0) * 1e3)
            f_ms.append(fl.build_stats()[0])
            n_inst = fl.instance_count
            _inst, _trans, f_bounds = fl.download() if _ == 0 else (None, None, None)
            if _ == 0:
                same_bounds = bool(np.array_equal(f_bounds, bounds))
            fl.close()
        fbytes = len(hier.nodes) * 144 + len(hier.children) * 4 + n_inst * (64 + 128 + 96 + 48)
        flat_line = {"device_ms": min(f_ms), "e2e_ms": min(f_e2e), "host_ms": _ffi.host.pth_flatten_seconds(scene._h) * 1e3,
                     "instances": int(n_inst), "bounds_identical_to_host": same_bounds,
                     "algorithmic_bytes": int(fbytes), "achieved_gbs": fbytes / (min(f_ms) * 1e-3) / 1e9,
                     "path": "pt_flatten (host graph in: H2D of nodes / children / geometries, level-parallel flatten); "
                             "e2e = wall clock of the call"}

    kdbuild.KdTree.build(bounds, cfg).close()  # warm-up: allocator, module load
    dev_ms, e2e_ms = [], []
    tree = None
    for _ in range(args.reps):
        t0 = time.perf_counter()
        tree = kdbuild.KdTree.build(bounds, cfg)
        nodes, items = tree.download()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
        dev_ms.append(tree.build_stats()[0])
        if _ + 1 < args.reps:
            tree.close()
    launches = tree.build_stats()[1]
    abytes = tree.algorithmic_bytes()

    host = None
    same = None
    if not args.skip_host:
        host = kdbuild.host_build(bounds, cfg)
        same = bool(np.array_equal(nodes["a"], host.nodes["a"]) and np.array_equal(nodes["b"], host.nodes["b"]) and
                    np.array_equal(nodes["split"], host.nodes["split"]) and np.array_equal(items, host.items))

    peak, src = 7700.0, "B200_PROFILING.md fallback"
    mp = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        with open(mp) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    traffic = None
    tp = os.path.join(REPO, "profiles", "traffic_kd_build.json")
    if os.path.exists(tp) and args.n == 1_000_000 and args.kind == "instances":  # the capture is of exactly this build
        with open(tp) as f:
            traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for k, v in json.load(f).items() if not k.startswith("_"))
    best = min(dev_ms)
    achieved = abytes / (best * 1e-3) / 1e9
    line = {
        "metric": "k-d tree build time", "unit": "ms", "higher_is_better": False,
        "config": {"workload": f"synthetic {args.kind} n={args.n} (SURVEY 8d M3b), max_depth={depth}, PartitionConfig 3/3/10",
                   "scene_prepare_s": scene_s},
        "value": best, "device_ms_median": statistics.median(dev_ms), "reps": args.reps,
        "tree": {"nodes": int(tree.node_count), "leaf_members": int(tree.item_count), "depth": int(tree.depth)},
        "gpu_launches": int(launches),
        "e2e": {"value": min(e2e_ms), "unit": "ms", "h2d_bytes": int(bounds.nbytes), "d2h_bytes": int(nodes.nbytes + items.nbytes),
                "path": "pt_kd_build (host bounds in) + pt_kd_tree_download (records out), wall clock"},
        "cpu_baseline": None if host is None else {"value": host.seconds * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                                                   "sample": "the whole build, once (C++ mirror of KDLeaf::partitioned)"},
        "identical_to_host_tree": same,
        "flatten": flat_line,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes": int(abytes), "peak_source": src, "traffic": traffic},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
