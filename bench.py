#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the per-pixel render loop on N B200s, beside the host-CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--samples S0]

One "step" = one pass of the hot path over the workload's frames (default: BASELINE.json configs[1] —
examples/primitives + texture-mapping + normal-mapping x3 at their native 910x512).  Rays are counted as in
SURVEY §8d: every ray_cast issued against the scene root (primary + shadow + reflection + refraction).

N > 1 is launched by torchrun, one rank per GPU.  The image shards by interleaved tiles with no data-path
collective; scaling is WEAK: a step at N GPUs renders SAMPLES = S0 * N per pixel, so every rank keeps
(pixels / N) * (S0 * N) = pixels * S0 paths per frame and the job's rays grow with N.  The scene blob is
NCCL-broadcast once (reported, not timed); each step ends with the NCCL gather of the RGB8 tiles to rank 0.

Prints ONE JSON line on rank 0 (contract in the task description).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
# stdout carries exactly ONE JSON line (rank 0): NCCL's version banner / debug lines go to stderr instead
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.path.insert(0, REPO)

WORKLOADS = {
    # BASELINE.json configs[1]
    "configs1": dict(frames=["primitives", "texture-mapping", "normal-mapping", "normal-mapping-left", "normal-mapping-right"],
                     samples=1, label="configs[1]: examples/primitives + texture-mapping + normal-mapping x3 @ 910x512"),
    # configs[0]
    "nonhier": dict(frames=["nonhier"], samples=1, label="configs[0]: examples/nonhier @ 256x256"),
    # configs[2]
    "big-scene": dict(frames=["big-scene"], samples=1, label="configs[2]: examples/big-scene @ 1980x1020, KD_DEPTH=10"),
    # configs[3]
    "secondary": dict(frames=["water-glass", "glossy-reflection", "soft-shadows"], samples=16,
                      label="configs[3]: water-glass + glossy-reflection + soft-shadows @ 910x512, SAMPLES=16"),
    # configs[4]: fixed total work, tiles partitioned across the ranks (strong scaling)
    "castle": dict(frames=["graphics-castle"], samples=64, size=(3840, 2160), scaling="strong", tolerate_kd_plane=True,
                   label="configs[4]: examples/graphics-castle @ 3840x2160, SAMPLES=64, KD_DEPTH=10"),
    "castle-hd": dict(frames=["graphics-castle"], samples=4, size=(1920, 1080), scaling="strong", tolerate_kd_plane=True,
                      label="examples/graphics-castle @ native 1920x1080, SAMPLES=4, KD_DEPTH=10"),
    # the two heaviest of the other example programs (linear meshes, area lights, dielectric / glossy materials)
    "monkeys": dict(frames=["monkeys-making-monkeys"], samples=4, label="examples/monkeys-making-monkeys @ 1920x1080, SAMPLES=4"),
    "robot": dict(frames=["robot-alarm-clock"], samples=4, label="examples/robot-alarm-clock @ 1920x1080, SAMPLES=4"),
    # configs[2], synthetic half: kd trees as deep as ceil(log2(N / 3))
    "synthetic-instances-1e5": dict(frames=["synthetic-instances:100000"], samples=1,
                                    label="configs[2]: synthetic 1e5 random instances @ 1980x1020, KD_DEPTH=16"),
    "synthetic-instances-1e6": dict(frames=["synthetic-instances:1000000"], samples=1,
                                    label="configs[2]: synthetic 1e6 random instances @ 1980x1020, KD_DEPTH=19"),
    "synthetic-triangles-1e6": dict(frames=["synthetic-triangles:1000000"], samples=1,
                                    label="configs[2]: one KDMesh of 1e6 random triangles @ 1980x1020, KD_MESH_DEPTH=19"),
}
SEED = 1


def build_scene(name):
    """an examples/*.rs scene program by name, or 'synthetic-instances:N' / 'synthetic-triangles:N'"""
    import math

    import portrayer_b200 as pt

    if ":" in name:
        kind, n = name.split(":")
        n = int(n)
        depth = math.ceil(math.log2(n / 3))
        return pt.Scene.synthetic_instances(n, kd_depth=depth) if kind == "synthetic-instances" else pt.Scene.synthetic_triangles(n, kd_mesh_depth=depth)
    return pt.Scene.example(name)


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def measured_peaks() -> tuple[float, str]:
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (fallback)"


# algorithmic bytes (restated in DESIGN.md §measurement): per traversal kernel, from the work counters
def algorithmic_bytes(kind: int, stats, n_rays: int) -> float:
    s, i, tr, bx = (stats.k_kd_splits[kind], stats.k_instance_tests[kind], stats.k_triangle_tests[kind], stats.k_bbox_gates[kind])
    walk = 16.0 * s + (4 + 96 + 8) * i + 72.0 * tr + 96.0 * bx
    if kind == 0:   # extend: ray record in (48 B), hit record out (t 8 + instance 4 + sub 4)
        return walk + n_rays * (48 + 16)
    # shadow: parent ray + hit in (48 + 16), path meta (8), instance invtrans + trans (192), light (120), occlusion byte out
    return walk + n_rays * (48 + 16 + 8 + 192 + 120 + 1)


def build_workload(args, world):
    import portrayer_b200 as pt

    wl = WORKLOADS[args.workload]
    samples = (args.samples or wl["samples"]) * (1 if wl.get("scaling") == "strong" else world)
    scenes = []
    for name in wl["frames"]:
        sc = build_scene(name)
        if wl.get("size"):
            sc.width, sc.height = wl["size"]
        scenes.append(sc)
    return wl, samples, scenes


def cpu_band_fraction(oracle, make_params, jobs, samples, threads, budget_s):
    """Fraction of every frame's rows (a centred band) the oracle can render within budget_s: calibrated on a
    thin band of the first frame.  jobs: (scene, camera, params, background)."""
    sc, cam, p, bg = jobs[0]
    band = max(1, sc.height // 64)
    y1 = (sc.height - band) // 2
    p_cal = make_params(sc.width, sc.height, samples, "hash", SEED, slice_=(0, y1, sc.width - 1, y1 + band - 1), bg_mode=p.bg_mode)
    t0 = time.perf_counter()
    oracle.render(sc.blob, cam, p_cal, bg, threads=threads)
    est = (time.perf_counter() - t0) * (sc.height / band) * len(jobs)
    return min(1.0, budget_s / max(est, 1e-6))


def band_params(make_params, sc, p, samples, frac):
    if frac >= 1.0:
        return p
    rows = max(1, int(sc.height * frac))
    y1 = (sc.height - rows) // 2
    return make_params(sc.width, sc.height, samples, "hash", SEED, slice_=(0, y1, sc.width - 1, y1 + rows - 1), bg_mode=p.bg_mode)


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path.  The reference is Rust and cannot be compiled in
    this image (no cargo), so this arm times the line-by-line C port (oracle/), on every host thread."""
    if rank != 0:
        return
    import portrayer_b200 as pt  # host mirror only: scene building + packing (no GPU call on this arm)
    from oracle import binding as oracle
    from portrayer_b200.render import _background_arg, make_params

    wl, samples, scenes = build_workload(args, world)
    threads = host_threads()
    jobs = []
    for sc in scenes:
        bg, bg_mode = _background_arg(sc, sc.width, sc.height)
        jobs.append((sc, sc.camera(), make_params(sc.width, sc.height, samples, "hash", SEED, bg_mode=bg_mode), bg))

    # bound the step: calibrate on a 1/16 slice of the first frame, then cut every frame to a row band that
    # keeps the whole run (steps + warmup) within ~2-3 minutes
    frac = cpu_band_fraction(oracle, make_params, jobs, samples, threads, 150.0 / max(1, args.steps + args.warmup))
    sample_desc = "full frames" if frac >= 1.0 else f"centre row band = {frac:.4f} of every frame"

    def one_step():
        rays = 0
        t0 = time.perf_counter()
        for sc, cam, p, bg in jobs:
            pp = band_params(make_params, sc, p, samples, frac)
            res = oracle.render(sc.blob, cam, pp, bg, threads=threads)
            assert res.rc == 0
            rays += res.stats.rays
        return rays, time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    total_rays, total_s = 0, 0.0
    for _ in range(args.steps):
        r, s = one_step()
        total_rays += r
        total_s += s
    value = total_rays / total_s / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": wl.get("scaling", "weak"),
        "vs_baseline": None, "dtype": "f64", "data": "reference example scenes; deterministic hashed jitter",
        "config": {"workload": wl["label"], "samples": samples, "rng": "hash", "seed": SEED,
                   "note": "reference is Rust (no toolchain here): this is the C port of its render loop (oracle/)"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch

    import portrayer_b200 as pt
    from portrayer_b200 import _ffi
    from portrayer_b200 import distributed as ptd
    from portrayer_b200.render import _background_arg, make_params

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: portrayer_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    _ffi.check(_ffi.gpu.pt_init(local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    wl = WORKLOADS[args.workload]
    samples = (args.samples or wl["samples"]) * (1 if wl.get("scaling") == "strong" else world)
    # Every workload runs with the reference's panic semantics except where a workload says otherwise: over the 2e9 rays
    # of the 4K x 64 castle frame rounding does trip the reference's "ray should definitely hit infinite plane" expect
    # (kdtree/node.rs:147,178; README.md:247-248) on a handful of rays; the frame is finished and the event reported.
    wl_flags = _ffi.PT_RENDER_TOLERATE_KD_PLANE if wl.get("tolerate_kd_plane") else 0

    # ---- scene preparation (host side, stays in the reference's own code in the target design): rank 0 only
    scenes = [build_scene(name) for name in wl["frames"]] if rank == 0 else [None] * len(wl["frames"])
    if rank == 0 and wl.get("size"):
        for sc in scenes:
            sc.width, sc.height = wl["size"]
    t_bcast0 = time.perf_counter()
    frames_meta = []
    for i, name in enumerate(wl["frames"]):
        sc = scenes[i]
        blob_dev = ptd.broadcast_blob(sc.blob if rank == 0 else None, dev)  # NCCL over NVLink when world > 1
        if rank == 0:
            bg, bg_mode = _background_arg(sc, sc.width, sc.height)
            meta = [sc.width, sc.height, bg_mode, bytes(sc.camera()), bg]
        else:
            meta = None
        if world > 1:
            box = [meta]
            torch.distributed.broadcast_object_list(box, src=0)
            meta = box[0]
        frames_meta.append((name, blob_dev, meta))
    torch.cuda.synchronize()
    scene_broadcast_ms = (time.perf_counter() - t_bcast0) * 1e3

    class Job:
        pass

    jobs = []
    for name, blob_dev, (w, h, bg_mode, cam_bytes, bg) in frames_meta:
        j = Job()
        j.name, j.w, j.h = name, w, h
        j.blob_dev = blob_dev
        j.dscene = pt.DeviceScene(device_ptr=blob_dev.data_ptr(), nbytes=blob_dev.numel())
        j.cam = pt.PtCamera.from_buffer_copy(cam_bytes)
        j.bg = np.ascontiguousarray(bg)
        j.bg_mode = bg_mode
        j.params = make_params(w, h, samples, "hash", SEED, bg_mode=bg_mode, rank=rank, world=world, flags=wl_flags)
        j.frame = pt.Frame(j.dscene, j.cam, j.params)
        j.frame.set_background(j.bg)
        j.rgb_dev = ptd.device_tensor(j.frame.rgb_device_ptr, (j.frame.owned_pixels, 3), "|u1", local_rank)
        jobs.append(j)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    same_geometry = len({(j.w, j.h) for j in jobs}) == 1
    # The exchange step at N > 1.  "peer" (default): rank 0's full images are mapped into every rank (CUDA IPC, peer
    # access over NVLink) and every rank's resolve kernel stores its tiles straight into them — the exchange is fused
    # into the last kernel of the frame and only a one-element all-reduce (completion signal) is left.  "nccl": compact
    # tiles + NCCL gather + un-tiling on rank 0 (the library-collective baseline).
    peer = None
    if world > 1 and args.exchange == "peer" and same_geometry:
        try:
            peer = ptd.PeerImage(len(jobs), jobs[0].h, jobs[0].w, local_rank, dst=0)
            ok = 1
        except Exception as exc:  # no peer access between some pair of GPUs on this box: every rank falls back together
            print(f"[rank {rank}] peer image unavailable ({exc}); using the NCCL gather", file=sys.stderr)
            ok = 0
        agree = torch.tensor([ok], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(agree, op=torch.distributed.ReduceOp.MIN)
        if int(agree.item()) == 0:
            if peer is not None:
                peer.close_local()
            peer = None
        else:
            for k, j in enumerate(jobs):
                j.frame.set_image_target(peer.ptr(k))
    n_streams = max(1, min(args.streams, len(jobs)))
    side_streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
    all_streams = [stream] + [s_.cuda_stream for s_ in side_streams]

    # which stream a frame goes to: round-robin until the frames' costs are known (the counting pass below measures
    # them), then longest-processing-time-first onto the least loaded stream, so that no stream is left finishing a
    # long chain of frames alone while the others idle
    stream_of = [k % len(all_streams) for k in range(len(jobs))]

    def balance_streams(costs):
        load = [0.0] * len(all_streams)
        for k in sorted(range(len(jobs)), key=lambda i: -costs[i]):
            t = min(range(len(load)), key=lambda i: load[i])
            stream_of[k] = t
            load[t] += costs[k]

    def step(collect=None):
        """one pass over the workload, device-resident; returns device ms (torch events on the launch stream)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        # all frames of the step are enqueued before anything waits for the device, round-robin over n_streams CUDA
        # streams: the frames are independent (own node pool, control block, __constant__ slot), and a 910x512 frame's
        # kernels are only a few waves each, so the tail of one frame's launch overlaps the next frame's
        for side in side_streams:
            side.wait_event(e0)
        for k, j in enumerate(jobs):
            j.frame.enqueue(stream=all_streams[stream_of[k]])
        for side in side_streams:
            ev = torch.cuda.Event()
            ev.record(side)
            torch.cuda.current_stream().wait_event(ev)
        for j in jobs:
            st = j.frame.finish()
            if collect is not None:
                collect.append(st)
        if peer is not None:
            peer.signal_done()  # the tiles are already in rank 0's images: order them before whatever reads the images
        elif world > 1:
            # the exchange step: NCCL gather of every rank's RGB8 tiles + device-side un-tiling on rank 0, one call
            # for all frames of the step that share a geometry (the exchange is latency-bound)
            if same_geometry:
                ptd.gather_images_device([j.rgb_dev for j in jobs], jobs[0].params, dst=0)
            else:
                for j in jobs:
                    ptd.gather_image_device(j.rgb_dev, j.params, dst=0)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # ---- work counters for the roofline (same deterministic workload, counting kernels, outside the timed region)
    counted = []
    for j in jobs:
        pc = make_params(j.w, j.h, samples, "hash", SEED, bg_mode=j.bg_mode, rank=rank, world=world, flags=_ffi.PT_RENDER_COUNTERS | wl_flags)
        fr = pt.Frame(j.dscene, j.cam, pc)
        fr.set_background(j.bg)
        counted.append(fr.render(stream=stream))
        fr.close()

    if args.balance and len(all_streams) > 1:
        balance_streams([max(float(st.device_ms), 1e-6) for st in counted])  # device time of each frame's counting pass
    for _ in range(args.warmup):
        flush.zero_()
        step()
    exchange_verified = None
    if peer is not None:
        # the fused exchange against the library collective, once, outside the timed region: same bytes on rank 0
        ref = ptd.gather_images_device([j.rgb_dev for j in jobs], jobs[0].params, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            got = peer.images().reshape(len(jobs), -1, 3)
            exchange_verified = bool(torch.equal(got, ref))
            if not exchange_verified:
                raise SystemExit("peer-store exchange differs from the NCCL gather")

    # ---- per-kernel launch durations for the roofline: the same frames on the kernel-by-kernel stream path with a
    # CUDA-event pair around every extend / shadow / shade launch (the timed region below replays CUDA graphs, whose
    # kernels cannot be bracketed individually), L2 flushed before every pass
    timed_stats = []
    for j in jobs:
        pk = make_params(j.w, j.h, samples, "hash", SEED, bg_mode=j.bg_mode, rank=rank, world=world, flags=_ffi.PT_RENDER_KERNEL_TIMES | wl_flags)
        fr = pt.Frame(j.dscene, j.cam, pk)
        fr.set_background(j.bg)
        fr.render(stream=stream)
        for _ in range(max(1, min(args.steps, 5))):
            flush.zero_()
            torch.cuda.synchronize()
            timed_stats.append(fr.render(stream=stream))
        fr.close()
    kernel_passes = max(1, min(args.steps, 5))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    step_ms, step_stats = [], []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        torch.cuda.synchronize()
        sts = []
        step_ms.append(step(sts))
        step_stats.append(sts)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    local_ms = sum(step_ms)
    rays_per_step_local = sum(st.rays for st in step_stats[0])
    launches_local = sum(st.kernel_launches for sts in step_stats for st in sts)
    # batches that ran out of node pool inside the timed steps and were redone (their first attempt is in the timed region too)
    overflow_retries = sum(st.retries for sts in step_stats for st in sts)
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    r = torch.tensor([float(rays_per_step_local), float(launches_local)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)  # max over ranks
        torch.distributed.all_reduce(r, op=torch.distributed.ReduceOp.SUM)  # whole-job rays
    total_ms = float(t.item())
    rays_per_step = float(r[0].item())
    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end through the public API, host buffers, copies inside the timed region
    pinned = []
    for j, sc in zip(jobs, scenes):
        blob_host = torch.empty(j.blob_dev.numel(), dtype=torch.uint8).pin_memory()
        blob_host.copy_(j.blob_dev)
        bg_host = torch.from_numpy(j.bg.copy()).pin_memory()
        rgb_host = torch.zeros((j.h, j.w, 3), dtype=torch.uint8).pin_memory()
        pinned.append((blob_host, bg_host, rgb_host))
    pe2e = [make_params(j.w, j.h, samples, "hash", SEED, bg_mode=j.bg_mode, rank=rank, world=world, flags=wl_flags) for j in jobs]

    def e2e_step():
        rays, h2d, d2h = 0, 0, 0
        for j, (blob_host, bg_host, rgb_host), p in zip(jobs, pinned, pe2e):
            ds = pt.DeviceScene(blob_host.numpy())                       # pt_scene_upload: H2D of the scene records (+ textures not yet resident)
            st = ds.render(j.cam, p, bg_host.numpy(), rgb_host.numpy())  # pt_render: H2D background, kernels, D2H image
            rays += st.rays
            h2d += ds.uploaded_bytes + st.h2d_bytes
            d2h += st.d2h_bytes
            ds.close()
        return rays, h2d, d2h

    e2e = None
    if args.device_only:
        pass
    elif world == 1:
        # cold: nothing cached in the library (first render of a program: textures cross PCIe, buffers are allocated)
        for j in jobs:  # the device-resident frames / scenes of the timed region hold texture references: drop them
            j.frame.close()
            j.dscene.close()
        _ffi.gpu.pt_release_cached_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        c_rays, c_h2d, c_d2h = e2e_step()
        torch.cuda.synchronize()
        cold_s = time.perf_counter() - t0
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e_rays = 0
        n_e2e = max(3, min(args.steps, 10))
        for _ in range(n_e2e):
            rr, h2d, d2h = e2e_step()
            e_rays += rr
        torch.cuda.synchronize()
        e_s = time.perf_counter() - t0
        e2e = {"value": e_rays / e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_s / n_e2e * 1e3, "steps": n_e2e,
               "cold_first_step": {"value": c_rays / cold_s / 1e6, "ms": cold_s * 1e3, "h2d_bytes": int(c_h2d), "d2h_bytes": int(c_d2h)},
               "path": "per frame: pt_scene_upload (pinned host blob: scene records every step; texels only when not already "
                       "resident in the library's keyed texture cache) + pt_render (host background in, kernels, RGB8 image "
                       "out to host), wall clock; cold_first_step = same with every cache emptied first"}
    else:
        # N > 1: upload + tile render + NCCL gather + D2H on rank 0, wall clock, max over ranks
        def e2e_step_multi():
            rays, h2d, d2h = 0, 0, 0
            for j, (blob_host, bg_host, rgb_host), p in zip(jobs, pinned, pe2e):
                ds = pt.DeviceScene(blob_host.numpy())   # scene records H2D (texels only when not resident)
                j.frame.rebind(ds, j.cam)                 # the rank's frame (buffers + graph) is kept, as pt_render does
                j.frame.set_background(bg_host.numpy())
                st = j.frame.render(stream=stream)
                if peer is not None:
                    peer.signal_done()  # tiles were stored into rank 0's image by the resolve kernels
                    if rank == 0:
                        rgb_host.copy_(peer.images()[jobs.index(j)], non_blocking=True)  # D2H on rank 0
                        torch.cuda.synchronize()
                        d2h += rgb_host.numel()
                else:
                    img = ptd.gather_image(j.rgb_dev, p, dst=0)  # NCCL gather, device un-tiling, D2H on rank 0
                    if rank == 0:
                        d2h += int(img.nbytes)
                rays += st.rays
                h2d += ds.uploaded_bytes + bg_host.numel() * 8
                j.frame.rebind(j.dscene)
                ds.close()
            return rays, h2d, d2h

        e2e_step_multi()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 10))
        e_rays = 0
        for _ in range(n_e2e):
            rr, h2d, d2h = e2e_step_multi()
            e_rays += rr
        barrier()
        e_s = time.perf_counter() - t0
        tt = torch.tensor([e_s], dtype=torch.float64, device=dev)
        rr_t = torch.tensor([float(e_rays), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(rr_t, op=torch.distributed.ReduceOp.SUM)
        e2e = {"value": float(rr_t[0]) / float(tt) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(rr_t[1]),
               "d2h_bytes_per_step": int(rr_t[2]), "ms_per_step": float(tt) / n_e2e * 1e3, "steps": n_e2e,
               "path": "per frame and rank: scene upload + tile render, " + ("tiles stored into rank 0's image by the resolve kernel "
                       "(peer memory) + completion all-reduce" if peer is not None else "NCCL gather of RGB8 tiles to rank 0") + ", D2H on rank 0"}

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (rank 0's share of the job)
    ms_ext = sum(st.ms_extend for st in timed_stats)
    ms_shd = sum(st.ms_shadow for st in timed_stats)
    ms_sha = sum(st.ms_shade for st in timed_stats)
    n_ext = sum(st.n_extend for st in timed_stats)
    n_shd = sum(st.n_shadow for st in timed_stats)
    kind = 1 if ms_shd >= ms_ext else 0
    kname = "shadow_kernel" if kind == 1 else "extend_kernel"
    bytes_step = 0.0
    for st in counted:
        n_rays = st.rays_shadow if kind == 1 else (st.rays_primary + st.rays_reflect + st.rays_refract)
        bytes_step += algorithmic_bytes(kind, st, n_rays)
    k_ms, k_n = (ms_shd, n_shd) if kind == 1 else (ms_ext, n_ext)
    peak, peak_src = measured_peaks()
    achieved = bytes_step * kernel_passes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(kname, {}).get("dram_bytes_per_launch")
    # second roof (report both, the binding one is the slower): algorithmic f64 flops of the same kernel,
    # 14 per kd split + (42 + P_type) per instance test + 58 per triangle test + 150 per bbox gate, against the
    # f64 issue ceiling measured on this GPU for separate DMUL + DADD (the parity build has no FMA)
    flops_step = 0.0
    for st in counted:
        flops_step += (14.0 * st.k_kd_splits[kind] + 42.0 * st.k_instance_tests[kind] + float(st.k_prim_flops[kind])
                       + 58.0 * st.k_triangle_tests[kind] + 150.0 * st.k_bbox_gates[kind])
    fp64_peak = C.c_double(0.0)
    _ffi.check(_ffi.gpu.pt_measure_fp64_rate(20.0, C.byref(fp64_peak)))
    fp64_achieved = flops_step * kernel_passes / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    roofline_fp64 = {"bound": "fp64 issue (no FMA)", "kernel": kname, "achieved": fp64_achieved, "peak": fp64_peak.value,
                     "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak.value if fp64_peak.value else None,
                     "peak_source": "pt_measure_fp64_rate: DMUL+DADD chains, measured in this run",
                     "algorithmic_flops_per_launch": flops_step * kernel_passes / max(k_n, 1),
                     "note": "flops of the reference's algorithm for the rays of the launch; candidates rejected by the "
                             "FP32 box cull are counted although their f64 work is skipped"}
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_step * kernel_passes / max(k_n, 1),
                "avg_launch_ms": k_ms / max(k_n, 1), "launches_timed": k_n,
                "kernel_ms_per_step": {"extend": ms_ext / kernel_passes, "shadow": ms_shd / kernel_passes, "shade": ms_sha / kernel_passes},
                "kernel_ms_per_level": {"extend": [round(sum(st.ms_extend_level[l] for st in timed_stats) / kernel_passes, 4) for l in range(12)],
                                        "shadow": [round(sum(st.ms_shadow_level[l] for st in timed_stats) / kernel_passes, 4) for l in range(12)]},
                "timing": "CUDA events around every launch of the kernel on the stream path (same frames, same kernels; the "
                          "timed region replays them inside CUDA graphs)",
                "note": "bytes TOUCHED per SURVEY 8d (16/kd split, 108/instance test, 72/triangle, 96/bbox gate + ray records); "
                        "the reference's scenes are KBs, so these are served by L1/L2 and the fraction of HBM peak can pass 1: "
                        "the kernel is bound by f64 issue + latency (roofline_fp64), DRAM traffic is in `traffic`"}

    # ---- CPU baseline: the oracle port on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.device_only:
        from oracle import binding as oracle

        threads = host_threads()
        cjobs = [(sc, j.cam, make_params(j.w, j.h, samples, "hash", SEED, bg_mode=j.bg_mode), j.bg) for j, sc in zip(jobs, scenes)]
        # about 10-25 s of CPU work: whole passes over the workload; a workload too big for that (4K x 64 samples) is
        # cut to a centred row band of every frame
        frac = cpu_band_fraction(oracle, make_params, cjobs, samples, threads, 12.0)
        t0 = time.perf_counter()
        c_rays, frames_done = 0, 0
        while time.perf_counter() - t0 < 10.0:
            for sc, cam, p, bg in cjobs:
                res = oracle.render(sc.blob, cam, band_params(make_params, sc, p, samples, frac), bg, threads=threads)
                c_rays += res.stats.rays
                frames_done += 1
                if time.perf_counter() - t0 > 25.0:
                    break
            if time.perf_counter() - t0 > 25.0:
                break
        c_s = time.perf_counter() - t0
        what = "full frames" if frac >= 1.0 else f"centre row band = {frac:.4f} of every frame"
        cpu = {"value": c_rays / c_s / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
               "sample": f"{frames_done} frame render(s) ({what}) = {frames_done / len(jobs):.2f} pass(es) over the {len(jobs)}-frame workload, {c_s:.1f} s",
               "note": "C port of the reference's render loop (oracle/); the Rust reference cannot be built here"}

    # rays on which the reference's kd walk would have panicked (rank 0's share), with the first location per frame
    panics = [{"frame": j.name, "device_error_bits": st.device_error_bits, "pixel": [st.err_pixel % j.w, st.err_pixel // j.w],
               "sample": st.err_sample, "path": st.err_pathid, "where": st.err_where}
              for j, st in zip(jobs, step_stats[0]) if st.device_error_bits] if wl_flags else None
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": wl.get("scaling", "weak"), "vs_baseline": None,
        "dtype": "f64",
        "data": "reference example scenes (procedural geometry, reference OBJ/texture assets" +
                (", stand-ins: " + "; ".join(f"{k}: {v}" for k, v in pt.assets.STAND_INS.items()) if pt.assets.STAND_INS else "") + ")",
        "config": {"workload": wl["label"], "frames_per_step": len(jobs), "samples": samples, "rng": "hash", "seed": SEED,
                   "rays_per_step": rays_per_step, "ms_per_frame": total_ms / args.steps / len(jobs),
                   "l2": "flushed between timed steps (256 MB write)", "tile": "32x32 interleaved over ranks",
                   "streams": n_streams, "stream_of_frame": list(stream_of),
                   "parallelism": f"tiles x{world}", "scene_broadcast_ms": scene_broadcast_ms,
                   "exchange": None if world == 1 else ("resolve kernel stores tiles into rank 0's image over peer memory (NVLink) + "
                                                        "1-element all-reduce" if peer is not None else "NCCL gather + un-tiling on rank 0"),
                   "exchange_verified_against_nccl_gather": exchange_verified,
                   "node_pool_retry_rounds_in_timed_steps": int(overflow_retries),
                   "reference_panics_tolerated": panics},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(r[1].item()), "roofline": roofline, "roofline_fp64": roofline_fp64,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="configs1")
    ap.add_argument("--samples", type=int, default=0, help="samples per pixel at N=1 (default: the workload's)")
    ap.add_argument("--streams", type=int, default=3, help="CUDA streams the frames of a step are spread over (1 = back to back on one)")
    ap.add_argument("--no-balance", dest="balance", action="store_false",
                    help="keep the frames round-robin over the streams instead of balancing the streams by the frames' measured device times")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="N > 1: how the tiles reach rank 0 (peer: stored by the resolve kernel into rank 0's image over NVLink; nccl: gather)")
    ap.add_argument("--device-only", action="store_true", help="skip the e2e and cpu_baseline legs (for runs under ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        from portrayer_b200 import distributed as ptd

        ptd.init_process_group("nccl")
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
