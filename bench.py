#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the per-pixel render loop on N B200s, beside the host-CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--samples S0]

One "step" = one pass of the hot path over the workload's frames.  The default workload is BASELINE.json configs[4],
the configuration `north_star` scales on and the largest one that fits a GPU: examples/graphics-castle at 3840x2160,
SAMPLES=64 — 2.2e9 rays per frame.  Rays are counted as in SURVEY §8d: every ray_cast issued against the scene root
(primary + shadow + reflection + refraction).

N > 1 is launched by torchrun, one rank per GPU.  The image shards by interleaved 32x32 tiles with no data-path
collective.  configs[4] is STRONG scaling (fixed frame, each rank renders 1/N of the tiles); the other workloads are
weak scaling (SAMPLES = S0 * N).  The scene blob is NCCL-broadcast once; every rank's resolve kernel stores its RGB8
tiles straight into rank 0's image over peer memory (NVLink), the step ends with a one-element all-reduce.

Prints ONE JSON line on rank 0 (contract in the task description).  With the default workload at N = 1 the line also
carries `other_configs`: device-timed lines of configs[0] .. configs[3].
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
    # BASELINE.json configs[4]: fixed total work, tiles partitioned across the ranks (strong scaling)
    "castle": dict(frames=["graphics-castle"], samples=64, size=(3840, 2160), scaling="strong", tolerate_kd_plane=True,
                   label="configs[4]: examples/graphics-castle @ 3840x2160, SAMPLES=64, KD_DEPTH=10"),
    "castle-hd": dict(frames=["graphics-castle"], samples=4, size=(1920, 1080), scaling="strong", tolerate_kd_plane=True,
                      label="examples/graphics-castle @ native 1920x1080, SAMPLES=4, KD_DEPTH=10"),
    # configs[0]
    "nonhier": dict(frames=["nonhier"], samples=1, label="configs[0]: examples/nonhier @ 256x256"),
    # configs[1]
    "configs1": dict(frames=["primitives", "texture-mapping", "normal-mapping", "normal-mapping-left", "normal-mapping-right"],
                     samples=1, label="configs[1]: examples/primitives + texture-mapping + normal-mapping x3 @ 910x512"),
    # configs[2]
    "big-scene": dict(frames=["big-scene"], samples=1, label="configs[2]: examples/big-scene @ 1980x1020, KD_DEPTH=10"),
    # configs[3]
    "secondary": dict(frames=["water-glass", "glossy-reflection", "soft-shadows"], samples=16,
                      label="configs[3]: water-glass + glossy-reflection + soft-shadows @ 910x512, SAMPLES=16"),
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
DEFAULT_WORKLOAD = "castle"
TABLE_WORKLOADS = ["nonhier", "configs1", "big-scene", "secondary"]  # configs[0] .. configs[3]: the per-config table
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


def measured_peaks() -> tuple[float, str, float]:
    """(HBM GB/s, where it came from, SM max MHz)"""
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (fallback)", 1965.0


def percentile(xs, q):
    xs = sorted(xs)
    if not xs:
        return None
    k = (len(xs) - 1) * q
    lo, hi = int(k), min(int(k) + 1, len(xs) - 1)
    return xs[lo] + (xs[hi] - xs[lo]) * (k - lo)


def workload_samples(wl, args, world):
    return (args.samples or wl["samples"]) * (1 if wl.get("scaling") == "strong" else world)


def build_scenes(wl):
    scenes = []
    for name in wl["frames"]:
        sc = build_scene(name)
        if wl.get("size"):
            sc.width, sc.height = wl["size"]
        scenes.append(sc)
    return scenes


# ----------------------------------------------------------------------------------------------- the CPU arm
class CpuSampler:
    """The oracle (C port of the reference's render loop) on every host thread, over a STRIDED sample of every frame's
    rows (row k * stride of each frame): unbiased over the picture, unlike a centred band."""

    def __init__(self, jobs, samples):
        from oracle import binding as oracle

        self.oracle = oracle
        self.jobs = jobs  # (scene, camera, params, background)
        self.samples = samples
        self.threads = host_threads()
        self.stride = 1

    def calibrate(self, budget_s: float):
        """row stride such that one pass over the workload takes about budget_s: probed on 8 rows of the first frame at
        one sample per pixel (cost is linear in SAMPLES), so that the probe itself stays cheap on a 4K x 64 frame"""
        from portrayer_b200._ffi import PtRenderParams

        sc, cam, p, bg = self.jobs[0]
        p1 = PtRenderParams.from_buffer_copy(bytes(p))
        p1.samples = 1
        probe = max(1, sc.height // 8)
        t0 = time.perf_counter()
        self.oracle.render(sc.blob, cam, p1, bg, threads=self.threads, row_stride=probe)
        est = (time.perf_counter() - t0) * probe * self.samples * len(self.jobs)
        self.stride = max(1, int(np.ceil(est / max(budget_s, 1e-3))))
        return self.stride

    def one_pass(self):
        rays = 0
        t0 = time.perf_counter()
        for sc, cam, p, bg in self.jobs:
            res = self.oracle.render(sc.blob, cam, p, bg, threads=self.threads, row_stride=self.stride)
            if res.rc not in (0, -5):  # -5: the reference's kd-plane panic, tolerated on the castle (see WORKLOADS)
                raise SystemExit(f"oracle failed: {res.rc}")
            rays += res.stats.rays
        return rays, time.perf_counter() - t0

    def describe(self) -> str:
        return "full frames" if self.stride == 1 else f"every {self.stride}th row of every frame (strided sample)"


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path.  The reference is Rust and cannot be compiled in this image
    (no cargo), so this arm times the line-by-line C port (oracle/, -O3), on every host thread.  The CUDA library is
    not even mapped into this process (PORTRAYER_NO_GPU_LIB)."""
    if rank != 0:
        return
    os.environ["PORTRAYER_NO_GPU_LIB"] = "1"
    import portrayer_b200 as pt  # noqa: F401  host mirror only: scene building + packing
    from portrayer_b200.render import _background_arg, make_params

    wl = WORKLOADS[args.workload]
    samples = workload_samples(wl, args, world)
    scenes = build_scenes(wl)
    jobs = []
    for sc in scenes:
        bg, bg_mode = _background_arg(sc, sc.width, sc.height)
        jobs.append((sc, sc.camera(), make_params(sc.width, sc.height, samples, "hash", SEED, bg_mode=bg_mode), bg))
    cpu = CpuSampler(jobs, samples)
    # bound the step: the whole run (steps + warmup) within about 2 minutes
    cpu.calibrate(120.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        cpu.one_pass()
    total_rays, times = 0, []
    for _ in range(args.steps):
        r, s = cpu.one_pass()
        total_rays += r
        times.append(s)
    total_s = sum(times)
    value = total_rays / total_s / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": wl.get("scaling", "weak"),
        "vs_baseline": None, "dtype": "f64", "data": "reference example scenes; deterministic hashed jitter",
        "config": {"workload": wl["label"], "samples": samples, "rng": "hash", "seed": SEED,
                   "note": "reference is Rust (no toolchain here): this is the C port of its render loop (oracle/, gcc -O3 "
                           "-ffp-contract=off); ms_per_step is the time of the SAMPLE, not of a whole frame"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cpu.threads, "kind": "port", "sample": cpu.describe()},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
class Bench:
    """Everything one workload needs on this rank: scenes (rank 0), device scenes, frames, the peer image."""

    def __init__(self, name, args, rank, world, local_rank, stream):
        import torch

        import portrayer_b200 as pt
        from portrayer_b200 import _ffi
        from portrayer_b200 import distributed as ptd
        from portrayer_b200.render import _background_arg, make_params

        self.torch, self.pt, self.ffi, self.ptd, self.make_params = torch, pt, _ffi, ptd, make_params
        self.name, self.args, self.rank, self.world, self.local_rank, self.stream = name, args, rank, world, local_rank, stream
        self.dev = torch.device("cuda", local_rank)
        wl = self.wl = WORKLOADS[name]
        self.samples = workload_samples(wl, args, world)
        # Every workload runs with the reference's panic semantics except where a workload says otherwise: over the 2e9
        # rays of the 4K x 64 castle frame rounding does trip the reference's "ray should definitely hit infinite plane"
        # expect (kdtree/node.rs:147,178; README.md:247-248) on a handful of rays; the frame is finished and the event reported.
        self.flags = _ffi.PT_RENDER_TOLERATE_KD_PLANE if wl.get("tolerate_kd_plane") else 0

        # ---- scene preparation (host side, stays in the reference's own code in the target design): rank 0 only
        t0 = time.perf_counter()
        self.scenes = build_scenes(wl) if rank == 0 else [None] * len(wl["frames"])
        self.scene_prepare_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        metas = []
        for i in range(len(wl["frames"])):
            sc = self.scenes[i]
            blob_dev = ptd.broadcast_blob(sc.blob if rank == 0 else None, self.dev)  # H2D on rank 0, NCCL over NVLink when world > 1
            if rank == 0:
                bg, bg_mode = _background_arg(sc, sc.width, sc.height)
                meta = [sc.width, sc.height, bg_mode, bytes(sc.camera()), bg]
            else:
                meta = None
            if world > 1:
                box = [meta]
                torch.distributed.broadcast_object_list(box, src=0)
                meta = box[0]
            metas.append((blob_dev, meta))
        torch.cuda.synchronize()
        self.scene_broadcast_ms = (time.perf_counter() - t0) * 1e3

        class Job:
            pass

        self.jobs = []
        for name_i, (blob_dev, (w, h, bg_mode, cam_bytes, bg)) in zip(wl["frames"], metas):
            j = Job()
            j.name, j.w, j.h = name_i, w, h
            j.blob_dev = blob_dev
            j.dscene = pt.DeviceScene(device_ptr=blob_dev.data_ptr(), nbytes=blob_dev.numel())
            j.cam = pt.PtCamera.from_buffer_copy(cam_bytes)
            j.bg = np.ascontiguousarray(bg)
            j.bg_mode = bg_mode
            j.params = self.params_of(j)
            j.frame = pt.Frame(j.dscene, j.cam, j.params)
            j.frame.set_background(j.bg)
            j.rgb_dev = ptd.device_tensor(j.frame.rgb_device_ptr, (j.frame.owned_pixels, 3), "|u1", local_rank)
            self.jobs.append(j)
        self.same_geometry = len({(j.w, j.h) for j in self.jobs}) == 1

        # The exchange step at N > 1.  "peer" (default): rank 0's full images are mapped into every rank (CUDA IPC, peer
        # access over NVLink) and every rank's resolve kernel stores its tiles straight into them — the exchange is fused
        # into the last kernel of the frame and only a one-element all-reduce (completion signal) is left.  "nccl": compact
        # tiles + NCCL gather + un-tiling on rank 0 (the library-collective baseline).
        self.peer = None
        t0 = time.perf_counter()
        if world > 1 and args.exchange == "peer" and self.same_geometry:
            try:
                self.peer = ptd.PeerImage(len(self.jobs), self.jobs[0].h, self.jobs[0].w, local_rank, dst=0)
                ok = 1
            except Exception as exc:  # no peer access between some pair of GPUs on this box: every rank falls back together
                print(f"[rank {rank}] peer image unavailable ({exc}); using the NCCL gather", file=sys.stderr)
                ok = 0
            agree = torch.tensor([ok], dtype=torch.int32, device=self.dev)
            torch.distributed.all_reduce(agree, op=torch.distributed.ReduceOp.MIN)
            if int(agree.item()) == 0:
                if self.peer is not None:
                    self.peer.close_local()
                self.peer = None
            else:
                for k, j in enumerate(self.jobs):
                    j.frame.set_image_target(self.peer.ptr(k))
        torch.cuda.synchronize()
        self.peer_setup_ms = (time.perf_counter() - t0) * 1e3

        n_streams = max(1, min(args.streams, len(self.jobs)))
        self.side_streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_streams - 1)]
        self.all_streams = [stream] + [s_.cuda_stream for s_ in self.side_streams]
        # which stream a frame goes to: round-robin until the frames' costs are known (the counting pass measures them),
        # then longest-processing-time-first onto the least loaded stream
        self.stream_of = [k % len(self.all_streams) for k in range(len(self.jobs))]

    def params_of(self, j, flags=0):
        return self.make_params(j.w, j.h, self.samples, "hash", SEED, bg_mode=j.bg_mode, rank=self.rank, world=self.world,
                                flags=flags | self.flags)

    def balance_streams(self, costs):
        load = [0.0] * len(self.all_streams)
        for k in sorted(range(len(self.jobs)), key=lambda i: -costs[i]):
            t = min(range(len(load)), key=lambda i: load[i])
            self.stream_of[k] = t
            load[t] += costs[k]

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def step(self, collect=None):
        """one pass over the workload, device-resident; returns device ms (torch events on the launch stream)"""
        torch, ptd = self.torch, self.ptd
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        # all frames of the step are enqueued before anything waits for the device, over n_streams CUDA streams: the
        # frames are independent (own node pool, control block, __constant__ slot)
        for side in self.side_streams:
            side.wait_event(e0)
        for k, j in enumerate(self.jobs):
            j.frame.enqueue(stream=self.all_streams[self.stream_of[k]])
        for side in self.side_streams:
            ev = torch.cuda.Event()
            ev.record(side)
            torch.cuda.current_stream().wait_event(ev)
        for j in self.jobs:
            st = j.frame.finish()
            if collect is not None:
                collect.append(st)
        if self.peer is not None:
            self.peer.signal_done()  # the tiles are already in rank 0's images: order them before whatever reads the images
        elif self.world > 1:
            # the exchange step: NCCL gather of every rank's RGB8 tiles + device-side un-tiling on rank 0
            if self.same_geometry:
                ptd.gather_images_device([j.rgb_dev for j in self.jobs], self.jobs[0].params, dst=0)
            else:
                for j in self.jobs:
                    ptd.gather_image_device(j.rgb_dev, j.params, dst=0)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def counting_pass(self):
        """work counters of the same deterministic workload (counting kernels), outside any timed region"""
        counted = []
        for j in self.jobs:
            fr = self.pt.Frame(j.dscene, j.cam, self.params_of(j, self.ffi.PT_RENDER_COUNTERS))
            fr.set_background(j.bg)
            counted.append(fr.render(stream=self.stream))
            fr.close()
        return counted

    def kernel_time_pass(self, flush, passes):
        """per-kernel launch durations: the same frames on the kernel-by-kernel stream path with a CUDA-event pair around
        every extend / shadow / shade launch (the timed region replays CUDA graphs, whose kernels cannot be bracketed
        individually), L2 flushed before every pass"""
        out = []
        for j in self.jobs:
            fr = self.pt.Frame(j.dscene, j.cam, self.params_of(j, self.ffi.PT_RENDER_KERNEL_TIMES))
            fr.set_background(j.bg)
            fr.render(stream=self.stream)
            for _ in range(passes):
                flush.zero_()
                self.torch.cuda.synchronize()
                out.append(fr.render(stream=self.stream))
            fr.close()
        return out

    def timed(self, flush, steps, warmup):
        """W warm-up steps, then K timed steps with the L2 flushed before each; per-step device ms, max over ranks"""
        torch = self.torch
        for _ in range(warmup):
            flush.zero_()
            self.step()
        self.barrier()
        step_ms, step_stats = [], []
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (not timed)
            torch.cuda.synchronize()
            sts = []
            step_ms.append(self.step(sts))
            step_stats.append(sts)
        self.barrier()
        t = torch.tensor(step_ms, dtype=torch.float64, device=self.dev)
        rays_local = sum(st.rays for st in step_stats[0])
        launches_local = sum(st.kernel_launches for sts in step_stats for st in sts)
        r = torch.tensor([float(rays_local), float(launches_local)], dtype=torch.float64, device=self.dev)
        # per rank: mean step time on ITS device and the rays it traced (what limits strong scaling: the slowest rank's share)
        own_ms = [sum(float(st.device_ms) for st in sts) for sts in step_stats]  # the rank's own frames (CUDA events), not the step with its completion signal
        mine = torch.tensor([sum(own_ms) / max(len(own_ms), 1), float(rays_local)], dtype=torch.float64, device=self.dev)
        self.by_rank = [[float(mine[0].item()), float(mine[1].item())]]
        if self.world > 1:
            every = [torch.zeros_like(mine) for _ in range(self.world)]
            torch.distributed.all_gather(every, mine)
            self.by_rank = [[float(x[0].item()), float(x[1].item())] for x in every]
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)  # every step: max over ranks
            torch.distributed.all_reduce(r, op=torch.distributed.ReduceOp.SUM)  # whole-job rays
        return [float(x) for x in t.tolist()], float(r[0].item()), int(r[1].item()), step_stats

    def close(self):
        for j in self.jobs:
            j.frame.close()
            j.dscene.close()
        if self.peer is not None:
            self.peer.close()
            self.peer = None


def quick_line(name, args, rank, world, local_rank, stream, flush):
    """device-timed line of another config for the table (N = 1): 3 warm-up + 5 timed steps"""
    b = Bench(name, args, rank, world, local_rank, stream)
    counted = b.counting_pass()
    if args.balance and len(b.all_streams) > 1:
        b.balance_streams([max(float(st.device_ms), 1e-6) for st in counted])
    ms, rays, launches, _ = b.timed(flush, 5, 3)
    b.close()
    return {"workload": b.wl["label"], "samples": b.samples, "value": rays * len(ms) / (sum(ms) * 1e-3) / 1e6, "unit": "Mrays/s",
            "ms_per_step": sum(ms) / len(ms), "ms_per_step_median": statistics.median(ms), "rays_per_step": rays, "steps": len(ms),
            "warmup": 3, "gpu_launches": launches}


def run_ours(args, rank, world, local_rank):
    import torch

    import portrayer_b200 as pt
    from portrayer_b200 import _ffi
    from portrayer_b200 import distributed as ptd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: portrayer_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    _ffi.check(_ffi.gpu.pt_init(local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    b = Bench(args.workload, args, rank, world, local_rank, stream)
    wl, jobs, scenes, samples, peer = b.wl, b.jobs, b.scenes, b.samples, b.peer

    counted = b.counting_pass()
    if args.balance and len(b.all_streams) > 1:
        b.balance_streams([max(float(st.device_ms), 1e-6) for st in counted])  # device time of each frame's counting pass

    exchange_verified = None
    if peer is not None:
        # the fused exchange against the library collective, once, outside the timed region: same bytes on rank 0
        flush.zero_()
        b.step()
        ref = ptd.gather_images_device([j.rgb_dev for j in jobs], jobs[0].params, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            got = peer.images().reshape(len(jobs), -1, 3)
            exchange_verified = bool(torch.equal(got, ref))
            if not exchange_verified:
                raise SystemExit("peer-store exchange differs from the NCCL gather")

    # a pass that takes seconds is not repeated five times for the per-kernel durations
    heavy = sum(float(st.device_ms) for st in counted) > 400.0
    kernel_passes = 1 if heavy else max(1, min(args.steps, 5))
    timed_stats = b.kernel_time_pass(flush, kernel_passes)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, rays_per_step, launches, step_stats = b.timed(flush, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(step_ms)
    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
    overflow_retries = sum(st.retries for sts in step_stats for st in sts)

    # ---- end to end through the public API, host buffers, copies inside the timed region
    pinned = []
    for j in jobs:
        blob_host = torch.empty(j.blob_dev.numel(), dtype=torch.uint8).pin_memory()
        blob_host.copy_(j.blob_dev)
        bg_host = torch.from_numpy(j.bg.copy()).pin_memory()
        rgb_host = torch.zeros((j.h, j.w, 3), dtype=torch.uint8).pin_memory()
        pinned.append((blob_host, bg_host, rgb_host))
    pe2e = [b.params_of(j) for j in jobs]
    n_e2e = 2 if heavy else max(3, min(args.steps, 10))

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

    def e2e_step_multi():
        """N > 1: rank 0 holds the scene in host memory; per frame: H2D on rank 0 + NCCL broadcast of the blob to every
        rank, pt_scene_upload_device, tile render with the tiles stored into rank 0's image (or NCCL gather), D2H on rank 0"""
        rays, h2d, d2h = 0, 0, 0
        for k, (j, (blob_host, bg_host, rgb_host)) in enumerate(zip(jobs, pinned)):
            blob_dev = ptd.broadcast_blob(blob_host.numpy() if rank == 0 else None, dev)
            ds = pt.DeviceScene(device_ptr=blob_dev.data_ptr(), nbytes=blob_dev.numel())
            j.frame.rebind(ds, j.cam)                 # the rank's frame (buffers + graph) is kept, as pt_render does
            j.frame.set_background(bg_host.numpy())
            st = j.frame.render(stream=stream)
            if peer is not None:
                peer.signal_done()
                if rank == 0:
                    rgb_host.copy_(peer.images()[k], non_blocking=True)  # D2H on rank 0
                    torch.cuda.synchronize()
                    d2h += rgb_host.numel()
            else:
                img = ptd.gather_image(j.rgb_dev, j.params, dst=0)  # NCCL gather, device un-tiling, D2H on rank 0
                if rank == 0:
                    d2h += int(img.nbytes)
            rays += st.rays
            h2d += (blob_host.numel() if rank == 0 else 0) + bg_host.numel() * 8
            j.frame.rebind(j.dscene)
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
        if not heavy:
            e2e_step()
        torch.cuda.synchronize()
        e_rays, e_times = 0, []
        for _ in range(n_e2e):
            t0 = time.perf_counter()
            rr, h2d, d2h = e2e_step()
            torch.cuda.synchronize()
            e_times.append(time.perf_counter() - t0)
            e_rays += rr
        e_s = sum(e_times)
        e2e = {"value": e_rays / e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_s / n_e2e * 1e3, "steps": n_e2e, "fraction_of_device_value": e_rays / e_s / 1e6 / value,
               "cold_first_step": {"value": c_rays / cold_s / 1e6, "ms": cold_s * 1e3, "h2d_bytes": int(c_h2d), "d2h_bytes": int(c_d2h)},
               "path": "per frame: pt_scene_upload (pinned host blob: scene records every step; texels only when not already "
                       "resident in the library's keyed texture cache) + pt_render (host background in, kernels, RGB8 image "
                       "out to host), wall clock; cold_first_step = same with every library cache emptied first"}
    else:
        e2e_step_multi()
        b.barrier()
        t0 = time.perf_counter()
        e_rays = 0
        for _ in range(n_e2e):
            rr, h2d, d2h = e2e_step_multi()
            e_rays += rr
        b.barrier()
        e_s = time.perf_counter() - t0
        tt = torch.tensor([e_s], dtype=torch.float64, device=dev)
        rr_t = torch.tensor([float(e_rays), float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(rr_t, op=torch.distributed.ReduceOp.SUM)
        e_val = float(rr_t[0]) / float(tt) / 1e6
        e2e = {"value": e_val, "unit": "Mrays/s", "h2d_bytes_per_step": int(rr_t[1]),
               "d2h_bytes_per_step": int(rr_t[2]), "ms_per_step": float(tt) / n_e2e * 1e3, "steps": n_e2e,
               "fraction_of_device_value": e_val / value,
               "path": "per frame: scene blob H2D on rank 0 + NCCL broadcast to every rank + pt_scene_upload_device, tile render, " +
                       ("tiles stored into rank 0's image by the resolve kernel (peer memory) + completion all-reduce"
                        if peer is not None else "NCCL gather of RGB8 tiles to rank 0") + ", D2H on rank 0; wall clock, max over ranks",
               "one_time_setup_ms": {"scene_broadcast_first": b.scene_broadcast_ms, "peer_image_setup": b.peer_setup_ms}}

    # ---- the other configs (device-timed), N = 1 only
    other = None
    if world == 1 and args.table and args.workload == DEFAULT_WORKLOAD and not args.device_only:
        other = []
        for name in TABLE_WORKLOADS:
            try:
                other.append(quick_line(name, args, rank, world, local_rank, stream, flush))
            except Exception as exc:  # a missing asset must not cost the headline
                other.append({"workload": WORKLOADS[name]["label"], "error": str(exc)})

    if rank != 0:
        return

    # ---- rooflines of the dominant traversal kernel (rank 0's share of the job)
    ms_ext = sum(st.ms_extend for st in timed_stats)
    ms_shd = sum(st.ms_shadow for st in timed_stats)
    ms_sha = sum(st.ms_shade for st in timed_stats)
    n_ext = sum(st.n_extend for st in timed_stats)
    n_shd = sum(st.n_shadow for st in timed_stats)
    kind = 1 if ms_shd >= ms_ext else 0
    kname = "shadow_kernel" if kind == 1 else "extend_kernel"
    k_ms, k_n = (ms_shd, n_shd) if kind == 1 else (ms_ext, n_ext)
    launches_per_pass = max(k_n, 1) / kernel_passes
    avg_launch_s = k_ms * 1e-3 / max(k_n, 1)
    peak, peak_src, sm_max_mhz = measured_peaks()
    # (1) HBM: the bytes one launch HAS to move — the rays' records in and the hit records out, plus every scene byte the
    # walk can touch fetched once (k-d nodes and leaf lists, cull boxes, instance / mesh / triangle records).  Everything
    # else the walk reads is served by L1 / L2: the reference's scenes are a few MB.
    scene_bytes = 0
    for sc in scenes:
        hh = sc.header
        scene_bytes += (16 * (hh.n_tlas_nodes + hh.n_blas_nodes) + 4 * (hh.n_tlas_items + hh.n_blas_items) + 224 * hh.n_instances
                        + 160 * hh.n_meshes + 72 * hh.n_triangles + 32 * (hh.n_tlas_items + hh.n_blas_items + hh.n_instances + hh.n_triangles))
    rays_k = sum((st.rays_shadow if kind == 1 else st.rays_primary + st.rays_reflect + st.rays_refract) for st in counted)
    per_ray = (48 + 16 + 8 + 1) if kind == 1 else (48 + 16)
    bytes_per_launch = rays_k * per_ray / launches_per_pass + min(scene_bytes / len(scenes), 126e6)
    achieved = bytes_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
    prof = {}
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            prof = json.load(f)
    pk = prof.get(kname, {})
    traffic = pk.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                "algorithmic_bytes": f"{per_ray} B per ray of the launch (ray record in, hit record out) + the scene records the walk can touch, once",
                "avg_launch_ms": avg_launch_s * 1e3, "launches_timed": k_n,
                "kernel_ms_per_step": {"extend": ms_ext / kernel_passes, "shadow": ms_shd / kernel_passes, "shade": ms_sha / kernel_passes},
                "kernel_ms_per_level": {"extend": [round(sum(st.ms_extend_level[l] for st in timed_stats) / kernel_passes, 4) for l in range(12)],
                                        "shadow": [round(sum(st.ms_shadow_level[l] for st in timed_stats) / kernel_passes, 4) for l in range(12)]},
                "timing": "CUDA events around every launch of the kernel on the stream path (same frames, same kernels; the "
                          "timed region replays them inside CUDA graphs)",
                "traffic_source": prof.get("_source"),
                "note": "the kernel is NOT bound by HBM (frac is a few percent, and ncu's dram bytes agree): the walk's data is "
                        "L1 / L2 resident; the binding roofs are f64 issue and total issue slots (roofline_fp64, roofline_issue)"}
    # (2) f64 issue: the f64 operations the device EXECUTED (counters of the counting kernels: kd splits walked, exact
    # instance / triangle tests and bbox gates that survived the FP32 cull), against the f64 issue ceiling measured on
    # this GPU for separate DMUL + DADD (the parity build has no FMA)
    flops = sum(14.0 * st.k_kd_splits[kind] + 42.0 * st.x_instance_tests[kind] + float(st.x_prim_flops[kind])
                + 58.0 * st.x_triangle_tests[kind] + 150.0 * st.x_bbox_gates[kind] for st in counted)
    fp64_peak = C.c_double(0.0)
    _ffi.check(_ffi.gpu.pt_measure_fp64_rate(20.0, C.byref(fp64_peak)))
    flops_per_launch = flops / launches_per_pass
    fp64_achieved = flops_per_launch / avg_launch_s / 1e12 if avg_launch_s > 0 else 0.0
    # the same roof as ncu saw it: FP64-pipe warp instructions of the captured launches against one per two cycles per scheduler
    ncu_fp64_frac = None
    if pk.get("fp64_inst_per_launch") and pk.get("avg_us_under_ncu"):
        sms_ = torch.cuda.get_device_properties(local_rank).multi_processor_count
        ncu_fp64_frac = pk["fp64_inst_per_launch"] / (pk["avg_us_under_ncu"] * 1e-6) / (sms_ * 4 * sm_max_mhz * 1e6 * 0.5)
    roofline_fp64 = {"bound": "fp64 issue (no FMA)", "kernel": kname, "achieved": fp64_achieved, "peak": fp64_peak.value,
                     "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak.value if fp64_peak.value else None,
                     "peak_source": "pt_measure_fp64_rate: DMUL+DADD chains, measured in this run",
                     "executed_flops_per_launch": flops_per_launch,
                     "ncu_fp64_pipe_active_frac": ncu_fp64_frac,
                     "note": "f64 operations executed: 14 per kd split walked + (42 + primitive) per exact instance test + 58 per "
                             "exact triangle test + 150 per bbox gate; tests the FP32 box cull rejected are not counted"}
    # (3) issue slots: warp instructions per launch from the ncu capture of the same kernels (profiles/), against
    # SMs x 4 schedulers x SM clock
    box_tests = sum(st.x_box_tests[kind] for st in counted)
    roofline_issue = None
    if pk.get("inst_per_launch") and pk.get("avg_us_under_ncu"):
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        slots_per_s = sms * 4 * sm_max_mhz * 1e6
        ncu_rate = pk["inst_per_launch"] / (pk["avg_us_under_ncu"] * 1e-6)
        roofline_issue = {"bound": "issue slots", "kernel": kname, "achieved": ncu_rate / 1e9, "peak": slots_per_s / 1e9,
                          "unit": "G warp-instructions/s", "frac": ncu_rate / slots_per_s,
                          "lanes_active_per_instruction": pk.get("lanes_per_inst"),
                          "source": "smsp__inst_executed.sum / gpu__time_duration of the captured launches (" + str(prof.get("_source")) + ")"}
    executed = {"kernel": kname,
                "per_ray": {"kd_splits": sum(st.k_kd_splits[kind] for st in counted) / max(rays_k, 1),
                            "fp32_box_tests": box_tests / max(rays_k, 1),
                            "exact_instance_tests": sum(st.x_instance_tests[kind] for st in counted) / max(rays_k, 1),
                            "exact_triangle_tests": sum(st.x_triangle_tests[kind] for st in counted) / max(rays_k, 1),
                            "reference_instance_tests": sum(st.k_instance_tests[kind] for st in counted) / max(rays_k, 1),
                            "reference_triangle_tests": sum(st.k_triangle_tests[kind] for st in counted) / max(rays_k, 1)}}

    # ---- CPU baseline: the oracle port on the host cores, bounded strided sample
    cpu = None
    if world == 1 and not args.device_only:
        from portrayer_b200.render import make_params

        cjobs = [(sc, j.cam, make_params(j.w, j.h, samples, "hash", SEED, bg_mode=j.bg_mode), j.bg) for j, sc in zip(jobs, scenes)]
        sampler_cpu = CpuSampler(cjobs, samples)
        sampler_cpu.calibrate(12.0)
        t0 = time.perf_counter()
        c_rays, passes_done = 0, 0
        while time.perf_counter() - t0 < 10.0:
            rr, _ = sampler_cpu.one_pass()
            c_rays += rr
            passes_done += 1
            if time.perf_counter() - t0 > 25.0:
                break
        c_s = time.perf_counter() - t0
        cpu = {"value": c_rays / c_s / 1e6, "unit": "Mrays/s", "cores": sampler_cpu.threads, "kind": "port",
               "sample": f"{passes_done} pass(es) over the {len(jobs)}-frame workload, {sampler_cpu.describe()}, {c_s:.1f} s",
               "note": "C port of the reference's render loop (oracle/, gcc -O3 -ffp-contract=off); the Rust reference cannot be built here"}

    # rays on which the reference's kd walk would have panicked (rank 0's share), with the first location per frame
    panics = [{"frame": j.name, "device_error_bits": st.device_error_bits, "pixel": [st.err_pixel % j.w, st.err_pixel // j.w],
               "sample": st.err_sample, "path": st.err_pathid, "where": st.err_where}
              for j, st in zip(jobs, step_stats[0]) if st.device_error_bits] if b.flags else None
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "ms_per_step_median": statistics.median(step_ms), "ms_per_step_p95": percentile(step_ms, 0.95),
        "ms_per_step_all": [round(x, 3) for x in step_ms],
        "higher_is_better": True, "scaling": wl.get("scaling", "weak"), "vs_baseline": None,
        "dtype": "f64",
        "data": "reference example scenes (procedural geometry, reference OBJ/texture assets" +
                (", stand-ins: " + "; ".join(f"{k}: {v}" for k, v in pt.assets.STAND_INS.items()) if pt.assets.STAND_INS else "") + ")",
        "config": {"workload": wl["label"], "frames_per_step": len(jobs), "samples": samples, "rng": "hash", "seed": SEED,
                   "rays_per_step": rays_per_step, "ms_per_frame": total_ms / args.steps / len(jobs),
                   "l2": "flushed between timed steps (256 MB write)", "tile": "32x32 interleaved over ranks",
                   "streams": len(b.all_streams), "stream_of_frame": list(b.stream_of),
                   "parallelism": f"tiles x{world}", "scene_prepare_ms_host": b.scene_prepare_ms, "scene_broadcast_ms": b.scene_broadcast_ms,
                   "peer_image_setup_ms": b.peer_setup_ms,
                   "exchange": None if world == 1 else ("resolve kernel stores tiles into rank 0's image over peer memory (NVLink) + "
                                                        "1-element all-reduce" if peer is not None else "NCCL gather + un-tiling on rank 0"),
                   "exchange_verified_against_nccl_gather": exchange_verified,
                   "node_pool_retry_rounds_in_timed_steps": int(overflow_retries),
                   # strong scaling is limited by the slowest rank: its own device time per step and its share of the rays
                   "ms_per_step_by_rank": [round(x[0], 3) for x in b.by_rank],
                   "rays_per_step_by_rank": [int(x[1]) for x in b.by_rank],
                   "rank_imbalance_max_over_mean": (max(x[0] for x in b.by_rank) / (sum(x[0] for x in b.by_rank) / len(b.by_rank))) if b.by_rank else None,
                   "reference_panics_tolerated": panics},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_fp64": roofline_fp64,
        "roofline_issue": roofline_issue, "executed_work": executed, "cpu_baseline": cpu, "other_configs": other,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--samples", type=int, default=0, help="samples per pixel at N=1 (default: the workload's)")
    ap.add_argument("--streams", type=int, default=3, help="CUDA streams the frames of a step are spread over (1 = back to back on one)")
    ap.add_argument("--no-balance", dest="balance", action="store_false",
                    help="keep the frames round-robin over the streams instead of balancing the streams by the frames' measured device times")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="N > 1: how the tiles reach rank 0 (peer: stored by the resolve kernel into rank 0's image over NVLink; nccl: gather)")
    ap.add_argument("--device-only", action="store_true", help="skip the e2e, cpu_baseline and other_configs legs (for runs under ncu)")
    ap.add_argument("--no-table", dest="table", action="store_false", help="skip the other_configs table")
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
