"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline /
--impl reference legs).  Nothing under portrayer_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
if not os.path.exists(_LIB):
    raise ImportError(f"{_LIB} is missing: run `make oracle`")
lib = C.CDLL(_LIB)


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "rays_primary", "rays_shadow", "rays_reflect", "rays_refract", "rays_depth_cut", "kd_splits",
        "instance_tests", "triangle_tests", "bbox_gates", "shaded_hits", "texel_lookups",
        "shadow_kd_splits", "shadow_instance_tests", "shadow_triangle_tests", "shadow_bbox_gates")]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}

    @property
    def rays(self) -> int:
        return self.rays_primary + self.rays_shadow + self.rays_reflect + self.rays_refract


lib.oracle_render.restype = C.c_int
lib.oracle_render.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OracleStats)]
lib.oracle_render_rows.restype = C.c_int
lib.oracle_render_rows.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(OracleStats)]
lib.oracle_trace_rays.restype = C.c_int
lib.oracle_trace_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                  C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.POINTER(OracleStats)]
lib.oracle_camera_rays.restype = None
lib.oracle_camera_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
lib.oracle_solve_quadratic.restype = C.c_int
lib.oracle_solve_quadratic.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]


def default_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


class OracleResult:
    def __init__(self, rc, rgb, hit_id, hit_t, color, stats):
        self.rc, self.rgb, self.hit_id, self.hit_t, self.color, self.stats = rc, rgb, hit_id, hit_t, color, stats


def render(blob: np.ndarray, camera, params, background: np.ndarray, rgb_init: np.ndarray | None = None,
           threads: int | None = None, row_stride: int = 1, row_offset: int = 0) -> OracleResult:
    """The pixel loop of src/render.rs:127-150 on the CPU. `camera`/`params` are the ctypes structs of the boundary.
    row_stride > 1: only every row_stride-th row of the slice (a benchmark sample of a big frame)."""
    h, w = params.height, params.width
    rgb = np.zeros((h, w, 3), np.uint8) if rgb_init is None else np.ascontiguousarray(rgb_init).copy()
    hit_id = np.full((h, w, 2), 0xFFFFFFFF, np.uint32)
    hit_t = np.full((h, w), np.inf, np.float64)
    color = np.zeros((h, w, 3), np.float64)
    stats = OracleStats()
    background = np.ascontiguousarray(background, dtype=np.float64)
    blob = np.ascontiguousarray(blob)
    rc = lib.oracle_render_rows(blob.ctypes.data, blob.nbytes, C.addressof(camera), C.addressof(params),
                                background.ctypes.data, rgb.ctypes.data, hit_id.ctypes.data, hit_t.ctypes.data,
                                color.ctypes.data, threads or default_threads(), row_stride, row_offset, C.byref(stats))
    return OracleResult(rc, rgb, hit_id, hit_t, color, stats)


def trace_rays(blob: np.ndarray, origins: np.ndarray, dirs: np.ndarray, background=(0.0, 0.0, 0.0), rng_mode: int = 0,
               seed: int = 1, max_depth: int = 0, threads: int | None = None):
    origins = np.ascontiguousarray(origins, dtype=np.float64)
    dirs = np.ascontiguousarray(dirs, dtype=np.float64)
    n = origins.shape[0]
    bg = np.asarray(background, dtype=np.float64)
    color = np.empty((n, 3), np.float64)
    hit_id = np.empty((n, 2), np.uint32)
    hit_t = np.empty(n, np.float64)
    stats = OracleStats()
    blob = np.ascontiguousarray(blob)
    rc = lib.oracle_trace_rays(blob.ctypes.data, blob.nbytes, n, origins.ctypes.data, dirs.ctypes.data, bg.ctypes.data,
                               rng_mode, seed, max_depth, color.ctypes.data, hit_id.ctypes.data, hit_t.ctypes.data,
                               threads or default_threads(), C.byref(stats))
    return rc, color, hit_id, hit_t, stats


def camera_rays(camera, xy: np.ndarray):
    """Camera::ray_at for [n, 2] pixel-space points -> (origins [n,3], dirs [n,3])."""
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    n = xy.shape[0]
    origins = np.empty((n, 3), np.float64)
    dirs = np.empty((n, 3), np.float64)
    lib.oracle_camera_rays(C.addressof(camera), n, xy.ctypes.data, origins.ctypes.data, dirs.ctypes.data)
    return origins, dirs


def solve_quadratic(a: float, b: float, c: float) -> list[float]:
    out = (C.c_double * 2)()
    n = lib.oracle_solve_quadratic(a, b, c, out)
    return [out[i] for i in range(n)]
