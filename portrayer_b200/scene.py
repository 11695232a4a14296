"""Scene programs: the Python handle on what an ``examples/*.rs`` main() builds.

A :class:`Scene` is the host-side result of "scene building + FlatScene::from +
KDTreeScene::from + flatten to SoA" (src/render.rs:124-126 and the glue of
SURVEY §8b), i.e. everything the reference does BEFORE its pixel loop: the
packed blob, the camera settings, the image size and the background closure.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi, assets  # noqa: F401  (assets configures the host library's loaders)
from ._ffi import PtBlobHeader, PtCamera, host


def example_names() -> list[str]:
    return [host.pth_example_name(i).decode() for i in range(host.pth_example_count())]


class Scene:
    def __init__(self, handle: int, name: str):
        if not handle:
            raise RuntimeError(f"building scene {name!r} failed: {(host.pth_last_error() or b'').decode()}")
        self._h = C.c_void_p(handle)
        self.name = name
        w, h = C.c_uint32(), C.c_uint32()
        host.pth_image_size(self._h, C.byref(w), C.byref(h))
        self.width, self.height = w.value, h.value
        n = host.pth_blob_size(self._h)
        buf = (C.c_uint8 * n).from_address(host.pth_blob_data(self._h))
        self.blob = np.frombuffer(buf, dtype=np.uint8)  # view; the handle owns the memory
        self.header = PtBlobHeader.from_buffer_copy(self.blob[: C.sizeof(PtBlobHeader)].tobytes())
        self.prepare_seconds = host.pth_prepare_seconds(self._h)

    # ---- constructors
    @classmethod
    def example(cls, name: str, kd_depth: int = -1, linear_tlas: bool = False, kd_mesh_depth: int | None = None) -> "Scene":
        """kd_depth < 0: KD_DEPTH env or 10.  kd_mesh_depth: sets KD_MESH_DEPTH for KDMesh::new (kdmesh.rs:51-53)."""
        old = os.environ.get("KD_MESH_DEPTH")
        if kd_mesh_depth is not None:
            os.environ["KD_MESH_DEPTH"] = str(kd_mesh_depth)
        try:
            return cls(host.pth_example_build(name.encode(), kd_depth, int(linear_tlas)), name)
        finally:
            if kd_mesh_depth is not None:
                if old is None:
                    os.environ.pop("KD_MESH_DEPTH", None)
                else:
                    os.environ["KD_MESH_DEPTH"] = old

    @classmethod
    def big_scene(cls, n: int = 10, kd_depth: int = -1, linear_tlas: bool = False) -> "Scene":
        return cls(host.pth_big_scene_build(n, kd_depth, int(linear_tlas)), f"big-scene-n{n}")

    @classmethod
    def synthetic_instances(cls, n_instances: int, seed: int = 1234939301, kd_depth: int = -1) -> "Scene":
        return cls(host.pth_synthetic_instances_build(n_instances, seed, kd_depth), f"synthetic-instances-{n_instances}")

    @classmethod
    def synthetic_triangles(cls, n_triangles: int, seed: int = 1234939301, kd_mesh_depth: int = 10) -> "Scene":
        return cls(host.pth_synthetic_triangles_build(n_triangles, seed, kd_mesh_depth), f"synthetic-triangles-{n_triangles}")

    # ---- what Image::render computes before the loop
    def camera(self, width: int | None = None, height: int | None = None) -> PtCamera:
        cam = PtCamera()
        host.pth_camera(self._h, float(width or self.width), float(height or self.height), C.byref(cam))
        return cam

    def background(self, width: int | None = None, height: int | None = None) -> np.ndarray:
        """background.at(x/w, y/h) for integer pixels (render.rs:31-34): float64 [H, W, 3]."""
        w, h = width or self.width, height or self.height
        out = np.empty((h, w, 3), dtype=np.float64)
        host.pth_background(self._h, w, h, out.ctypes.data)
        return out

    def background_rows(self, width: int | None = None, height: int | None = None) -> np.ndarray | None:
        """[H, 3] when the closure depends on v only (every shipped example), else None."""
        bg = self.background(width, height)
        if np.array_equal(bg, np.broadcast_to(bg[:, :1, :], bg.shape)):
            return np.ascontiguousarray(bg[:, 0, :])
        return None

    def item_bounds(self) -> np.ndarray:
        """FlatSceneNode::bounds of every flat instance (flat_scene.rs:63-69): [n, 6] = min xyz, max xyz — the input of
        the scene-tree build (kdscene.rs:24-28)."""
        n = host.pth_scene_item_count(self._h)
        out = np.empty((n, 6), dtype=np.float64)
        if n:
            host.pth_scene_item_bounds(self._h, out.ctypes.data)
        return out

    def section(self, name: str, dtype, width: int) -> np.ndarray:
        """a record section of the blob as an [n, width] array view, e.g. section('tri_pos', np.float64, 9)"""
        h = self.header
        n = getattr(h, {"tlas_nodes": "n_tlas_nodes", "tlas_items": "n_tlas_items", "blas_nodes": "n_blas_nodes",
                        "blas_items": "n_blas_items", "tri_pos": "n_triangles"}[name])
        off = getattr(h, "off_" + name)
        dt = np.dtype(dtype)
        return np.frombuffer(self.blob, dtype=dt, count=n * width, offset=off).reshape(n, width) if width > 1 else \
            np.frombuffer(self.blob, dtype=dt, count=n, offset=off)

    def close(self) -> None:
        if self._h:
            self.blob = None
            host.pth_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
