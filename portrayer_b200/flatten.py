"""Scene flattening — ``FlatScene::from`` (src/flat_scene.rs:18-46) on the device through ``pt_flatten``
(SURVEY §8f rank 2).  ``hierarchy_of(scene)`` exports an example scene's graph in the boundary's records."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import check, gpu, host

HIER_NODE_DTYPE = np.dtype([("trans", "<f8", 16), ("geometry", "<u4"), ("first_child", "<u4"), ("child_count", "<u4"),
                            ("reserved", "<u4")])  # PtHierNode, 144 B
GEOMETRY_DTYPE = np.dtype([("bounds", "<f8", 6), ("prim", "<u4"), ("mesh", "<u4"), ("material", "<u4"), ("reserved", "<u4")])  # 64 B
INSTANCE_DTYPE = np.dtype([("invtrans", "<f8", 12), ("prim", "<u4"), ("mesh", "<u4"), ("material", "<u4"), ("reserved", "<u4"),
                           ("pad", "<f8", 2)])  # PtInstance, 128 B
assert HIER_NODE_DTYPE.itemsize == 144 and GEOMETRY_DTYPE.itemsize == 64 and INSTANCE_DTYPE.itemsize == 128


class Hierarchy:
    def __init__(self, nodes, children, geometries, root):
        self.nodes, self.children, self.geometries, self.root = nodes, children, geometries, root


def hierarchy_of(scene) -> Hierarchy:
    """the scene graph of a host-built scene (host mirror: pack.cpp export_hierarchy)"""
    nn, nc, ng, root = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    if host.pth_scene_hierarchy_sizes(scene._h, C.byref(nn), C.byref(nc), C.byref(ng), C.byref(root)) != 0:
        raise ValueError(f"scene {scene.name!r} has no scene graph (hand-built known-answer scene)")
    nodes = np.zeros(nn.value, HIER_NODE_DTYPE)
    children = np.zeros(max(nc.value, 1), np.uint32)
    geoms = np.zeros(max(ng.value, 1), GEOMETRY_DTYPE)
    host.pth_scene_hierarchy(scene._h, nodes.ctypes.data, children.ctypes.data, geoms.ctypes.data)
    return Hierarchy(nodes, children[: nc.value], geoms[: ng.value], root.value)


class FlatScene:
    """Flat instances built on, and resident in, the device (``PtFlatScene``)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def build(cls, hier: Hierarchy) -> "FlatScene":
        nodes = np.ascontiguousarray(hier.nodes)
        children = np.ascontiguousarray(hier.children, dtype=np.uint32)
        geoms = np.ascontiguousarray(hier.geometries)
        h = C.c_void_p()
        check(gpu.pt_flatten(nodes.ctypes.data, len(nodes), children.ctypes.data if len(children) else None, len(children),
                             hier.root, geoms.ctypes.data if len(geoms) else None, len(geoms), C.byref(h)))
        return cls(h)

    @property
    def handle(self):
        return self._h

    @property
    def instance_count(self) -> int:
        return gpu.pt_flat_instance_count(self._h)

    @property
    def bounds_device_ptr(self) -> int:
        return gpu.pt_flat_bounds_device(self._h) or 0

    def build_stats(self) -> tuple[float, int]:
        ms, n = C.c_double(), C.c_uint32()
        check(gpu.pt_flat_build_stats(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def download(self):
        """(instances [n] INSTANCE_DTYPE, trans [n, 12], bounds [n, 6])"""
        n = self.instance_count
        inst = np.zeros(n, INSTANCE_DTYPE)
        trans = np.zeros((n, 12), np.float64)
        bounds = np.zeros((n, 6), np.float64)
        if n:
            check(gpu.pt_flat_download(self._h, inst.ctypes.data, trans.ctypes.data, bounds.ctypes.data))
        return inst, trans, bounds

    def close(self) -> None:
        if self._h:
            gpu.pt_flat_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
