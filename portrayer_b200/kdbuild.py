"""k-d tree build — the step right before the render loop.

``KdTree.build`` runs ``KDLeaf::partitioned`` (src/kdtree/leaf.rs:89-231) on the device through the C ABI
(``pt_kd_build``); ``host_build`` runs the C++ mirror of the same reference code on the CPU (the role the Rust
code plays in the target design) and is what the device trees are compared against.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import check, gpu, host

KD_NODE_DTYPE = np.dtype([("split", "<f8"), ("a", "<u4"), ("b", "<u4")])  # PtKdNode


class PtKdBuildConfig(C.Structure):
    _fields_ = [("max_depth", C.c_uint32), ("target_max_nodes", C.c_uint32), ("target_max_merit", C.c_int32),
                ("max_tries", C.c_uint32)]


def config(max_depth: int = 10, target_max_nodes: int = 3, target_max_merit: int = 3, max_tries: int = 10) -> PtKdBuildConfig:
    """The reference's PartitionConfig (kdscene.rs:30-34, kdmesh.rs:45-49) and MAX_TREE_DEPTH (kdscene.rs:13)."""
    return PtKdBuildConfig(max_depth, target_max_nodes, target_max_merit, max_tries)


class KdTree:
    """A tree built on, and resident in, the device (``PtKdTree``)."""

    def __init__(self, handle: C.c_void_p):
        self._h = handle

    @classmethod
    def build(cls, bounds: np.ndarray, cfg: PtKdBuildConfig | None = None) -> "KdTree":
        """bounds: [n, 6] float64 host array = {min x, y, z, max x, y, z} of every item."""
        bounds = np.ascontiguousarray(bounds, dtype=np.float64).reshape(-1, 6)
        cfg = cfg or config()
        h = C.c_void_p()
        check(gpu.pt_kd_build(bounds.ctypes.data, bounds.shape[0], C.byref(cfg), C.byref(h)))
        return cls(h)

    @classmethod
    def build_device(cls, d_bounds: int, n: int, cfg: PtKdBuildConfig | None = None, stream: int | None = None) -> "KdTree":
        cfg = cfg or config()
        h = C.c_void_p()
        check(gpu.pt_kd_build_device(C.c_void_p(d_bounds), n, C.byref(cfg), C.c_void_p(stream) if stream else None, C.byref(h)))
        return cls(h)

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def node_count(self) -> int:
        return gpu.pt_kd_tree_node_count(self._h)

    @property
    def item_count(self) -> int:
        return gpu.pt_kd_tree_item_count(self._h)

    @property
    def depth(self) -> int:
        return gpu.pt_kd_tree_depth(self._h)

    def root_bounds(self) -> tuple[np.ndarray, float]:
        b = np.empty(6, np.float64)
        ext = C.c_double()
        check(gpu.pt_kd_tree_root_bounds(self._h, b.ctypes.data, C.byref(ext)))
        return b, ext.value

    def build_stats(self) -> tuple[float, int]:
        """(device milliseconds, kernel launches) of the build"""
        ms, n = C.c_double(), C.c_uint32()
        check(gpu.pt_kd_tree_build_stats(self._h, C.byref(ms), C.byref(n), None))
        return ms.value, n.value

    def algorithmic_bytes(self) -> int:
        """bytes the build has to move at the least (members x passes): the numerator of its HBM roofline"""
        b = C.c_uint64()
        check(gpu.pt_kd_tree_build_stats(self._h, None, None, C.byref(b)))
        return b.value

    def download(self) -> tuple[np.ndarray, np.ndarray]:
        nodes = np.empty(self.node_count, KD_NODE_DTYPE)
        items = np.empty(self.item_count, np.uint32)
        check(gpu.pt_kd_tree_download(self._h, nodes.ctypes.data, items.ctypes.data))
        return nodes, items

    def close(self) -> None:
        if self._h:
            gpu.pt_kd_tree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostTree:
    def __init__(self, nodes, items, depth, extent, seconds):
        self.nodes, self.items, self.depth, self.extent, self.seconds = nodes, items, depth, extent, seconds


def host_build(bounds: np.ndarray, cfg: PtKdBuildConfig | None = None) -> HostTree:
    """The same build by the C++ mirror of the reference (CPU, single thread like the reference's)."""
    bounds = np.ascontiguousarray(bounds, dtype=np.float64).reshape(-1, 6)
    cfg = cfg or config()
    h = host.pth_kd_build(bounds.ctypes.data, bounds.shape[0], cfg.max_depth, cfg.target_max_nodes, cfg.target_max_merit, cfg.max_tries)
    if not h:
        raise RuntimeError((host.pth_last_error() or b"").decode())
    try:
        nn, ni = host.pth_kd_tree_node_count(h), host.pth_kd_tree_item_count(h)
        nodes = np.frombuffer((C.c_uint8 * (nn * 16)).from_address(host.pth_kd_tree_nodes(h)), dtype=KD_NODE_DTYPE).copy() if nn else np.empty(0, KD_NODE_DTYPE)
        items = np.frombuffer((C.c_uint8 * (ni * 4)).from_address(host.pth_kd_tree_items(h)), dtype=np.uint32).copy() if ni else np.empty(0, np.uint32)
        return HostTree(nodes, items, host.pth_kd_tree_depth(h), host.pth_kd_tree_extent(h), host.pth_kd_tree_build_seconds(h))
    finally:
        host.pth_kd_tree_free(h)


def triangle_bounds(tri_pos: np.ndarray) -> np.ndarray:
    """Triangle::bounds (src/primitive/triangle.rs:30-35) of [n, 9] vertex records -> [n, 6]"""
    v = np.asarray(tri_pos, np.float64).reshape(-1, 3, 3)
    return np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1)
