"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch).

The image shards naturally (src/render.rs:127-150 keeps no cross-pixel state), so the data path
needs no collective: the scene blob is broadcast once, every rank renders its interleaved tiles
(all samples of a pixel stay on the owner), and the compact RGB8 results are gathered to rank 0.
Nothing is reduced, so the gathered image is bit-identical for any world size.

The same functions run on CPU tensors over gloo — that is how the plumbing is tested without GPUs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi
from ._ffi import PtRenderParams


def init_process_group(backend: str | None = None) -> tuple[int, int, int]:
    """Rendezvous from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def owned_pixel_index(params: PtRenderParams, rank: int | None = None, world: int | None = None) -> np.ndarray:
    """Global pixel indices rendered by (rank, world), in output order (pure host code, csrc/tiles.c)."""
    p = PtRenderParams.from_buffer_copy(bytes(params))
    if rank is not None:
        p.rank, p.world = rank, world
    n = _ffi.gpu.pt_owned_pixels(C.byref(p), None, 0)
    out = np.empty(n, np.uint32)
    _ffi.gpu.pt_owned_pixels(C.byref(p), out.ctypes.data, n)
    return out


def broadcast_blob(blob: np.ndarray | None, device: torch.device, src: int = 0) -> torch.Tensor:
    """Broadcast the pointer-free scene blob from ``src`` to every rank; returns a uint8 tensor on ``device``."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return torch.from_numpy(np.ascontiguousarray(blob)).to(device)
    size = torch.tensor([blob.nbytes if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(size, src)
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(blob)).to(device)
    else:
        t = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return t


class _CudaView:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, shape: tuple, typestr: str = "|u1", device: int = 0) -> torch.Tensor:
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device=torch.device("cuda", device))


class GatherPlan:
    """Everything about the exchange step that depends only on the image geometry, computed once: how many pixels
    every rank owns, the padded message size, and (on ``dst``) the scatter indices — on the device for CUDA tensors,
    so a step's gather is one NCCL call plus one device-side scatter per rank, with no host work and no D2H."""

    _cache: dict = {}

    def __init__(self, params: PtRenderParams, device: torch.device, dst: int):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst = dst
        self.h, self.w = params.height, params.width
        self.counts = [int(_ffi.gpu.pt_owned_pixels(C.byref(_with_rank(params, r, self.world)), None, 0)) for r in range(self.world)]
        self.pad = max(self.counts) if self.counts else 0
        self.device = device
        self.send = torch.zeros((self.pad, 3), dtype=torch.uint8, device=device) if self.world > 1 else None
        self.recv = None
        self.index = None
        self.multi: dict = {}  # n_frames -> (send, recv, image) of gather_images_device
        if self.rank == dst:
            self.index = [torch.from_numpy(owned_pixel_index(params, r, self.world).astype(np.int64)).to(device) for r in range(self.world)]
            if self.world > 1:
                self.recv = [torch.empty((self.pad, 3), dtype=torch.uint8, device=device) for _ in range(self.world)]
            self.image = torch.zeros((self.h * self.w, 3), dtype=torch.uint8, device=device)

    @classmethod
    def of(cls, params: PtRenderParams, device: torch.device, dst: int) -> "GatherPlan":
        key = (params.width, params.height, params.x1, params.y1, params.x2, params.y2, params.tile_w, params.tile_h,
               dist.get_world_size() if dist.is_initialized() else 1, str(device), dst)
        plan = cls._cache.get(key)
        if plan is None:
            plan = cls._cache[key] = cls(params, device, dst)
        return plan


def gather_image_device(local_rgb: torch.Tensor, params: PtRenderParams, dst: int = 0) -> torch.Tensor | None:
    """The exchange step, device-resident: every rank's compact [owned, 3] uint8 pixels -> a [H*W, 3] image tensor on
    ``dst`` (returned there, None elsewhere).  Stream-ordered, no host synchronisation."""
    plan = GatherPlan.of(params, local_rgb.device, dst)
    assert local_rgb.shape[0] == plan.counts[plan.rank], (local_rgb.shape, plan.counts[plan.rank])
    if plan.world == 1:
        plan.image.index_copy_(0, plan.index[0], local_rgb)
        return plan.image
    plan.send[: plan.counts[plan.rank]].copy_(local_rgb)
    dist.gather(plan.send, plan.recv, dst=dst)
    if plan.rank != dst:
        return None
    for r in range(plan.world):
        plan.image.index_copy_(0, plan.index[r], plan.recv[r][: plan.counts[r]])
    return plan.image


def gather_images_device(local_rgbs: list[torch.Tensor], params: PtRenderParams, dst: int = 0) -> torch.Tensor | None:
    """The exchange step for SEVERAL frames of the same geometry at once (a step of the benchmark renders five):
    one NCCL gather of an [F, pad, 3] message per rank and one device-side scatter per rank, instead of F of each —
    at 1.4 MB per frame the exchange is latency-bound, so the call count is what matters.  Returns [F, H*W, 3] on
    ``dst`` (None elsewhere).  Stream-ordered, no host synchronisation."""
    plan = GatherPlan.of(params, local_rgbs[0].device, dst)
    n_frames = len(local_rgbs)
    own = plan.counts[plan.rank]
    bufs = plan.multi.get(n_frames)
    if bufs is None:
        send = torch.zeros((n_frames, plan.pad, 3), dtype=torch.uint8, device=plan.device)
        recv = [torch.empty((n_frames, plan.pad, 3), dtype=torch.uint8, device=plan.device) for _ in range(plan.world)] \
            if plan.rank == dst and plan.world > 1 else None
        image = torch.zeros((n_frames, plan.h * plan.w, 3), dtype=torch.uint8, device=plan.device) if plan.rank == dst else None
        bufs = plan.multi[n_frames] = (send, recv, image)
    send, recv, image = bufs
    for f, rgb in enumerate(local_rgbs):
        assert rgb.shape[0] == own, (rgb.shape, own)
        send[f, :own].copy_(rgb)
    if plan.world == 1:
        image.index_copy_(1, plan.index[0], send[:, :own])
        return image
    dist.gather(send, recv, dst=dst)
    if plan.rank != dst:
        return None
    for r in range(plan.world):
        image.index_copy_(1, plan.index[r], recv[r][:, : plan.counts[r]])
    return image


def gather_image(local_rgb: torch.Tensor, params: PtRenderParams, dst: int = 0, out: np.ndarray | None = None):
    """Gather every rank's compact [owned, 3] uint8 pixels to ``dst`` and place them into a [H, W, 3] host image.
    Returns the image on ``dst`` (None elsewhere). ``out`` lets the caller keep pixels outside the slice."""
    image_dev = gather_image_device(local_rgb, params, dst)
    if image_dev is None:
        return None
    plan = GatherPlan.of(params, local_rgb.device, dst)
    host = image_dev.cpu().numpy().reshape(plan.h, plan.w, 3)
    if out is None:
        if (params.x1, params.y1, params.x2, params.y2) == (0, 0, plan.w - 1, plan.h - 1):
            return host
        out = np.zeros((plan.h, plan.w, 3), np.uint8)
    out[params.y1:params.y2 + 1, params.x1:params.x2 + 1] = host[params.y1:params.y2 + 1, params.x1:params.x2 + 1]
    return out


class PeerImage:
    """``n_images`` full RGB8 images ([n, H*W, 3] bytes) that live on ``dst``'s GPU and are mapped into every other
    rank's address space (CUDA IPC -> peer access over NVLink / NVSwitch): frames pointed at ``ptr(i)`` with
    ``Frame.set_image_target`` have their resolve kernel store the rank's tiles straight into the collecting rank's
    picture, so the exchange step needs no gather, no padding copy and no un-tiling — only ``signal_done()``, a
    one-element all-reduce that orders every rank's stores before ``dst`` reads the images.

    ``exchange_handle`` is the only part that touches ``torch.distributed`` (a 64-byte broadcast); it is separate so
    that the handle plumbing is testable over gloo on CPU with a stand-in allocator."""

    def __init__(self, n_images: int, height: int, width: int, device: int, dst: int = 0):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst, self.n, self.h, self.w, self.device = dst, n_images, height, width, device
        self.image_bytes = height * width * 3
        self._owned = self.rank == dst
        self._ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        alloc_error = None
        if self._owned:
            try:
                _ffi.check(_ffi.gpu.pt_peer_alloc(self.image_bytes * n_images, C.byref(self._ptr), handle))
            except _ffi.PortrayerError as exc:  # still take part in the broadcast: an all-zero handle tells every rank
                alloc_error = exc
                handle = (C.c_ubyte * 64)()
        payload = exchange_handle(bytes(handle) if self._owned else None, dst,
                                  torch.device("cuda", device) if dist.is_initialized() and dist.get_backend() == "nccl" else torch.device("cpu"))
        if not any(payload):
            raise RuntimeError(f"the collecting rank could not allocate the shared image: {alloc_error or 'see rank ' + str(dst)}")
        if not self._owned:
            buf = (C.c_ubyte * 64).from_buffer_copy(payload)
            _ffi.check(_ffi.gpu.pt_peer_open(buf, C.byref(self._ptr)))
        self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", device))

    def ptr(self, image: int = 0) -> int:
        return self._ptr.value + image * self.image_bytes

    def signal_done(self) -> None:
        """stream-ordered completion signal: when it has passed on ``dst``, every rank's resolve kernels that were
        enqueued before it have finished and their peer stores are visible"""
        if self.world > 1:
            dist.all_reduce(self._flag)

    def images(self) -> torch.Tensor | None:
        """[n, H, W, 3] uint8 view of the images on ``dst`` (None elsewhere)"""
        if not self._owned:
            return None
        return device_tensor(self._ptr.value, (self.n, self.h, self.w, 3), "|u1", self.device)

    def close_local(self) -> None:
        """release this rank's mapping / allocation without any collective (error paths)"""
        if self._ptr.value:
            (_ffi.gpu.pt_peer_free if self._owned else _ffi.gpu.pt_peer_close)(self._ptr)
            self._ptr = C.c_void_p()

    def close(self) -> None:
        if self._ptr.value:
            if self._owned:
                if self.world > 1:
                    dist.barrier()  # nobody may still be storing into memory that is about to be freed
                _ffi.gpu.pt_peer_free(self._ptr)
            else:
                _ffi.gpu.pt_peer_close(self._ptr)
                if self.world > 1:
                    dist.barrier()
            self._ptr = C.c_void_p()


def exchange_handle(handle: bytes | None, src: int, device: torch.device) -> bytes:
    """broadcast a 64-byte peer-memory handle from ``src`` to every rank"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return handle
    t = torch.zeros(64, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(handle), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def _with_rank(params: PtRenderParams, rank: int, world: int) -> PtRenderParams:
    p = PtRenderParams.from_buffer_copy(bytes(params))
    p.rank, p.world = rank, world
    return p
