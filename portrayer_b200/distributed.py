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


def gather_image(local_rgb: torch.Tensor, params: PtRenderParams, dst: int = 0, out: np.ndarray | None = None):
    """Gather every rank's compact [owned, 3] uint8 pixels to ``dst`` and place them into a [H, W, 3] image.
    Returns the image on ``dst`` (None elsewhere). ``out`` lets the caller keep pixels outside the slice."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    h, w = params.height, params.width
    counts = [int(_ffi.gpu.pt_owned_pixels(C.byref(_with_rank(params, r, world)), None, 0)) for r in range(world)]
    assert local_rgb.shape[0] == counts[rank], (local_rgb.shape, counts[rank])
    if world == 1:
        pieces = [local_rgb]
    else:
        # ranks own slightly different pixel counts: pad to the maximum so one gather moves everything
        pad = max(counts)
        send = torch.zeros((pad, 3), dtype=torch.uint8, device=local_rgb.device)
        send[: counts[rank]] = local_rgb
        recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send, recv, dst=dst)
        if rank != dst:
            return None
        pieces = [recv[r][: counts[r]] for r in range(world)]
    image = out if out is not None else np.zeros((h, w, 3), np.uint8)
    flat = image.reshape(-1, 3)
    for r, piece in enumerate(pieces):
        flat[owned_pixel_index(params, r, world).astype(np.int64)] = piece.cpu().numpy()
    return image


def _with_rank(params: PtRenderParams, rank: int, world: int) -> PtRenderParams:
    p = PtRenderParams.from_buffer_copy(bytes(params))
    p.rank, p.world = rank, world
    return p
