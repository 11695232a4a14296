"""world_size-2 test of the multi-GPU plumbing on CPU (gloo): scene broadcast, interleaved tile ownership,
gather to rank 0.  The per-rank renderer here is the ORACLE (tests may use it); what is under test is
portrayer_b200.distributed and the tile partition of csrc/tiles.c."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch

    import portrayer_b200 as pt
    from oracle import binding as oracle
    from portrayer_b200 import distributed as ptd
    from portrayer_b200.render import _background_arg, make_params

    ptd.init_process_group("gloo")
    scene = pt.Scene.example("nonhier") if rank == 0 else None
    blob = ptd.broadcast_blob(scene.blob if rank == 0 else None, torch.device("cpu")).numpy()
    ref_scene = pt.Scene.example("nonhier")  # camera / background are tiny host-side inputs every rank rebuilds
    assert rank != 0 or np.array_equal(blob, scene.blob)
    w, h = 96, 80
    bg, bg_mode = _background_arg(ref_scene, w, h)
    params = make_params(w, h, 2, "hash", 7, bg_mode=bg_mode, rank=rank, world=world, tile=16)
    res = oracle.render(blob, ref_scene.camera(w, h), params, bg, threads=2)
    assert res.rc == 0
    index = ptd.owned_pixel_index(params)
    local = torch.from_numpy(res.rgb.reshape(-1, 3)[index.astype(np.int64)].copy())
    image = ptd.gather_image(local, params, dst=0)
    # several frames of one geometry in one exchange (what a benchmark step does): frame 1 = frame 0 inverted
    multi = ptd.gather_images_device([local, 255 - local], params, dst=0)
    if rank == 0:
        full = oracle.render(blob, ref_scene.camera(w, h), make_params(w, h, 2, "hash", 7, bg_mode=bg_mode), bg, threads=2)
        multi = multi.numpy().reshape(2, h, w, 3)
        np.save(out_path, np.stack([image, full.rgb, multi[0], 255 - multi[1]]))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_tiles_gather_to_rank0(tmp_path):
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    gathered, full, multi0, multi1 = np.load(out)
    assert np.array_equal(gathered, full), "the gathered 2-rank image must be bit-identical to the 1-rank image"
    assert np.array_equal(multi0, full) and np.array_equal(multi1, full), "the batched multi-frame exchange must agree"


def _handle_worker(rank, world, port, out_path):
    sys.path.insert(0, REPO)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch

    from portrayer_b200 import distributed as ptd

    ptd.init_process_group("gloo")
    handle = bytes(range(100, 164)) if rank == 1 else None  # the collecting rank need not be rank 0
    got = ptd.exchange_handle(handle, 1, torch.device("cpu"))
    assert len(got) == 64
    if rank == 0:
        np.save(out_path, np.frombuffer(got, np.uint8))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_peer_handle_reaches_every_rank(tmp_path):
    """the 64-byte peer-memory handle of distributed.PeerImage (pt_peer_alloc -> pt_peer_open) crosses ranks intact"""
    out = str(tmp_path / "handle.npy")
    mp.spawn(_handle_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert np.array_equal(np.load(out), np.arange(100, 164, dtype=np.uint8))


def test_tile_partition_is_a_partition():
    sys.path.insert(0, REPO)
    from portrayer_b200 import distributed as ptd
    from portrayer_b200.render import make_params

    w, h = 100, 70
    for world in (1, 2, 3, 8):
        seen = np.zeros(w * h, np.int32)
        for rank in range(world):
            idx = ptd.owned_pixel_index(make_params(w, h, 1, rank=rank, world=world, tile=32))
            seen[idx] += 1
        assert np.all(seen == 1), f"world={world}: every pixel must be owned exactly once"
    # slices restrict ownership to the inclusive rectangle (render.rs:116-119)
    idx = ptd.owned_pixel_index(make_params(w, h, 1, slice_=(10, 5, 19, 9)))
    assert len(idx) == 10 * 5 and set(idx // w) == set(range(5, 10)) and set(idx % w) == set(range(10, 20))
    # a warp's 32 consecutive pixels form an 8x4 block
    idx = ptd.owned_pixel_index(make_params(64, 64, 1))
    first = idx[:32]
    assert set(first % 64) == set(range(8)) and set(first // 64) == set(range(4))
