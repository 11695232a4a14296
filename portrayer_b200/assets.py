"""Asset lookup for the host mirror.

OBJ meshes are read from ``assets/`` (a git-ignored copy of the reference's
assets made by ``tools/sync_assets.py``) or from their baked form under
``tests/golden/meshes``.  Textures are decoded ONCE here with Pillow and
registered with the host library so that host, oracle and device all read the
same RGB8 bytes (the reference's jpeg-decoder 0.1.15 output cannot be
reproduced offline — SURVEY §2a; decode-once keeps that out of the comparison).

Three images are absent from the reference checkout itself
(``.MISSING_LARGE_BLOBS``): ``earth_cube.png`` is rebuilt as the 4x3 montage of
``earth.jpg`` that ``make-cube-map.sh`` would produce; ``shrub.png`` and
``cpu_cubemap.png`` get a deterministic procedural stand-in.  Any texture that
cannot be found gets the same stand-in, and its name is recorded in
``STAND_INS`` so reports can say so.
"""
from __future__ import annotations

import os

import numpy as np

from . import _ffi

ASSETS_DIR = os.environ.get("PORTRAYER_ASSETS", os.path.join(_ffi.REPO_ROOT, "assets"))
BAKED_MESH_DIR = os.path.join(_ffi.REPO_ROOT, "tests", "golden", "meshes")
STAND_INS: dict[str, str] = {}
_registered: set[str] = set()


def _procedural(name: str, size: int = 1024) -> np.ndarray:
    """Deterministic value-noise stand-in, seeded by the file name."""
    seed = sum(ord(c) * (i + 1) for i, c in enumerate(name)) & 0xFFFFFFFF
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(size // 16, size // 16, 3), dtype=np.uint8)
    img = np.kron(base, np.ones((16, 16, 1), dtype=np.uint8))
    if "shrub" in name.lower():  # hedge-like greens for the castle's maze walls (assets/shrub.png is missing upstream)
        g = base[..., 1].astype(np.int32)
        img = np.kron(np.stack([g // 6 + 10, g // 3 + 70, g // 8 + 8], axis=-1).astype(np.uint8), np.ones((16, 16, 1), dtype=np.uint8))
    if "nor" in name.lower():  # keep stand-in normal maps pointing mostly "out of the surface"
        img = (img.astype(np.int32) // 4 + np.array([96, 96, 191])).clip(0, 255).astype(np.uint8)
    return np.ascontiguousarray(img)


def _decode(path: str) -> np.ndarray | None:
    if not os.path.exists(path):
        return None
    from PIL import Image

    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB"), dtype=np.uint8))  # image::open(..).to_rgb()


def load_texture(path: str) -> np.ndarray:
    """RGB8 array for an example's texture path ("assets/earth.jpg")."""
    name = os.path.basename(path)
    rel = path.split("assets/", 1)[1] if "assets/" in path else name  # "assets/robot-alarm-clock/wallpaper.jpg" keeps its sub-directory
    img = _decode(os.path.join(ASSETS_DIR, rel))
    if img is None:
        img = _decode(os.path.join(ASSETS_DIR, name))
    if img is None and name == "earth_cube.png":
        earth = _decode(os.path.join(ASSETS_DIR, "earth.jpg"))
        if earth is not None:  # montage -mode concatenate -tile 4x3 (make-cube-map.sh:12)
            img = np.ascontiguousarray(np.tile(earth, (3, 4, 1)))
            STAND_INS[name] = "4x3 montage of earth.jpg (make-cube-map.sh recipe); original missing upstream"
    if img is None:
        img = _procedural(name)
        STAND_INS[name] = "procedural stand-in (file not available)"
    return img


@_ffi.TEXTURE_LOADER_FN
def _texture_loader(path_bytes: bytes) -> None:
    path = path_bytes.decode()
    name = os.path.basename(path)
    if name in _registered:
        return
    img = load_texture(path)
    _ffi.host.pth_register_texture(path.encode(), img.shape[1], img.shape[0], img.ctypes.data)
    _registered.add(name)
    if os.environ.get("PORTRAYER_WRITE_DECODED"):
        write_decoded(name, img)


def write_decoded(name: str, img: np.ndarray) -> str:
    """assets/_decoded/<name>.ptex ("PTEX", u32 width, u32 height, RGB8 rows): the decoded form a process WITHOUT this
    Python layer reads (host/scene.cpp RgbImageBuffer::open) — the standalone C++ drivers under tests/cpp and host/tools."""
    out_dir = os.path.join(ASSETS_DIR, "_decoded")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, name + ".ptex")
    with open(out, "wb") as f:
        f.write(b"PTEX" + np.array([img.shape[1], img.shape[0]], dtype="<u4").tobytes())
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())
    return out


def configure() -> None:
    _ffi.host.pth_set_assets_dir(ASSETS_DIR.encode())
    _ffi.host.pth_set_baked_mesh_dir(BAKED_MESH_DIR.encode())
    _ffi.host.pth_set_texture_loader(_texture_loader)


configure()
