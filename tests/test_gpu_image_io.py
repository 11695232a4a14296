"""The byte-moving steps either side of the render loop on the device, -m gpu:
texture ingest (`.to_rgb()` + residency, src/texture.rs:96-102) and PNG encode (Image::save, src/render.rs:200-208)."""
import io
import struct
import zlib

import numpy as np
import pytest

import parity
import portrayer_b200 as pt
from conftest import has_reference_assets
from portrayer_b200 import _ffi

pytestmark = pytest.mark.gpu


def _reference_png(rgb: np.ndarray) -> bytes:
    """the same file written on the CPU: 8-bit RGB, filter 0, stored deflate blocks, one IDAT (host/render.cpp save_as)"""
    h, w = rgb.shape[:2]
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))
    z = bytearray(b"\x78\x01")
    pos = 0
    while True:
        n = min(65535, len(raw) - pos)
        last = pos + n >= len(raw)
        z += struct.pack("<BHH", 1 if last else 0, n, ~n & 0xFFFF) + raw[pos:pos + n]
        pos += n
        if last:
            break
    z += struct.pack(">I", zlib.adler32(raw))

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data))

    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", bytes(z)) + chunk(b"IEND", b"")


# sizes: one pixel; a scanline stream shorter than one deflate block; rows that straddle block boundaries; a stream that
# ends exactly on a block boundary (65535 = 3 * 5 * 17 * 257 -> rows of 1 + 3 * 1456 = 4369 bytes, 15 rows); the
# configs' frame sizes
@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (1456, 15), (1456, 30), (21845, 3), (910, 512), (1920, 1080), (3840, 2160)])
def test_png_encode_on_the_device(gpu_ready, w, h):
    from PIL import Image as PILImage

    rng = np.random.default_rng(w * 7919 + h)
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    png = pt.png_encode(rgb)
    assert len(png) == _ffi.gpu.pt_png_size(w, h)
    assert png == _reference_png(rgb)  # every byte: framing, block headers, Adler-32, CRC-32
    with PILImage.open(io.BytesIO(png)) as im:  # a decoder checks both checksums and returns the pixels
        im.load()
        assert im.mode == "RGB" and im.size == (w, h)
        assert np.array_equal(np.asarray(im), rgb)


def test_frame_png_is_the_rendered_picture(gpu_ready):
    """render -> PNG without the pixels leaving the device in between: the file decodes to what pt_frame_read returns"""
    from PIL import Image as PILImage

    scene = pt.Scene.example("primitives")
    w, h = scene.width // 2, scene.height // 2
    img, _ = parity.render_gpu(scene, samples=1, rng="hash", size=(w, h))
    ds = pt.DeviceScene(scene.blob)
    from portrayer_b200.render import _background_arg, make_params

    bg, bg_mode = _background_arg(scene, w, h)
    fr = pt.Frame(ds, scene.camera(w, h), make_params(w, h, 1, "hash", 1, bg_mode=bg_mode, flags=_ffi.PT_RENDER_ROW_MAJOR))
    fr.set_background(np.ascontiguousarray(bg))
    fr.render()
    png = fr.encode_png()
    with PILImage.open(io.BytesIO(png)) as im:
        assert np.array_equal(np.asarray(im), img.buffer)
    fr.close()
    ds.close()


@pytest.mark.parametrize("channels,bgr", [(1, False), (2, False), (3, False), (4, False), (3, True), (4, True)])
def test_texture_ingest_to_rgb(gpu_ready, channels, bgr):
    """`.to_rgb()` of every DynamicImage layout of image 0.21 (gray replicated, alpha dropped, BGR swapped), ragged sizes"""
    rng = np.random.default_rng(channels * 31 + bgr)
    for k, (w, h) in enumerate([(1, 1), (5, 3), (257, 129), (1024, 768)]):
        px = rng.integers(0, 256, size=(h, w, channels) if channels > 1 else (h, w), dtype=np.uint8)
        key = 0xABCD0000 + channels * 1000 + bgr * 100 + k
        pt.texture_ingest(px, key, bgr=bgr)
        got = pt.texture_read(key)
        p3 = px.reshape(h, w, channels)
        want = np.repeat(p3[..., :1], 3, axis=2) if channels <= 2 else (p3[..., 2::-1] if bgr else p3[..., :3])
        assert np.array_equal(got, want)


@pytest.mark.skipif(not has_reference_assets(), reason="reference textures not synced (tools/sync_assets.py)")
def test_ingested_textures_serve_a_records_only_scene(gpu_ready):
    """decoder output (here with an alpha channel added) -> pt_texture_ingest -> the scene is uploaded WITHOUT its texel
    section and renders exactly like the full blob"""
    import ctypes as C
    import gc

    scene = pt.Scene.example("texture-mapping")
    h = scene.header
    kw = dict(samples=1, rng="fixed", size=(227, 128))
    full = pt.DeviceScene(scene.blob)
    want, _ = parity.render_gpu(scene, dscene=full, **kw)
    full.close()
    gc.collect()
    _ffi.gpu.pt_release_cached_memory()
    assert _ffi.gpu.pt_resident_texture_bytes() == 0
    tex = np.frombuffer(scene.blob, dtype=np.uint8, count=h.n_textures * 32, offset=h.off_textures)
    for i in range(h.n_textures):
        width, height = np.frombuffer(tex[i * 32:i * 32 + 8].tobytes(), dtype="<u4")
        offset, key = np.frombuffer(tex[i * 32 + 8:i * 32 + 24].tobytes(), dtype="<u8")
        assert key != 0
        rgb = np.frombuffer(scene.blob, dtype=np.uint8, count=int(width) * int(height) * 3, offset=h.off_texels + int(offset)).reshape(height, width, 3)
        rgba = np.concatenate([rgb, np.full((height, width, 1), 200, np.uint8)], axis=2)
        pt.texture_ingest(rgba, int(key))
    assert _ffi.gpu.pt_resident_texture_bytes() >= h.n_texel_bytes * 0.9
    records_only = pt.DeviceScene(scene.blob[: h.off_texels])
    assert records_only.uploaded_bytes < h.off_texels + 4096
    got, _ = parity.render_gpu(scene, dscene=records_only, **kw)
    records_only.close()
    assert np.array_equal(got.buffer, want.buffer) and np.array_equal(got.hit_id, want.hit_id)
