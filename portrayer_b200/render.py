"""Python surface of the render entry point — ``Image::render`` of src/render.rs.

``Image.render(scene, ...)`` is the call a user of the reference makes; here it
goes through the C ABI (``pt_scene_upload`` + ``pt_render``) with host buffers.
``DeviceScene`` / ``Frame`` expose the device-resident form used by the
multi-GPU path and the benchmark.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi
from ._ffi import (PT_BG_CONSTANT, PT_BG_PER_PIXEL, PT_BG_PER_ROW, PT_DEFAULT_SAMPLES, PT_RNG_FIXED, PT_RNG_HASH,
                   PtCamera, PtRenderParams, PtStats, check, gpu)
from .scene import Scene


def samples_from_env(default: int = PT_DEFAULT_SAMPLES) -> int:
    """SAMPLES: must parse and be > 0, else the default of 100 (render.rs:107-113)."""
    try:
        v = int(os.environ.get("SAMPLES", ""))
        return v if v > 0 else default
    except ValueError:
        return default


def rng_mode_of(name: str | int) -> int:
    if isinstance(name, int):
        return name
    return {"fixed": PT_RNG_FIXED, "hash": PT_RNG_HASH}[name]


def make_params(width: int, height: int, samples: int, rng: str | int = "hash", seed: int = 1, slice_=None,
                bg_mode: int = PT_BG_PER_PIXEL, rank: int = 0, world: int = 1, tile: int = 32, flags: int = 0,
                max_depth: int = 0, max_batch_paths: int = 0, node_pool_capacity: int = 0) -> PtRenderParams:
    x1, y1, x2, y2 = slice_ if slice_ is not None else (0, 0, width - 1, height - 1)
    return PtRenderParams(width=width, height=height, x1=x1, y1=y1, x2=x2, y2=y2, samples=samples,
                          rng_mode=rng_mode_of(rng), seed=seed, bg_mode=bg_mode, max_depth=max_depth, tile_w=tile,
                          tile_h=tile, rank=rank, world=world, max_batch_paths=max_batch_paths,
                          node_pool_capacity=node_pool_capacity, flags=flags)


def _background_arg(scene: Scene, width: int, height: int) -> tuple[np.ndarray, int]:
    rows = scene.background_rows(width, height)
    if rows is not None:
        return rows, PT_BG_PER_ROW
    return np.ascontiguousarray(scene.background(width, height)), PT_BG_PER_PIXEL


def png_encode(rgb: np.ndarray) -> bytes:
    """Image::save's encode step (src/render.rs:200-208) on the device: [H, W, 3] uint8 -> the bytes of a PNG file"""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w = rgb.shape[:2]
    buf = np.empty(gpu.pt_png_size(w, h), np.uint8)
    n = C.c_uint64(0)
    check(gpu.pt_png_encode(rgb.ctypes.data, w, h, buf.ctypes.data, buf.nbytes, C.byref(n)))
    return buf[: n.value].tobytes()


_LAYOUTS = {1: _ffi.PT_PIXELS_LUMA8, 2: _ffi.PT_PIXELS_LUMAA8, 3: _ffi.PT_PIXELS_RGB8, 4: _ffi.PT_PIXELS_RGBA8}


def texture_ingest(pixels: np.ndarray, key: int, bgr: bool = False) -> None:
    """RgbImageBuffer::open's `.to_rgb()` + upload (src/texture.rs:96-102) on the device: the decoder's output
    ([H, W] or [H, W, C] uint8, C = 1..4) becomes resident RGB8 texels under ``key``."""
    pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
    h, w = pixels.shape[:2]
    c = 1 if pixels.ndim == 2 else pixels.shape[2]
    layout = _LAYOUTS[c]
    if bgr:
        layout = {3: _ffi.PT_PIXELS_BGR8, 4: _ffi.PT_PIXELS_BGRA8}[c]
    check(gpu.pt_texture_ingest(pixels.ctypes.data, w, h, layout, key))


def texture_read(key: int) -> np.ndarray:
    w, h = C.c_uint32(0), C.c_uint32(0)
    check(gpu.pt_texture_read(key, None, 0, C.byref(w), C.byref(h)))
    out = np.empty((h.value, w.value, 3), np.uint8)
    check(gpu.pt_texture_read(key, out.ctypes.data, out.nbytes, C.byref(w), C.byref(h)))
    return out


def init_devices(ids) -> int:
    """pt_init_devices: make ``ids`` one device group in this process (ids[0] = primary); from then on ``Image.render`` /
    ``DeviceScene.render`` fan their tiles over all members.  Handles created before the call are invalid."""
    ids = [int(i) for i in ids]
    arr = (C.c_int * len(ids))(*ids)
    check(gpu.pt_init_devices(arr, len(ids)))
    return gpu.pt_device_group_size()


class DeviceScene:
    """A scene blob resident in HBM (``PtScene``)."""

    def __init__(self, blob: np.ndarray | None = None, device_ptr: int | None = None, nbytes: int | None = None):
        self._h = C.c_void_p()
        if device_ptr is not None:
            check(gpu.pt_scene_upload_device(C.c_void_p(device_ptr), nbytes, C.byref(self._h)))
        else:
            blob = np.ascontiguousarray(blob)
            check(gpu.pt_scene_upload(blob.ctypes.data, blob.nbytes, C.byref(self._h)))

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def set_tlas(self, tree) -> None:
        """make a device-built tree (kdbuild.KdTree) the scene tree of this uploaded scene"""
        check(gpu.pt_scene_set_tlas(self._h, tree.handle))

    def set_instances(self, flat, tree) -> None:
        """make device-flattened instances (flatten.FlatScene) and a tree built over their bounds the scene's"""
        check(gpu.pt_scene_set_instances(self._h, flat.handle, tree.handle))

    @property
    def uploaded_bytes(self) -> int:
        """host -> device bytes of the upload (records + textures that were not already resident)"""
        return gpu.pt_scene_uploaded_bytes(self._h)

    def render(self, camera: PtCamera, params: PtRenderParams, background: np.ndarray, rgb: np.ndarray,
               hit_id: np.ndarray | None = None, hit_t: np.ndarray | None = None, progress=None) -> PtStats:
        """pt_render with HOST buffers: H2D of the background, kernels, D2H of the outputs."""
        stats = PtStats()
        cb = _ffi.PROGRESS_FN(lambda _user, n: progress(n)) if progress else None
        check(gpu.pt_render(self._h, C.byref(camera), C.byref(params), background.ctypes.data, rgb.ctypes.data,
                            hit_id.ctypes.data if hit_id is not None else None,
                            hit_t.ctypes.data if hit_t is not None else None,
                            C.cast(cb, C.c_void_p) if cb else None, None, C.byref(stats)))
        return stats

    def trace_rays(self, origins: np.ndarray, dirs: np.ndarray, background=(0.0, 0.0, 0.0), rng: str | int = "fixed",
                   seed: int = 1, max_depth: int = 0, flags: int = 0):
        """Ray::color for explicit rays -> (color [n,3] f64, hit_id [n,2] u32, hit_t [n] f64, stats)."""
        origins = np.ascontiguousarray(origins, dtype=np.float64)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64)
        n = origins.shape[0]
        bg = np.asarray(background, dtype=np.float64)
        color = np.empty((n, 3), np.float64)
        hit_id = np.empty((n, 2), np.uint32)
        hit_t = np.empty(n, np.float64)
        stats = PtStats()
        check(gpu.pt_trace_rays(self._h, n, origins.ctypes.data, dirs.ctypes.data, bg.ctypes.data, rng_mode_of(rng), seed,
                                max_depth, flags, color.ctypes.data, hit_id.ctypes.data, hit_t.ctypes.data, C.byref(stats)))
        return color, hit_id, hit_t, stats

    def close(self) -> None:
        if self._h:
            gpu.pt_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Frame:
    """Device-resident render target (``PtFrame``): inputs and outputs stay in HBM."""

    def __init__(self, dscene: DeviceScene, camera: PtCamera, params: PtRenderParams):
        self.dscene = dscene
        self.params = params
        self._h = C.c_void_p()
        check(gpu.pt_frame_create(dscene.handle, C.byref(camera), C.byref(params), C.byref(self._h)))
        self.owned_pixels = gpu.pt_frame_owned_pixels(self._h)
        self.background_doubles = gpu.pt_frame_background_doubles(self._h)

    def rebind(self, dscene: "DeviceScene | None" = None, camera: PtCamera | None = None, seed: int | None = None) -> None:
        """point this frame at another scene / camera of the same image geometry (pt_frame_rebind)"""
        check(gpu.pt_frame_rebind(self._h, dscene.handle if dscene is not None else None,
                                  C.byref(camera) if camera is not None else None,
                                  C.byref(C.c_uint64(seed)) if seed is not None else None, None))
        if dscene is not None:
            self.dscene = dscene

    def set_background(self, background: np.ndarray) -> None:
        background = np.ascontiguousarray(background, dtype=np.float64)
        assert background.size == self.background_doubles, (background.size, self.background_doubles)
        check(gpu.pt_frame_set_background(self._h, background.ctypes.data))

    def set_background_device(self, device_ptr: int) -> None:
        check(gpu.pt_frame_set_background_device(self._h, C.c_void_p(device_ptr)))

    def set_image_target(self, device_ptr: int | None) -> None:
        """the resolve kernel also stores this rank's pixels into a full-image RGB8 buffer (W*H*3 bytes, local or peer
        device memory, distributed.PeerImage): the multi-GPU exchange without a gather.  None detaches."""
        check(gpu.pt_frame_set_image_target(self._h, C.c_void_p(device_ptr) if device_ptr else None))

    def render(self, stream: int | None = None, progress=None) -> PtStats:
        stats = PtStats()
        cb = _ffi.PROGRESS_FN(lambda _user, n: progress(n)) if progress else None
        check(gpu.pt_frame_render(self._h, C.c_void_p(stream) if stream else None,
                                  C.cast(cb, C.c_void_p) if cb else None, None, C.byref(stats)))
        return stats

    def enqueue(self, stream: int | None = None) -> None:
        """put the whole frame on the stream without waiting (pt_frame_enqueue); pair with finish()"""
        check(gpu.pt_frame_enqueue(self._h, C.c_void_p(stream) if stream else None))

    def finish(self) -> PtStats:
        stats = PtStats()
        check(gpu.pt_frame_finish(self._h, C.byref(stats)))
        return stats

    def pixel_index(self) -> np.ndarray:
        out = np.empty(self.owned_pixels, np.uint32)
        check(gpu.pt_frame_pixel_index(self._h, out.ctypes.data))
        return out

    @property
    def rgb_device_ptr(self) -> int:
        return gpu.pt_frame_rgb_device(self._h)

    @property
    def hit_id_device_ptr(self) -> int:
        return gpu.pt_frame_hit_id_device(self._h)

    def read(self, rgb: np.ndarray | None = None, hit_id: np.ndarray | None = None, hit_t: np.ndarray | None = None) -> PtStats:
        stats = PtStats()
        check(gpu.pt_frame_read(self._h, rgb.ctypes.data if rgb is not None else None,
                                hit_id.ctypes.data if hit_id is not None else None,
                                hit_t.ctypes.data if hit_t is not None else None, C.byref(stats)))
        return stats

    def encode_png(self) -> bytes:
        """the picture of the last render as a PNG FILE, assembled on the device (pt_frame_encode_png; world <= 1)"""
        n = C.c_uint64(0)
        gpu.pt_frame_encode_png(self._h, None, 0, C.byref(n))  # size query (fails with "too small" and fills n)
        buf = np.empty(n.value, np.uint8)
        check(gpu.pt_frame_encode_png(self._h, buf.ctypes.data, buf.nbytes, C.byref(n)))
        return buf[: n.value].tobytes()

    def close(self) -> None:
        if self._h:
            gpu.pt_frame_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Image:
    """``Image`` of src/render.rs:154-224: an RGB8 buffer that scenes are rendered onto."""

    def __init__(self, width: int, height: int, path: str | None = None):
        self.path = path
        self.buffer = np.zeros((height, width, 3), np.uint8)
        self.hit_id = None
        self.hit_t = None
        self.stats: PtStats | None = None
        if path and os.path.exists(path):  # keep the pixels of an existing image of equal size (render.rs:165-176)
            from PIL import Image as PILImage

            with PILImage.open(path) as im:
                arr = np.asarray(im.convert("RGB"))
            if arr.shape == self.buffer.shape:
                self.buffer = np.ascontiguousarray(arr)

    @property
    def width(self) -> int:
        return self.buffer.shape[1]

    @property
    def height(self) -> int:
        return self.buffer.shape[0]

    def save(self, path: str | None = None) -> None:
        path = path or self.path
        if str(path).lower().endswith(".png") and os.environ.get("PORTRAYER_PNG", "device") == "device" and gpu.pt_device_count() > 0:
            with open(path, "wb") as f:  # Image::save, render.rs:200-208: the file is assembled on the device
                f.write(png_encode(self.buffer))
            return
        from PIL import Image as PILImage

        PILImage.fromarray(self.buffer).save(path)

    def render(self, scene: Scene, samples: int | None = None, rng: str | int = "hash", seed: int = 1, slice_=None,
               want_hit_ids: bool = False, progress=None, flags: int = 0, dscene: DeviceScene | None = None,
               **tuning) -> PtStats:
        """Render ``scene`` onto this image (all of it, or the inclusive ``slice_`` = (x1, y1, x2, y2):
        ``slice_mut`` of render.rs:211-213).  ``samples`` defaults to the SAMPLES env var."""
        if slice_ is not None:
            x1, y1, x2, y2 = slice_
            if x1 >= self.width or y1 >= self.height or x2 >= self.width or y2 >= self.height:
                raise IndexError(  # render.rs:83-86 panics with this text
                    f"The positions {{x: {x1}, y: {y1}}} and/or {{x: {x2}, y: {y2}}} are not within an image with "
                    f"width = {self.width} and height = {self.height}")
        samples = samples if samples is not None else samples_from_env()
        bg, bg_mode = _background_arg(scene, self.width, self.height)
        params = make_params(self.width, self.height, samples, rng, seed, slice_, bg_mode, flags=flags, **tuning)
        own = dscene is None
        dscene = dscene or DeviceScene(scene.blob)
        try:
            if want_hit_ids:
                self.hit_id = np.full((self.height, self.width, 2), 0xFFFFFFFF, np.uint32)
                self.hit_t = np.full((self.height, self.width), np.inf, np.float64)
            self.stats = dscene.render(scene.camera(self.width, self.height), params, bg, self.buffer,
                                       self.hit_id if want_hit_ids else None, self.hit_t if want_hit_ids else None,
                                       progress)
        finally:
            if own:
                dscene.close()
        return self.stats
