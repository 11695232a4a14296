"""portrayer_b200 — B200-native implementation of portrayer's per-pixel render loop.

Only what the hot path needs lives here:

* ``csrc/``  hand-written CUDA kernels for sm_100a + the C ABI (``include/portrayer_gpu.h``)
* ``host/``  C++ mirror of the reference's scene API (stays Rust in the target design)
* this Python layer: ctypes bindings, the ``Image.render`` entry point, multi-GPU plumbing.
"""
from ._ffi import (PT_RENDER_COUNTERS, PT_RENDER_LINEAR_TLAS, PT_RNG_FIXED, PT_RNG_HASH, PortrayerError, PtCamera, PtRenderParams,  # noqa: F401
                   PtStats)
from .render import (DeviceScene, Frame, Image, init_devices, make_params, png_encode, samples_from_env,  # noqa: F401
                     texture_ingest, texture_read)
from .scene import Scene, example_names  # noqa: F401

__all__ = ["Scene", "example_names", "Image", "DeviceScene", "Frame", "make_params", "samples_from_env", "init_devices", "png_encode", "texture_ingest", "texture_read", "PtStats",
           "PtCamera", "PtRenderParams", "PortrayerError", "PT_RNG_FIXED", "PT_RNG_HASH", "PT_RENDER_COUNTERS",
           "PT_RENDER_LINEAR_TLAS"]
