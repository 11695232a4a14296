"""ctypes bindings of the product libraries.

* ``libportrayer_gpu.so``    — the C ABI of ``include/portrayer_gpu.h`` (CUDA, sm_100a).
* ``libportrayer_host.so``   — the C++ host mirror of the reference's scene API (``portrayer_b200/host/capi.h``):
  scene programs, flatten, kd build, packing.  GPU-free: it links only ``libportrayer_blob.so`` (the pure-host
  pack / unpack / tile-ownership part of the C ABI).
* ``libportrayer_render.so`` — ``Image::render`` of the host mirror, the one part of it that calls the GPU library.

There is no CPU fallback: if the libraries are not built the import raises.  ``PORTRAYER_NO_GPU_LIB=1`` (set by
``bench.py --impl reference``, which times the CPU oracle and must not map the CUDA library) loads the GPU-free
libraries only; every ``gpu.*`` call then fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIB_DIR = os.environ.get("PORTRAYER_LIB_DIR") or os.path.join(_HERE, "lib")  # override: A/B builds of the native libraries


def _load(name: str) -> C.CDLL:
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the native libraries first "
            f"(`make -C {REPO_ROOT}` or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "portrayer_b200 has no CPU fallback."
        )
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


class _NoGpuLibrary:
    """stands in for the CUDA library in a PORTRAYER_NO_GPU_LIB process: any use is an error, never a fallback"""

    def __getattr__(self, name):
        raise RuntimeError(f"libportrayer_gpu.so is not loaded in this process (PORTRAYER_NO_GPU_LIB is set): {name} is unavailable")


NO_GPU_LIB = os.environ.get("PORTRAYER_NO_GPU_LIB", "") not in ("", "0")
blob = _load("libportrayer_blob.so")
host = _load("libportrayer_host.so")
gpu = _NoGpuLibrary() if NO_GPU_LIB else _load("libportrayer_gpu.so")
host_render = None if NO_GPU_LIB else _load("libportrayer_render.so")

# ----------------------------------------------------------------------------- constants (portrayer_gpu.h)
PT_OK = 0
PT_ERR_INVALID, PT_ERR_CUDA = -1, -2
PT_ERR_NO_TEXCOORD_NORMALMAP, PT_ERR_NO_TEXCOORD_TEXTURE = -3, -4
PT_ERR_KD_PLANE_MISS, PT_ERR_TIR_INSIDE, PT_ERR_KD_TOO_DEEP, PT_ERR_OVERFLOW = -5, -6, -7, -8
PT_RNG_FIXED, PT_RNG_HASH = 0, 1
PT_BG_PER_PIXEL, PT_BG_PER_ROW, PT_BG_CONSTANT = 0, 1, 2
PT_RENDER_COUNTERS, PT_RENDER_LINEAR_TLAS, PT_RENDER_KERNEL_TIMES = 1, 2, 4
PT_RENDER_ROW_MAJOR, PT_RENDER_NO_GRAPH = 8, 16
PT_RENDER_TOLERATE_KD_PLANE = 32
PT_RENDER_EXACT_WALK = 64
PT_PIXELS_LUMA8, PT_PIXELS_LUMAA8, PT_PIXELS_RGB8, PT_PIXELS_RGBA8, PT_PIXELS_BGR8, PT_PIXELS_BGRA8 = 1, 2, 3, 4, 5, 6
PT_DEVERR_NORMALMAP, PT_DEVERR_TEXTURE, PT_DEVERR_KD_PLANE, PT_DEVERR_TIR, PT_DEVERR_OVERFLOW = 1, 2, 4, 8, 16
PT_EPSILON = 0.00001
PT_MAX_RECURSION_DEPTH = 10
PT_DEFAULT_SAMPLES = 100
NO_HIT = 0xFFFFFFFF


class PtCamera(C.Structure):
    _fields_ = [
        ("eye", C.c_double * 3),
        ("view_to_world", C.c_double * 16),
        ("fov_factor", C.c_double),
        ("aspect_ratio", C.c_double),
        ("width", C.c_double),
        ("height", C.c_double),
    ]


class PtRenderParams(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("x1", C.c_uint32), ("y1", C.c_uint32), ("x2", C.c_uint32), ("y2", C.c_uint32),
        ("samples", C.c_uint32), ("rng_mode", C.c_uint32),
        ("seed", C.c_uint64),
        ("bg_mode", C.c_uint32), ("max_depth", C.c_uint32),
        ("tile_w", C.c_uint32), ("tile_h", C.c_uint32),
        ("rank", C.c_uint32), ("world", C.c_uint32),
        ("max_batch_paths", C.c_uint64), ("node_pool_capacity", C.c_uint64),
        ("flags", C.c_uint32), ("reserved", C.c_uint32),
    ]


class PtStats(C.Structure):
    _fields_ = [
        ("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_reflect", C.c_uint64),
        ("rays_refract", C.c_uint64), ("rays_depth_cut", C.c_uint64),
        ("kd_splits", C.c_uint64), ("instance_tests", C.c_uint64), ("triangle_tests", C.c_uint64),
        ("bbox_gates", C.c_uint64), ("shaded_hits", C.c_uint64), ("texel_lookups", C.c_uint64),
        ("nodes_total", C.c_uint64),
        ("batches", C.c_uint32), ("retries", C.c_uint32), ("max_level", C.c_uint32),
        ("device_error_bits", C.c_uint32), ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32),
        ("device_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("k_kd_splits", C.c_uint64 * 2), ("k_instance_tests", C.c_uint64 * 2), ("k_triangle_tests", C.c_uint64 * 2),
        ("k_bbox_gates", C.c_uint64 * 2),
        ("ms_extend", C.c_double), ("ms_shadow", C.c_double), ("ms_shade", C.c_double),
        ("n_extend", C.c_uint32), ("n_shadow", C.c_uint32), ("n_shade", C.c_uint32), ("reserved2", C.c_uint32),
        ("k_prim_flops", C.c_uint64 * 2),
        ("err_bit", C.c_uint32), ("err_pixel", C.c_uint32), ("err_sample", C.c_uint32), ("err_pathid", C.c_uint32),
        ("err_where", C.c_uint32), ("reserved3", C.c_uint32),
        ("ms_extend_level", C.c_double * 16), ("ms_shadow_level", C.c_double * 16),
        ("x_box_tests", C.c_uint64 * 2), ("x_instance_tests", C.c_uint64 * 2), ("x_triangle_tests", C.c_uint64 * 2),
        ("x_bbox_gates", C.c_uint64 * 2), ("x_prim_flops", C.c_uint64 * 2),
    ]

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            if name.startswith("reserved"):
                continue
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out

    @property
    def rays(self) -> int:
        """every ray_cast issued against the scene root"""
        return self.rays_primary + self.rays_shadow + self.rays_reflect + self.rays_refract


class PtBlobHeader(C.Structure):
    _fields_ = [
        ("magic", C.c_uint32), ("version", C.c_uint32), ("total_bytes", C.c_uint64),
        ("ambient", C.c_double * 3), ("tlas_extent", C.c_double),
        ("tlas_depth", C.c_uint32), ("blas_max_depth", C.c_uint32),
        ("n_tlas_nodes", C.c_uint32), ("n_tlas_items", C.c_uint32), ("n_instances", C.c_uint32), ("n_meshes", C.c_uint32),
        ("n_blas_nodes", C.c_uint32), ("n_blas_items", C.c_uint32), ("n_triangles", C.c_uint32), ("n_tri_normals", C.c_uint32),
        ("n_tri_uvs", C.c_uint32), ("n_materials", C.c_uint32), ("n_lights", C.c_uint32), ("n_textures", C.c_uint32),
        ("n_texel_bytes", C.c_uint64),
        ("off_tlas_nodes", C.c_uint64), ("off_tlas_items", C.c_uint64), ("off_instances", C.c_uint64),
        ("off_instance_trans", C.c_uint64), ("off_meshes", C.c_uint64), ("off_blas_nodes", C.c_uint64),
        ("off_blas_items", C.c_uint64), ("off_tri_pos", C.c_uint64), ("off_tri_normals", C.c_uint64),
        ("off_tri_uvs", C.c_uint64), ("off_materials", C.c_uint64), ("off_lights", C.c_uint64),
        ("off_textures", C.c_uint64), ("off_texels", C.c_uint64),
    ]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64)
TEXTURE_LOADER_FN = C.CFUNCTYPE(None, C.c_char_p)

# every symbol include/portrayer_gpu.h declares, with its signature
GPU_SYMBOLS = {
    "pt_init": (C.c_int, [C.c_int]),
    "pt_init_devices": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "pt_texture_ingest": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]),
    "pt_texture_read": (C.c_int, [C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pt_png_size": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "pt_png_encode_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "pt_png_encode": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "pt_frame_encode_png": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "pt_device_group_size": (C.c_int, []),
    "pt_shutdown": (None, []),
    "pt_last_error": (C.c_char_p, []),
    "pt_error_string": (C.c_char_p, [C.c_int]),
    "pt_device_count": (C.c_int, []),
    "pt_release_cached_memory": (None, []),
    "pt_resident_texture_bytes": (C.c_uint64, []),
    "pt_abi_sizeof": (C.c_uint64, [C.c_int]),
    "pt_measure_fp64_rate": (C.c_int, [C.c_double, C.POINTER(C.c_double)]),
    "pt_scene_blob_size": (C.c_uint64, [C.c_void_p]),
    "pt_scene_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "pt_scene_unpack": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "pt_scene_unpack_records": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "pt_scene_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "pt_scene_upload_device": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "pt_scene_free": (None, [C.c_void_p]),
    "pt_scene_uploaded_bytes": (C.c_uint64, [C.c_void_p]),
    "pt_render": (C.c_int, [C.c_void_p, C.POINTER(PtCamera), C.POINTER(PtRenderParams), C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PtStats)]),
    "pt_trace_rays": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PtStats)]),
    "pt_owned_pixels": (C.c_uint64, [C.POINTER(PtRenderParams), C.c_void_p, C.c_uint64]),
    "pt_frame_create": (C.c_int, [C.c_void_p, C.POINTER(PtCamera), C.POINTER(PtRenderParams), C.POINTER(C.c_void_p)]),
    "pt_frame_free": (None, [C.c_void_p]),
    "pt_frame_rebind": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PtCamera), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "pt_frame_owned_pixels": (C.c_uint64, [C.c_void_p]),
    "pt_frame_background_doubles": (C.c_uint64, [C.c_void_p]),
    "pt_frame_set_background": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_frame_set_background_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_frame_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PtStats)]),
    "pt_frame_enqueue": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_frame_finish": (C.c_int, [C.c_void_p, C.POINTER(PtStats)]),
    "pt_frame_rgb_device": (C.c_void_p, [C.c_void_p]),
    "pt_frame_hit_id_device": (C.c_void_p, [C.c_void_p]),
    "pt_frame_hit_t_device": (C.c_void_p, [C.c_void_p]),
    "pt_frame_pixel_index": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_frame_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PtStats)]),
    "pt_peer_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]),
    "pt_peer_free": (C.c_int, [C.c_void_p]),
    "pt_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "pt_peer_close": (C.c_int, [C.c_void_p]),
    "pt_frame_set_image_target": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_kd_build": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "pt_kd_build_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "pt_kd_tree_free": (None, [C.c_void_p]),
    "pt_kd_tree_node_count": (C.c_uint32, [C.c_void_p]),
    "pt_kd_tree_item_count": (C.c_uint32, [C.c_void_p]),
    "pt_kd_tree_depth": (C.c_uint32, [C.c_void_p]),
    "pt_kd_tree_root_bounds": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "pt_kd_tree_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pt_kd_tree_build_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "pt_scene_set_tlas": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pt_flatten": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "pt_flat_free": (None, [C.c_void_p]),
    "pt_flat_instance_count": (C.c_uint32, [C.c_void_p]),
    "pt_flat_bounds_device": (C.c_void_p, [C.c_void_p]),
    "pt_flat_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pt_flat_build_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]),
    "pt_scene_set_instances": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
}
for _name, (_res, _args) in GPU_SYMBOLS.items():
    if NO_GPU_LIB:
        break
    _fn = getattr(gpu, _name)  # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args

HOST_SYMBOLS = {
    "pth_last_error": (C.c_char_p, []),
    "pth_set_assets_dir": (None, [C.c_char_p]),
    "pth_set_baked_mesh_dir": (None, [C.c_char_p]),
    "pth_set_texture_loader": (None, [TEXTURE_LOADER_FN]),
    "pth_register_texture": (None, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "pth_bake_obj": (C.c_int, [C.c_char_p, C.c_char_p]),
    "pth_mesh_info": (C.c_int64, [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pth_example_count": (C.c_int, []),
    "pth_example_name": (C.c_char_p, [C.c_int]),
    "pth_example_build": (C.c_void_p, [C.c_char_p, C.c_int64, C.c_int]),
    "pth_big_scene_build": (C.c_void_p, [C.c_uint64, C.c_int64, C.c_int]),
    "pth_synthetic_instances_build": (C.c_void_p, [C.c_uint64, C.c_uint64, C.c_int64]),
    "pth_synthetic_triangles_build": (C.c_void_p, [C.c_uint64, C.c_uint64, C.c_int64]),
    "pth_scene_free": (None, [C.c_void_p]),
    "pth_blob_size": (C.c_uint64, [C.c_void_p]),
    "pth_blob_data": (C.c_void_p, [C.c_void_p]),
    "pth_image_size": (None, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pth_camera": (None, [C.c_void_p, C.c_double, C.c_double, C.POINTER(PtCamera)]),
    "pth_background": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "pth_prepare_seconds": (C.c_double, [C.c_void_p]),
    "pth_flatten_seconds": (C.c_double, [C.c_void_p]),
    "pth_scene_item_count": (C.c_uint64, [C.c_void_p]),
    "pth_scene_item_bounds": (None, [C.c_void_p, C.c_void_p]),
    "pth_scene_hierarchy_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "pth_scene_hierarchy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pth_kd_build": (C.c_void_p, [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32]),
    "pth_kd_tree_free": (None, [C.c_void_p]),
    "pth_kd_tree_node_count": (C.c_uint64, [C.c_void_p]),
    "pth_kd_tree_item_count": (C.c_uint64, [C.c_void_p]),
    "pth_kd_tree_depth": (C.c_uint32, [C.c_void_p]),
    "pth_kd_tree_extent": (C.c_double, [C.c_void_p]),
    "pth_kd_tree_build_seconds": (C.c_double, [C.c_void_p]),
    "pth_kd_tree_nodes": (C.c_void_p, [C.c_void_p]),
    "pth_kd_tree_items": (C.c_void_p, [C.c_void_p]),
    "pth_image_render": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p,
                                   C.POINTER(PtStats)]),
}
for _name, (_res, _args) in HOST_SYMBOLS.items():
    if _name == "pth_image_render":
        if host_render is None:
            continue
        _fn = getattr(host_render, _name)
    else:
        _fn = getattr(host, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class PortrayerError(RuntimeError):
    """A non-zero PtError; the message is the reference's panic text where the reference would panic."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


def check(rc: int) -> None:
    if rc != PT_OK:
        raise PortrayerError(rc, (gpu.pt_last_error() or b"").decode() or (gpu.pt_error_string(rc) or b"").decode())
