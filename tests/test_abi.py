"""The C-ABI library without a GPU: it loads, exports every symbol include/portrayer_gpu.h declares, and its
pure-host entry points (blob pack / unpack / validation, tile ownership, error strings) behave.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(REPO, "include", "portrayer_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native_libraries):
    from portrayer_b200 import _ffi

    names = _declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(_ffi.gpu, name), f"{name} is declared in include/portrayer_gpu.h but not exported"
        assert name in _ffi.GPU_SYMBOLS, f"{name} has no ctypes signature in portrayer_b200/_ffi.py"


def test_struct_sizes_match_the_header(native_libraries):
    """ctypes mirrors vs the sizes the header documents (a mismatch would silently corrupt the blob)."""
    from portrayer_b200 import _ffi

    assert C.sizeof(_ffi.PtCamera) == 23 * 8
    assert C.sizeof(_ffi.PtRenderParams) == 88
    assert C.sizeof(_ffi.PtBlobHeader) % 8 == 0
    # and against the compiled library itself
    assert _ffi.gpu.pt_abi_sizeof(0) == C.sizeof(_ffi.PtCamera)
    assert _ffi.gpu.pt_abi_sizeof(1) == C.sizeof(_ffi.PtRenderParams)
    assert _ffi.gpu.pt_abi_sizeof(2) == C.sizeof(_ffi.PtStats)
    assert _ffi.gpu.pt_abi_sizeof(3) == C.sizeof(_ffi.PtBlobHeader)
    assert _ffi.gpu.pt_abi_sizeof(99) == 0


def test_error_strings_are_the_reference_panics(native_libraries):
    from portrayer_b200 import _ffi

    s = lambda code: _ffi.gpu.pt_error_string(code).decode()
    assert s(_ffi.PT_ERR_NO_TEXCOORD_NORMALMAP) == "Normal/Texture mapping is not supported for this primitive!"  # material.rs:133
    assert s(_ffi.PT_ERR_NO_TEXCOORD_TEXTURE) == "Texture mapping is not supported for this primitive!"  # material.rs:141
    assert s(_ffi.PT_ERR_KD_PLANE_MISS) == "bug: ray should definitely hit infinite plane"  # kdtree/node.rs:147,178
    assert s(_ffi.PT_ERR_TIR_INSIDE) == "bug: should not have total internal reflection when casting inside surface"  # material.rs:258


def test_blob_unpack_validates(native_libraries):
    import portrayer_b200 as pt
    from portrayer_b200 import _ffi

    scene = pt.Scene.example("nonhier")
    blob = scene.blob.copy()
    desc = (C.c_uint8 * 512)()  # PtSceneDesc is < 512 bytes
    assert _ffi.gpu.pt_scene_unpack(blob.ctypes.data, blob.nbytes, desc) == 0
    # records-only view: stops at the texel section
    h = scene.header
    assert _ffi.gpu.pt_scene_unpack_records(blob.ctypes.data, h.off_texels, desc) == 0
    assert _ffi.gpu.pt_scene_unpack(blob.ctypes.data, h.off_texels - 1, desc) != 0
    # truncated
    assert _ffi.gpu.pt_scene_unpack(blob.ctypes.data, blob.nbytes - 1, desc) != 0
    # bad magic
    bad = blob.copy()
    bad[0] ^= 0xFF
    assert _ffi.gpu.pt_scene_unpack(bad.ctypes.data, bad.nbytes, desc) != 0
    # an instance index out of range in a leaf
    bad = blob.copy()
    items = np.frombuffer(bad, dtype=np.uint32, count=h.n_tlas_items, offset=h.off_tlas_items)
    items[0] = h.n_instances
    assert _ffi.gpu.pt_scene_unpack(bad.ctypes.data, bad.nbytes, desc) != 0
    # a kd child pointing backwards (cycle)
    bad = blob.copy()
    nodes = np.frombuffer(bad, dtype=np.uint32, count=h.n_tlas_nodes * 4, offset=h.off_tlas_nodes).reshape(-1, 4)
    split = np.where((nodes[:, 2] & 3) != 3)[0]
    if len(split):
        nodes[split[0], 3] = 0
        assert _ffi.gpu.pt_scene_unpack(bad.ctypes.data, bad.nbytes, desc) != 0
    # a material pointing at a texture that does not exist
    bad = blob.copy()
    mats = np.frombuffer(bad, dtype=np.int32, count=h.n_materials * 40, offset=h.off_materials).reshape(-1, 40)
    mats[0, 38] = 5
    assert _ffi.gpu.pt_scene_unpack(bad.ctypes.data, bad.nbytes, desc) != 0


def test_no_cpu_fallback_without_a_device(native_libraries):
    """On a machine without a GPU the library must refuse to render, loudly (this container has none)."""
    import torch

    from portrayer_b200 import _ffi

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _ffi.gpu.pt_device_count() == 0
    assert _ffi.gpu.pt_init(0) == _ffi.PT_ERR_CUDA
    assert b"no CPU fallback" in _ffi.gpu.pt_last_error()
    import portrayer_b200 as pt

    with pytest.raises(pt.PortrayerError):
        pt.DeviceScene(pt.Scene.example("nonhier").blob)


def test_product_code_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import / link / execute oracle/."""
    pattern = re.compile(r"import\s+oracle|from\s+oracle|liboracle|oracle\.h|oracle/oracle|oracle_render|oracle_trace")
    offenders = []
    for root in ("portrayer_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(REPO, root)):
            for f in files:
                if f.endswith((".py", ".c", ".cpp", ".cu", ".cuh", ".h", ".hpp")):
                    if pattern.search(open(os.path.join(dirpath, f), errors="ignore").read()):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_header_is_plain_c(tmp_path):
    """include/portrayer_gpu.h is what a cgo / bindgen / ctypes consumer compiles: it must be valid C99 on its own
    (no C++ types, no torch, no CUDA headers) and valid C++ for the host mirror."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "portrayer_gpu.h"\nint main(void) { PtPeerHandle h; PtStats s; (void)h; (void)s; return 0; }\n')
    inc = os.path.join(REPO, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    if shutil.which("g++"):
        subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)


def test_png_size_is_the_stored_deflate_layout(native_libraries):
    """pt_png_size is pure arithmetic (no device): signature + IHDR + one IDAT of stored deflate blocks + IEND, the layout
    tests/test_gpu_image_io.py checks byte for byte on the device"""
    from portrayer_b200 import _ffi

    for w, h in ((1, 1), (7, 5), (1456, 15), (910, 512), (3840, 2160)):
        raw = h * (1 + 3 * w)
        blocks = max(1, -(-raw // 65535))
        assert _ffi.gpu.pt_png_size(w, h) == 8 + 25 + 12 + (2 + 5 * blocks + raw + 4) + 12
    assert _ffi.gpu.pt_png_size(0, 5) == 0


def test_integration_doc_matches_the_abi(native_libraries):
    """INTEGRATION.md's Rust mirror states PtStats' size: keep it in step with the header"""
    import ctypes as C

    from portrayer_b200 import _ffi

    n = C.sizeof(_ffi.PtStats)
    assert _ffi.gpu.pt_abi_sizeof(2) == n
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    assert f"_opaque: [u64; {n // 8}]" in doc and f"({n} bytes" in doc


def test_device_group_members_partition_the_callers_tiles(native_libraries):
    """pt_render over a device group gives member i of n rank r + w*i of world w*n (api.cu render_group), r / w the
    caller's own rank / world: the members' pixels must be exactly the caller's pixels, each once — with slices and
    ragged image sizes.  Pure host logic (tiles.c), no device."""
    import numpy as np

    from portrayer_b200 import _ffi
    from portrayer_b200.render import make_params

    def owned(p):
        n = _ffi.gpu.pt_owned_pixels(p, None, 0)
        out = np.empty(n, np.uint32)
        assert _ffi.gpu.pt_owned_pixels(p, out.ctypes.data, n) == n
        return out

    for (w, h, slice_, tile) in ((200, 120, None, 16), (333, 77, (5, 3, 300, 70), 32), (64, 64, None, 32)):
        for world in (1, 2, 3):
            for rank in range(world):
                mine = owned(make_params(w, h, 1, "hash", 1, slice_, rank=rank, world=world, tile=tile))
                for n in (2, 8):
                    parts = [owned(make_params(w, h, 1, "hash", 1, slice_, rank=rank + world * i, world=world * n, tile=tile)) for i in range(n)]
                    union = np.concatenate(parts)
                    assert len(union) == len(mine) and np.array_equal(np.sort(union), np.sort(mine))


def test_python_constants_match_the_header(native_libraries):
    """the ctypes layer restates the header's #defines: every PT_RENDER_* / PT_PIXELS_* value must agree"""
    from portrayer_b200 import _ffi

    text = open(os.path.join(REPO, "include", "portrayer_gpu.h")).read()
    found = dict(re.findall(r"#define\s+(PT_(?:RENDER|PIXELS)_[A-Z0-9_]+)\s+(\d+)u", text))
    assert len(found) >= 13, found
    for name, value in found.items():
        assert getattr(_ffi, name) == int(value), name
