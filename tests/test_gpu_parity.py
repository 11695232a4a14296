"""Parity of the CUDA path (through the C ABI) against the CPU oracle. -m gpu."""
import numpy as np
import pytest

import parity
import portrayer_b200 as pt
from conftest import has_reference_assets
from oracle import binding as oracle

pytestmark = pytest.mark.gpu


def _report(name, **kw):
    scene = pt.Scene.example(name)
    img, stats = parity.render_gpu(scene, **kw)
    ref = parity.render_oracle(scene, **kw)
    assert ref.rc == 0
    rep = parity.compare(img, ref, name)
    rep["rays"] = stats.rays
    print(rep)
    # the device and the oracle issue exactly the same rays
    assert stats.rays_primary == ref.stats.rays_primary
    assert stats.rays_shadow == ref.stats.rays_shadow
    assert stats.rays_reflect == ref.stats.rays_reflect
    assert stats.rays_refract == ref.stats.rays_refract
    assert stats.rays_depth_cut == ref.stats.rays_depth_cut
    return rep


# BASELINE.json configs[0]: nonhier at its native resolution, SAMPLES=1, pixel-centre rays
def test_nonhier_native(gpu_ready):
    rep = _report("nonhier", samples=1, rng="fixed")
    parity.assert_parity(rep)
    assert rep["max_lsb_diff"] <= 1


# configs[1]: all analytic primitives, instancing, a reflective dome
def test_primitives_native(gpu_ready):
    parity.assert_parity(_report("primitives", samples=1, rng="fixed"))


@pytest.mark.skipif(not has_reference_assets(), reason="reference textures not synced (tools/sync_assets.py)")
@pytest.mark.parametrize("name", ["texture-mapping", "normal-mapping", "normal-mapping-left", "normal-mapping-right"])
def test_textured_scenes(gpu_ready, name):
    parity.assert_parity(_report(name, samples=1, rng="fixed"))


# the other example programs of the reference (examples/*.rs beyond BASELINE.json's configs): deep hierarchies,
# instancing through shared nodes, a Triangle as a primitive, flat + smooth + textured meshes, a dielectric glossy mesh,
# constant backgrounds.  Native size for the small ones, half size where a linear 5 804-triangle Mesh makes the oracle slow.
@pytest.mark.parametrize("name,samples,rng,scale", [
    ("simple", 1, "fixed", 1), ("hier", 1, "fixed", 1), ("instance", 1, "fixed", 1), ("nonhier2", 1, "fixed", 1),
    ("macho-cows", 1, "fixed", 1), ("simple-cows", 1, "fixed", 1), ("single-triangle", 1, "fixed", 1),
    ("smooth-shading", 1, "hash", 2), ("primitives-simple", 1, "fixed", 1), ("four-shapes", 1, "fixed", 2),
    ("antialiasing", 4, "hash", 1), ("fish", 2, "hash", 2), ("graphics-poster", 2, "hash", 1),
    ("cube-mapping", 1, "fixed", 1), ("entering-the-mirror-dimension", 2, "hash", 2), ("transmission-refraction", 2, "hash", 2),
    ("robot-alarm-clock", 2, "hash", 4),
    # the last two of the reference's 28 examples: two area lights + 8 linear Meshes (3 uses of monkey.obj, one a flat
    # dielectric), a cube-mapped texture on a Cube, glass ball; 5 KDMeshes (one instanced 3x), glossy dielectric lake
    ("monkeys-making-monkeys", 2, "hash", 6), ("graphics-temple", 2, "hash", 1),
])
def test_more_example_scenes(gpu_ready, name, samples, rng, scale):
    if name in ("fish", "cube-mapping", "transmission-refraction", "robot-alarm-clock", "monkeys-making-monkeys", "graphics-temple") and not has_reference_assets():
        pytest.skip("reference textures not synced")
    scene = pt.Scene.example(name)
    rep = _report(name, samples=samples, rng=rng, size=(scene.width // scale, scene.height // scale))
    parity.assert_parity(rep)
    assert rep["hit_t_bit_identical"]


# degenerate inputs (host/examples/kats.cpp): no geometry at all, no lights, zero-area / zero-scale primitives whose
# arithmetic runs through NaN and inf; at an image size that is neither a multiple of the 32x32 tile nor of the 8x4
# warp block, and at 1x1 and 1-pixel-wide images
@pytest.mark.parametrize("name", ["edge-empty", "edge-no-lights", "edge-degenerate"])
@pytest.mark.parametrize("size", [None, (1, 1), (37, 1), (1, 53)])
def test_degenerate_scenes_and_ragged_images(gpu_ready, name, size):
    scene = pt.Scene.example(name)
    kw = dict(samples=3, rng="hash", size=size or (scene.width, scene.height))
    img, stats = parity.render_gpu(scene, **kw)
    ref = parity.render_oracle(scene, **kw)
    assert ref.rc == 0
    rep = parity.compare(img, ref, name)
    assert rep["rgb_exact"] == 1.0 and rep["hit_id_mismatches"] == 0 and rep["hit_t_bit_identical"], rep
    assert stats.rays == ref.stats.rays
    if name == "edge-empty":
        assert stats.rays_shadow == 0 and np.all(img.hit_id[..., 0] == 0xFFFFFFFF)
    if name == "edge-no-lights":
        assert stats.rays_shadow == 0 and stats.rays_primary > 0


# configs[2]: kd-tree traversal stress, at reduced size so the oracle finishes in seconds
@pytest.mark.parametrize("kd_depth", [10, 18])
def test_big_scene(gpu_ready, kd_depth):
    scene = pt.Scene.big_scene(10, kd_depth=kd_depth)
    kw = dict(samples=1, rng="fixed", size=(495, 255))
    img, stats = parity.render_gpu(scene, **kw)
    ref = parity.render_oracle(scene, **kw)
    rep = parity.compare(img, ref, f"big-scene kd{kd_depth}")
    print(rep)
    parity.assert_parity(rep)
    assert stats.rays == ref.stats.rays


# configs[2], synthetic half: N random instances on a grid (TLAS stress, deep scene tree) and one KDMesh of N random
# triangles (BLAS stress), generator of SURVEY 8d M3b, trees as deep as ceil(log2(N / 3))
@pytest.mark.parametrize("kind,n,depth", [("instances", 10_000, 12), ("instances", 100_000, 16), ("triangles", 100_000, 16)])
def test_synthetic_kd_stress(gpu_ready, kind, n, depth):
    scene = pt.Scene.synthetic_instances(n, kd_depth=depth) if kind == "instances" else pt.Scene.synthetic_triangles(n, kd_mesh_depth=depth)
    assert max(scene.header.tlas_depth, scene.header.blas_max_depth) == depth
    kw = dict(samples=1, rng="hash", size=(330, 170))
    img, stats = parity.render_gpu(scene, **kw)
    ref = parity.render_oracle(scene, **kw)
    assert ref.rc == 0
    rep = parity.compare(img, ref, f"synthetic {kind} {n}")
    print(rep)
    parity.assert_parity(rep)
    assert rep["hit_t_bit_identical"]
    assert stats.rays == ref.stats.rays


# The work counters of the closest-hit kernel (kd splits visited, instances / triangles tested, bbox gates) are the
# reference's algorithmic work for primary + reflected + refracted rays: they must equal the oracle's, also where the
# device skips work it can prove useless (FP32 box culls of instances and of triangles inside Mesh / KDMesh folds).
# Shadow rays are any-hit on the device (the reference runs a full closest-hit cast and only looks at is_some(),
# material.rs:174-179), so there the device may only count LESS.
@pytest.mark.parametrize("name", ["nonhier", "soft-shadows", "primitives", "glossy-reflection", "kat-mesh-equivalence-kdmesh"])
def test_work_counters_match_the_reference_algorithm(gpu_ready, name):
    from portrayer_b200 import _ffi

    scene = pt.Scene.example(name)
    kw = dict(samples=1, rng="hash", size=(200, 120))
    img, stats = parity.render_gpu(scene, flags=_ffi.PT_RENDER_COUNTERS, **kw)
    ref = parity.render_oracle(scene, **kw)
    parity.assert_parity(parity.compare(img, ref, name))
    assert np.array_equal(img.hit_id, ref.hit_id) and np.array_equal(img.hit_t, ref.hit_t)
    r = ref.stats
    assert stats.k_kd_splits[0] == r.kd_splits - r.shadow_kd_splits
    assert stats.k_instance_tests[0] == r.instance_tests - r.shadow_instance_tests
    assert stats.k_triangle_tests[0] == r.triangle_tests - r.shadow_triangle_tests
    assert stats.k_bbox_gates[0] == r.bbox_gates - r.shadow_bbox_gates
    assert stats.k_kd_splits[1] <= r.shadow_kd_splits and stats.k_instance_tests[1] <= r.shadow_instance_tests
    assert stats.k_triangle_tests[1] <= r.shadow_triangle_tests and stats.k_bbox_gates[1] <= r.shadow_bbox_gates
    assert stats.k_instance_tests[0] > 0 and stats.k_instance_tests[1] > 0


# configs[3]: secondary-ray-bound recursion with hashed jitter, glossy + area-light + dielectric RNG dimensions
@pytest.mark.parametrize("name,samples", [("glossy-reflection", 4), ("soft-shadows", 2)])
def test_secondary_ray_scenes(gpu_ready, name, samples):
    parity.assert_parity(_report(name, samples=samples, rng="hash", size=(455, 256)))


@pytest.mark.skipif(not has_reference_assets(), reason="reference textures not synced (tools/sync_assets.py)")
def test_water_glass(gpu_ready):
    parity.assert_parity(_report("water-glass", samples=4, rng="hash", size=(455, 256)))


# src/kdtree/node.rs:219-352 on the device
@pytest.mark.parametrize("name,origin,direction", [
    ("kat-edge-case", (0.0, 0.5, 0.9), (0.0, 0.0, -1.0)),
    ("kat-edge-case-flipped", (0.0, 0.5, -0.9), (0.0, 0.0, 1.0)),
])
def test_ray_cast_edge_case_device(gpu_ready, name, origin, direction):
    scene = pt.Scene.example(name)
    ds = pt.DeviceScene(scene.blob)
    color, hit_id, hit_t, _ = ds.trace_rays(np.array([origin]), np.array([direction]))
    assert hit_id[0, 0] == 0 and tuple(color[0]) == (1.0, 0.0, 0.0)


# exact ties inside a linear Mesh: the device folds in Morton order (fold_order.cu) and must still return the FIRST
# listed triangle among those with bit-identical t
def test_mesh_fold_ties_device(gpu_ready):
    scene = pt.Scene.example("edge-mesh-ties")
    origins, dirs, expect_sub, expect_t = parity.mesh_tie_rays()
    _c, hit_id, hit_t, _ = pt.DeviceScene(scene.blob).trace_rays(origins, dirs)
    rc, _co, io, to, _ = oracle.trace_rays(scene.blob, origins, dirs)
    assert rc == 0
    assert np.array_equal(hit_id, io) and np.array_equal(hit_t, to)
    hit = expect_sub >= 0
    assert np.array_equal(hit_id[hit, 1].astype(np.int64), expect_sub[hit]) and np.array_equal(hit_t[hit], expect_t[hit])
    # and through the camera path (shadow rays: any-hit folds) like every other scene
    parity.assert_parity(_report("edge-mesh-ties", samples=2, rng="hash"))


# src/kdtree/kdmesh.rs:99-166 on the device: Mesh == KDMesh bit-exact, and both == the oracle
def test_mesh_equivalence_device(gpu_ready):
    mesh = pt.Scene.example("kat-mesh-equivalence-mesh")
    kdmesh = pt.Scene.example("kat-mesh-equivalence-kdmesh")
    n = 100000
    i = np.arange(n, dtype=np.float64)
    origins, dirs = oracle.camera_rays(mesh.camera(533, 300), np.stack([533.0 * i / n, 300.0 * i / n], axis=1))
    ca, ia, ta, _ = pt.DeviceScene(mesh.blob).trace_rays(origins, dirs)
    cb, ib, tb, _ = pt.DeviceScene(kdmesh.blob).trace_rays(origins, dirs)
    assert np.array_equal(ca, cb) and np.array_equal(ia, ib) and np.array_equal(ta, tb)
    rc, co, io, to, _ = oracle.trace_rays(mesh.blob, origins, dirs)
    assert rc == 0
    assert np.array_equal(ia, io) and np.array_equal(ta, to)
    # colours go through pow(): CUDA's and glibc's differ by <= 2 ulp
    assert np.allclose(ca, co, rtol=1e-13, atol=0)


def test_determinism_and_batch_independence(gpu_ready):
    """The image must not depend on batch size (nor, therefore, on tile/GPU partitioning)."""
    scene = pt.Scene.example("glossy-reflection")
    a, _ = parity.render_gpu(scene, samples=2, rng="hash", size=(320, 180))
    b, _ = parity.render_gpu(scene, samples=2, rng="hash", size=(320, 180), max_batch_paths=4096)
    c, sc = parity.render_gpu(scene, samples=2, rng="hash", size=(320, 180), max_batch_paths=20000, node_pool_capacity=30000)
    assert np.array_equal(a.buffer, b.buffer) and np.array_equal(a.buffer, c.buffer)
    assert np.array_equal(a.hit_id, b.hit_id)
    assert sc.batches > 1


def test_slice_only_writes_slice(gpu_ready):
    """slice_mut (render.rs:211-213): pixels outside the slice keep their old value (render.rs:136-138)."""
    scene = pt.Scene.example("nonhier")
    full, _ = parity.render_gpu(scene, samples=1, rng="fixed", size=(96, 96))
    img = pt.Image(96, 96)
    img.buffer[:] = 7
    img.render(scene, samples=1, rng="fixed", slice_=(10, 20, 50, 40))
    assert np.array_equal(img.buffer[20:41, 10:51], full.buffer[20:41, 10:51])
    mask = np.ones((96, 96), bool)
    mask[20:41, 10:51] = False
    assert np.all(img.buffer[mask] == 7)
    with pytest.raises(IndexError):
        img.render(scene, samples=1, slice_=(0, 0, 96, 10))


def test_tile_ownership_matches_single_rank(gpu_ready):
    """Interleaved tile ownership: the union of world=3 partial renders is the single-rank image."""
    scene = pt.Scene.example("primitives")
    w, h = 200, 120
    full, _ = parity.render_gpu(scene, samples=1, rng="hash", size=(w, h))
    img = pt.Image(w, h)
    for rank in range(3):
        img.render(scene, samples=1, rng="hash", rank=rank, world=3, tile=16)
    assert np.array_equal(img.buffer, full.buffer)


def test_resolve_kernel_stores_tiles_into_a_shared_image(gpu_ready):
    """The fused exchange of the multi-GPU path (pt_frame_set_image_target): frames of world=3 ranks all point at ONE
    full-image buffer allocated with pt_peer_alloc; every resolve kernel stores its own tiles at their place, and the
    result is the single-rank image — no gather, no un-tiling.  (Across processes the other ranks map the same buffer
    with pt_peer_open; bench.py --gpus N checks that path against the NCCL gather on every run.)"""
    import ctypes as C

    import torch

    from portrayer_b200 import _ffi
    from portrayer_b200 import distributed as ptd
    from portrayer_b200.render import _background_arg, make_params

    scene = pt.Scene.example("primitives")
    w, h = 200, 120
    full, _ = parity.render_gpu(scene, samples=2, rng="hash", size=(w, h))
    ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
    _ffi.check(_ffi.gpu.pt_peer_alloc(w * h * 3, C.byref(ptr), handle))
    assert any(handle), "the handle another process would open must be filled in"
    try:
        ds = pt.DeviceScene(scene.blob)
        bg, bg_mode = _background_arg(scene, w, h)
        compact = []
        for rank in range(3):
            params = make_params(w, h, 2, "hash", 1, bg_mode=bg_mode, rank=rank, world=3, tile=16)
            fr = pt.Frame(ds, scene.camera(w, h), params)
            fr.set_background(bg)
            fr.set_image_target(ptr.value)
            fr.render()
            # the compact owned-pixel output is still written, and agrees
            own = ptd.device_tensor(fr.rgb_device_ptr, (fr.owned_pixels, 3)).cpu().numpy()
            compact.append((fr.pixel_index(), own))
            fr.close()
        image = ptd.device_tensor(ptr.value, (h, w, 3)).cpu().numpy()
        assert np.array_equal(image, full.buffer)
        for index, own in compact:
            assert np.array_equal(image.reshape(-1, 3)[index.astype(np.int64)], own)
        # detaching: a later render leaves the image alone
        torch.cuda.synchronize()
        ptd.device_tensor(ptr.value, (h, w, 3)).zero_()
        params = make_params(w, h, 2, "hash", 1, bg_mode=bg_mode)
        fr = pt.Frame(ds, scene.camera(w, h), params)
        fr.set_background(bg)
        fr.set_image_target(ptr.value)
        fr.set_image_target(None)
        fr.render()
        fr.close()
        assert not ptd.device_tensor(ptr.value, (h, w, 3)).any()
        ds.close()
    finally:
        _ffi.check(_ffi.gpu.pt_peer_free(ptr))


def test_reference_panic_is_reported(gpu_ready):
    """A textured material on a cylinder panics in the reference (material.rs:141); the C ABI returns that text."""
    import ctypes as C

    from portrayer_b200 import _ffi

    scene = pt.Scene.example("water-glass") if has_reference_assets() else None
    if scene is None:
        pytest.skip("needs a textured scene")
    # patch instance 2 (the water cylinder) to use the textured table material
    blob = scene.blob.copy()
    h = scene.header
    inst = np.frombuffer(blob, dtype=np.uint8, count=h.n_instances * 128, offset=h.off_instances).view(np.uint32).reshape(h.n_instances, 32)
    prims = inst[:, 24]
    cyl = int(np.where(prims == 6)[0][0])
    mats = np.frombuffer(blob, dtype=np.uint8, count=h.n_materials * 160, offset=h.off_materials).view(np.int32).reshape(h.n_materials, 40)
    textured = int(np.where(mats[:, 38] >= 0)[0][0])
    inst[cyl, 26] = textured
    ds = pt.DeviceScene(blob)
    img = pt.Image(64, 36)
    params = pt.make_params(64, 36, 1, "fixed", bg_mode=_ffi.PT_BG_CONSTANT)
    with pytest.raises(pt.PortrayerError) as err:
        ds.render(scene.camera(64, 36), params, np.zeros(3), img.buffer)
    assert "mapping is not supported for this primitive!" in str(err.value)


def test_graph_and_stream_paths_agree(gpu_ready):
    """The frame CUDA graph (conditional WHILE over recursion levels) and the kernel-by-kernel stream path run the
    same kernels: identical images, hit ids, ray counts — also with several batches per frame."""
    from portrayer_b200 import _ffi

    scene = pt.Scene.example("glossy-reflection")
    kw = dict(samples=3, rng="hash", size=(300, 170))
    a, sa = parity.render_gpu(scene, **kw)
    b, sb = parity.render_gpu(scene, flags=_ffi.PT_RENDER_NO_GRAPH, **kw)
    c, sc = parity.render_gpu(scene, max_batch_paths=30000, **kw)  # 6 graph replays
    d, sd = parity.render_gpu(scene, flags=_ffi.PT_RENDER_KERNEL_TIMES, **kw)
    for other, st in ((b, sb), (c, sc), (d, sd)):
        assert np.array_equal(a.buffer, other.buffer) and np.array_equal(a.hit_id, other.hit_id)
        assert (st.rays_primary, st.rays_shadow, st.rays_reflect, st.rays_refract) == \
               (sa.rays_primary, sa.rays_shadow, sa.rays_reflect, sa.rays_refract)
    assert sc.batches >= 5 and sa.batches == 1
    assert sd.n_extend > 0 and sd.ms_extend > 0.0
    # only the levels that hold rays are launched: far fewer than 3 kernels x 11 levels
    assert sa.kernel_launches <= 3 + 3 * (sa.max_level + 1)


def test_node_pool_overflow_falls_back_and_matches(gpu_ready):
    """A node pool too small for the batch's ray trees: the graph path notices after its single sync, the frame is
    redone with halved batches, and the image is the same."""
    scene = pt.Scene.example("glossy-reflection")
    kw = dict(samples=2, rng="hash", size=(200, 120))
    a, _ = parity.render_gpu(scene, **kw)
    b, sb = parity.render_gpu(scene, max_batch_paths=48000, node_pool_capacity=50000, **kw)
    assert np.array_equal(a.buffer, b.buffer)
    assert sb.retries >= 1


@pytest.mark.skipif(not has_reference_assets(), reason="reference textures not synced (tools/sync_assets.py)")
def test_texture_residency_cache(gpu_ready):
    """PtTexture.key keeps texels in HBM between scenes: the second upload of a scene copies records only, a
    records-only blob (cut at off_texels) is accepted while its textures are resident, and renders are identical."""
    import gc

    from portrayer_b200 import _ffi

    gc.collect()  # device scenes of earlier tests hold texture references until they are collected
    _ffi.gpu.pt_release_cached_memory()
    assert _ffi.gpu.pt_resident_texture_bytes() == 0
    scene = pt.Scene.example("normal-mapping")
    h = scene.header
    assert h.n_textures > 0 and h.n_texel_bytes > 1 << 20
    first = pt.DeviceScene(scene.blob)
    assert first.uploaded_bytes > h.n_texel_bytes * 0.9  # cold: the texels crossed PCIe
    resident = _ffi.gpu.pt_resident_texture_bytes()
    assert resident > 0
    second = pt.DeviceScene(scene.blob)
    assert second.uploaded_bytes < h.off_texels + 4096  # warm: records (+ the texture table) only
    assert _ffi.gpu.pt_resident_texture_bytes() == resident
    records_only = pt.DeviceScene(scene.blob[: h.off_texels])
    kw = dict(samples=1, rng="fixed", size=(227, 128))
    a, _ = parity.render_gpu(scene, dscene=first, **kw)
    b, _ = parity.render_gpu(scene, dscene=second, **kw)
    c, _ = parity.render_gpu(scene, dscene=records_only, **kw)
    ref = parity.render_oracle(scene, **kw)
    parity.assert_parity(parity.compare(a, ref, "resident textures"))
    assert np.array_equal(a.buffer, b.buffer) and np.array_equal(a.buffer, c.buffer)
    for ds in (first, second, records_only):
        ds.close()
    # unreferenced now: released on request; a records-only upload must then be refused, loudly
    _ffi.gpu.pt_release_cached_memory()
    assert _ffi.gpu.pt_resident_texture_bytes() == 0
    with pytest.raises(pt.PortrayerError) as err:
        pt.DeviceScene(scene.blob[: h.off_texels])
    assert "not resident" in str(err.value)


def test_frame_cache_rebinds_between_scenes(gpu_ready):
    """pt_render keeps frames (buffers + graph) per image geometry and re-binds them to the next scene."""
    a1, _ = parity.render_gpu(pt.Scene.example("nonhier"), samples=1, rng="fixed", size=(128, 128))
    b1, _ = parity.render_gpu(pt.Scene.example("primitives"), samples=1, rng="fixed", size=(128, 128))
    a2, _ = parity.render_gpu(pt.Scene.example("nonhier"), samples=1, rng="fixed", size=(128, 128))
    b2, _ = parity.render_gpu(pt.Scene.example("primitives"), samples=1, rng="fixed", size=(128, 128))
    assert np.array_equal(a1.buffer, a2.buffer) and np.array_equal(b1.buffer, b2.buffer)
    assert not np.array_equal(a1.buffer, b1.buffer)


# configs[4]: graphics-castle (13 KDMesh instances, dielectrics, glossy lake, ~6.9 k maze cubes) at reduced size
@pytest.mark.parametrize("kd_depth", [10, 16])
def test_graphics_castle(gpu_ready, kd_depth):
    scene = pt.Scene.example("graphics-castle", kd_depth=kd_depth)
    assert scene.header.n_instances > 6000 and scene.header.n_meshes >= 12
    kw = dict(samples=1, rng="hash", size=(384, 216))
    img, stats = parity.render_gpu(scene, **kw)
    ref = parity.render_oracle(scene, **kw)
    rep = parity.compare(img, ref, f"graphics-castle kd{kd_depth}")
    print(rep)
    parity.assert_parity(rep)
    assert (stats.rays_primary, stats.rays_shadow, stats.rays_reflect, stats.rays_refract) == \
           (ref.stats.rays_primary, ref.stats.rays_shadow, ref.stats.rays_reflect, ref.stats.rays_refract)


# configs[4] at full size holds one ray on which the reference's kd walk panics (see tests/test_oracle_kats.py).  With
# PT_RENDER_EXACT_WALK — the device walks every subtree the reference walks — the call fails with the reference's panic
# text and says where; with PT_RENDER_TOLERATE_KD_PLANE on top the frame is finished, the event stays visible in the
# stats, and every other sample is the oracle's.  The default walk skips subtrees the ray cannot hit anything in: it
# either reports the same event or never walks the subtree that holds it — and renders the same pixels either way.
@pytest.mark.skipif(not has_reference_assets(), reason="reference assets not synced")
def test_castle_kd_plane_panic_device(gpu_ready):
    from portrayer_b200 import _ffi
    from test_oracle_kats import CASTLE_PANIC_PIXEL

    scene = pt.Scene.example("graphics-castle")
    w, h = 3840, 2160
    x, y = CASTLE_PANIC_PIXEL
    kw = dict(samples=64, rng="hash", seed=1, size=(w, h), slice_=(x - 1, y, x, y))
    with pytest.raises(pt.PortrayerError) as err:
        parity.render_gpu(scene, flags=_ffi.PT_RENDER_EXACT_WALK, **kw)
    assert err.value.code == _ffi.PT_ERR_KD_PLANE_MISS
    assert "bug: ray should definitely hit infinite plane" in str(err.value) and f"pixel ({x}, {y}) sample 57" in str(err.value)
    img, stats = parity.render_gpu(scene, flags=_ffi.PT_RENDER_TOLERATE_KD_PLANE | _ffi.PT_RENDER_EXACT_WALK, **kw)
    assert stats.device_error_bits == _ffi.PT_DEVERR_KD_PLANE and stats.err_bit == _ffi.PT_DEVERR_KD_PLANE
    assert (stats.err_pixel % w, stats.err_pixel // w, stats.err_sample) == (x, y, 57)
    ref = parity.render_oracle(scene, samples=64, rng="hash", seed=1, size=(w, h), slice_=(x - 1, y, x - 1, y))
    assert ref.rc == 0 and np.array_equal(img.buffer[y, x - 1], ref.rgb[y, x - 1])
    # the default (pruning) walk: same pixels; the event, if the walk still meets it, is the same one
    fast, fstats = parity.render_gpu(scene, flags=_ffi.PT_RENDER_TOLERATE_KD_PLANE, **kw)
    assert np.array_equal(fast.buffer[y, x - 1:x + 1], img.buffer[y, x - 1:x + 1])
    assert np.array_equal(fast.hit_id[y, x - 1:x + 1], img.hit_id[y, x - 1:x + 1]) and np.array_equal(fast.hit_t[y, x - 1:x + 1], img.hit_t[y, x - 1:x + 1])
    assert fstats.device_error_bits in (0, _ffi.PT_DEVERR_KD_PLANE)
    print("default walk met the kd-plane event:", bool(fstats.device_error_bits))


# The default walk prunes subtrees (PT_RENDER_EXACT_WALK unset) and is what every other test of this file runs; here the
# exact walk — every subtree the reference walks — is held against the oracle and against the pruning walk: identical
# pictures, hit ids, hit parameters and ray counts.
@pytest.mark.parametrize("name,samples,rng,scale", [
    ("nonhier", 1, "fixed", 1), ("primitives", 2, "hash", 1), ("big-scene", 1, "hash", 4), ("glossy-reflection", 2, "hash", 2),
    ("graphics-castle", 1, "hash", 6), ("macho-cows", 1, "fixed", 2), ("kat-mesh-equivalence-kdmesh", 1, "fixed", 1)])
def test_exact_walk_and_pruning_walk_agree(gpu_ready, name, samples, rng, scale):
    from portrayer_b200 import _ffi

    scene = pt.Scene.example(name)
    size = (max(scene.width // scale, 16), max(scene.height // scale, 16))
    exact, st_e = parity.render_gpu(scene, samples=samples, rng=rng, size=size, flags=_ffi.PT_RENDER_EXACT_WALK)
    fast, st_f = parity.render_gpu(scene, samples=samples, rng=rng, size=size)
    ref = parity.render_oracle(scene, samples=samples, rng=rng, size=size)
    rep = parity.compare(exact, ref, name + " exact walk")
    parity.assert_parity(rep)
    assert rep["hit_t_bit_identical"], rep
    assert np.array_equal(exact.buffer, fast.buffer)
    assert np.array_equal(exact.hit_id, fast.hit_id) and np.array_equal(exact.hit_t, fast.hit_t)
    assert (st_e.rays_primary, st_e.rays_shadow, st_e.rays_reflect, st_e.rays_refract) == \
           (st_f.rays_primary, st_f.rays_shadow, st_f.rays_reflect, st_f.rays_refract)


# SURVEY 8 row a20, PT_RENDER_LINEAR_TLAS: the scene WITHOUT its k-d tree — `[FlatSceneNode]::ray_cast`
# (ray.rs:87-99 over flat_scene.rs:71-99) — on the device against the oracle's linear mode (bit-exact like the tree
# walk), and against the device's own tree walk: the tree is only an index, so hit ids agree wherever the walk's probe /
# EPSILON quirks lose nothing (they lose nothing on these scenes).
@pytest.mark.parametrize("name,samples,rng,scale", [
    ("nonhier", 1, "fixed", 1), ("primitives", 1, "fixed", 1), ("instance", 1, "fixed", 1), ("macho-cows", 1, "fixed", 2),
    ("glossy-reflection", 2, "hash", 2), ("water-glass", 2, "hash", 2), ("big-scene", 1, "fixed", 8)])
def test_linear_scene_cross_check(gpu_ready, name, samples, rng, scale):
    scene = pt.Scene.example(name)
    size = (scene.width // scale, scene.height // scale)
    lin = _report(name, samples=samples, rng=rng, size=size, flags=pt.PT_RENDER_LINEAR_TLAS)
    parity.assert_parity(lin)
    assert lin["hit_t_bit_identical"], lin
    img_lin, _ = parity.render_gpu(scene, samples=samples, rng=rng, size=size, flags=pt.PT_RENDER_LINEAR_TLAS)
    img_kd, _ = parity.render_gpu(scene, samples=samples, rng=rng, size=size)
    same = np.all(img_lin.hit_id == img_kd.hit_id, axis=2)
    assert same.mean() >= (0.995 if name == "big-scene" else 0.9999), (name, float(same.mean()))
    assert np.array_equal(img_lin.hit_t[same], img_kd.hit_t[same])
    assert (np.abs(img_lin.buffer.astype(int) - img_kd.buffer.astype(int)).max(axis=2) <= 1).mean() >= (0.99 if name == "big-scene" else 0.999)
