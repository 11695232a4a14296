"""The oracle against every known-answer test the reference holds for the hot path.
CPU only: these pin the checker before it is trusted to check the CUDA path."""
import math

import numpy as np
import pytest

import portrayer_b200 as pt
from conftest import has_reference_assets
from oracle import binding as oracle


def approx(a, b, eps=1e-6):  # assert_approx_eq!'s default
    return abs(a - b) < eps


# src/math.rs:159-172 solve_quadratic_equations
def test_solve_quadratic_equations():
    s = oracle.solve_quadratic(2.0, 8.0, 3.0)  # discriminant > 0, ascending
    assert len(s) == 2 and approx(s[0], -2.0 - math.sqrt(5.0 / 2.0)) and approx(s[1], math.sqrt(5.0 / 2.0) - 2.0)
    s = oracle.solve_quadratic(4.0, -4.0, 1.0)  # discriminant == 0
    assert len(s) == 1 and approx(s[0], 0.5)
    assert oracle.solve_quadratic(3.0, 4.0, 2.0) == []  # discriminant < 0


# src/math.rs:174-179 solution_order
def test_solution_order():
    s = oracle.solve_quadratic(-2.0, 8.0, 3.0)
    assert len(s) == 2 and approx(s[0], 2.0 - math.sqrt(11.0 / 2.0)) and approx(s[1], 2.0 + math.sqrt(11.0 / 2.0))
    assert s[0] < s[1]


def test_quadratic_linear_fallback():
    assert oracle.solve_quadratic(0.0, 2.0, -4.0) == [2.0]  # a == 0 falls back to the linear solver
    assert oracle.solve_quadratic(0.0, 0.0, 1.0) == []


# src/kdtree/node.rs:219-293 / :295-351 — B (instance 0, red) must win over C (instance 1, blue)
@pytest.mark.parametrize("name,origin,direction", [
    ("kat-edge-case", (0.0, 0.5, 0.9), (0.0, 0.0, -1.0)),
    ("kat-edge-case-flipped", (0.0, 0.5, -0.9), (0.0, 0.0, 1.0)),
])
def test_ray_cast_edge_case(name, origin, direction):
    scene = pt.Scene.example(name)
    rc, color, hit_id, hit_t, _ = oracle.trace_rays(scene.blob, np.array([origin]), np.array([direction]), threads=1)
    assert rc == 0
    assert hit_id[0, 0] == 0, "the nearer plane B must be returned, not C which straddles the split"
    assert tuple(color[0]) == (1.0, 0.0, 0.0)  # mat_b = Rgb::red()
    assert math.isfinite(hit_t[0]) and hit_t[0] > 0


def _mesh_equivalence_rays(scene, n):
    # kdmesh.rs:155-160: x = width * i / n, y = height * i / n
    i = np.arange(n, dtype=np.float64)
    xy = np.stack([533.0 * i / n, 300.0 * i / n], axis=1)
    return oracle.camera_rays(scene.camera(533, 300), xy)


# src/kdtree/kdmesh.rs:99-166 mesh_equivalence: Mesh and KDMesh give the SAME colour, bit for bit, for 100 000 rays
def test_mesh_equivalence():
    mesh = pt.Scene.example("kat-mesh-equivalence-mesh")
    kdmesh = pt.Scene.example("kat-mesh-equivalence-kdmesh")
    assert mesh.header.n_triangles == 1804  # castle.obj
    n = 100000
    origins, dirs = _mesh_equivalence_rays(mesh, n)
    rc_a, color_a, id_a, t_a, st_a = oracle.trace_rays(mesh.blob, origins, dirs)
    rc_b, color_b, id_b, t_b, st_b = oracle.trace_rays(kdmesh.blob, origins, dirs)
    assert rc_a == 0 and rc_b == 0
    assert np.array_equal(color_a, color_b), "pixels were not the same"
    assert np.array_equal(t_a, t_b)
    assert np.array_equal(id_a, id_b)  # same triangle index: both are MeshData.triangles order
    assert (id_a[:, 0] != 0xFFFFFFFF).sum() > n // 10  # the sweep really crosses the castle
    # the kd walk tests far fewer triangles than the linear scan
    assert st_b.triangle_tests * 5 < st_a.triangle_tests


# The reference's kd walk panics when rounding puts a split-plane crossing outside the ray's range
# (.expect("bug: ray should definitely hit infinite plane"), kdtree/node.rs:147,178; README.md:247-248 calls the
# kd-tree "occasionally buggy").  The 4K x 64-sample graphics-castle frame of BASELINE.json configs[4] holds one such
# ray under the hashed jitter of seed 1 (found by the device, which reports where the panic fires): the restatement
# panics on exactly that pixel and not on its neighbour.  tests/test_gpu_parity.py checks the device against this.
CASTLE_PANIC_PIXEL = (1784, 1742)


@pytest.mark.skipif(not has_reference_assets(), reason="reference assets not synced (tools/sync_assets.py)")
def test_castle_kd_plane_panic_is_reproduced(native_libraries):
    from portrayer_b200 import _ffi
    from portrayer_b200.render import _background_arg, make_params

    scene = pt.Scene.example("graphics-castle")
    w, h = 3840, 2160
    bg, bg_mode = _background_arg(scene, w, h)
    x, y = CASTLE_PANIC_PIXEL
    for px, expect in (((x, y), _ffi.PT_ERR_KD_PLANE_MISS), ((x - 1, y), 0)):
        p = make_params(w, h, 64, "hash", 1, slice_=(px[0], px[1], px[0], px[1]), bg_mode=bg_mode)
        res = oracle.render(scene.blob, scene.camera(w, h), p, bg, threads=1)
        assert res.rc == expect


def test_mesh_fold_ties_first_listed_wins():
    """ray.rs:50-63 / mesh.rs:157-167: among triangles that return exactly the same t the first listed wins.  Pins the
    oracle's fold against a brute-force answer on a mesh built to tie (the device's Morton-order fold is then held to
    the oracle in test_gpu_parity.py)."""
    import parity

    scene = pt.Scene.example("edge-mesh-ties")
    origins, dirs, expect_sub, expect_t = parity.mesh_tie_rays()
    rc, _color, hit_id, hit_t, _ = oracle.trace_rays(scene.blob, origins, dirs)
    assert rc == 0
    hit = expect_sub >= 0
    assert hit.sum() > 1000 and (expect_t == 4.0).sum() > 300
    assert np.array_equal(hit_id[hit, 0], np.zeros(hit.sum(), np.uint32))
    assert np.array_equal(hit_id[hit, 1].astype(np.int64), expect_sub[hit])
    assert np.array_equal(hit_t[hit], expect_t[hit])
    assert np.all(hit_id[~hit, 0] == 0xFFFFFFFF)


# the oracle's linear mode (`[FlatSceneNode]::ray_cast`, ray.rs:87-99 — no k-d tree) is the semantic
# cross-check of its tree walk: same nearest hit on scenes where the walk's quirks lose nothing.
@pytest.mark.parametrize("name,scale", [("nonhier", 2), ("primitives", 4), ("instance", 2), ("hier", 2)])
def test_tree_walk_equals_linear_scan(name, scale):
    import parity
    scene = pt.Scene.example(name)
    size = (scene.width // scale, scene.height // scale)
    kd = parity.render_oracle(scene, samples=1, rng="fixed", size=size)
    lin = parity.render_oracle(scene, samples=1, rng="fixed", size=size, flags=pt.PT_RENDER_LINEAR_TLAS)
    assert kd.rc == 0 and lin.rc == 0
    assert np.array_equal(kd.hit_id, lin.hit_id)
    assert np.array_equal(kd.hit_t, lin.hit_t)
    assert np.array_equal(kd.rgb, lin.rgb)
    assert lin.stats.kd_splits < kd.stats.kd_splits  # only KDMesh trees are walked in linear mode
