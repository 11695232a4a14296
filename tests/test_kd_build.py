"""k-d tree build on the device (pt_kd_build) against the host mirror of
KDLeaf::partitioned (src/kdtree/leaf.rs:89-231) and against the reference's own exact-tree tests
(src/kdtree/leaf.rs:248-361)."""
import math

import numpy as np
import pytest

import portrayer_b200 as pt
from portrayer_b200 import kdbuild

from conftest import has_reference_assets


def plane_bounds(xs):
    """leaf.rs:263-267: a unit Plane rotated 90 degrees about z and translated to x: bounds x +- ~0, y, z +-0.5"""
    c, s = math.cos(math.radians(90.0)), math.sin(math.radians(90.0))
    out = []
    for x in xs:
        # corners (+-0.5, 0, +-0.5) through Rz(90) then T(x): x' = c*px - s*py + x, y' = s*px + c*py
        px = np.array([-0.5, 0.5])
        xs_ = c * px + x
        ys_ = s * px
        out.append([xs_.min(), ys_.min(), -0.5, xs_.max(), ys_.max(), 0.5])
    return np.array(out, dtype=np.float64)


def leaves_of(nodes, items):
    """(split, axis, front, back) of the root and the member lists of its two children (which must be leaves)"""
    root = nodes[0]
    assert root["a"] & 3 != 3
    front, back = nodes[root["a"] >> 2], nodes[root["b"]]
    assert front["a"] & 3 == 3 and back["a"] & 3 == 3
    fl = items[front["a"] >> 2:(front["a"] >> 2) + front["b"]]
    bl = items[back["a"] >> 2:(back["a"] >> 2) + back["b"]]
    return float(root["split"]), int(root["a"] & 3), list(fl), list(bl)


KATS = [
    # leaf.rs:248-300 single_axis_center_partition: A B | C D E around x = 0
    ([-8.0, -5.0, 3.0, 5.0, 8.0], 3, 0.0, [0, 1], [2, 3, 4]),
    # leaf.rs:302-360 single_axis_uneven_partition: the bisection ends on x = 4.0
    ([-8.0, 0.0, 3.0, 5.0, 8.0], 2, 4.0, [0, 1, 2], [3, 4]),
]


@pytest.mark.parametrize("xs,merit,plane,back,front", KATS)
def test_reference_partition_kats_host(native_libraries, xs, merit, plane, back, front):
    t = kdbuild.host_build(plane_bounds(xs), kdbuild.config(5, 3, merit, 10))
    split, axis, fl, bl = leaves_of(t.nodes, t.items)
    assert axis == 0 and round(split, 9) == plane and fl == front and bl == back


def assert_same_tree(dev: "kdbuild.KdTree", ref: "kdbuild.HostTree"):
    nodes, items = dev.download()
    assert dev.depth == ref.depth
    assert nodes.shape == ref.nodes.shape and items.shape == ref.items.shape
    assert np.array_equal(nodes["a"], ref.nodes["a"]) and np.array_equal(nodes["b"], ref.nodes["b"])
    assert np.array_equal(items, ref.items)
    # planes: equal as numbers everywhere (bit-identical except possibly the sign of a zero, see kd_build.cu)
    assert np.array_equal(nodes["split"], ref.nodes["split"])
    nz = ref.nodes["split"] != 0.0
    assert np.array_equal(nodes["split"][nz].view(np.uint64), ref.nodes["split"][nz].view(np.uint64))
    _, extent = dev.root_bounds()
    assert extent == ref.extent


@pytest.mark.gpu
@pytest.mark.parametrize("xs,merit,plane,back,front", KATS)
def test_reference_partition_kats_device(gpu_ready, xs, merit, plane, back, front):
    cfg = kdbuild.config(5, 3, merit, 10)
    t = kdbuild.KdTree.build(plane_bounds(xs), cfg)
    nodes, items = t.download()
    split, axis, fl, bl = leaves_of(nodes, items)
    assert axis == 0 and round(split, 9) == plane and fl == front and bl == back
    assert_same_tree(t, kdbuild.host_build(plane_bounds(xs), cfg))


@pytest.mark.gpu
def test_degenerate_inputs_device(gpu_ready):
    for bounds, cfg in [
        (np.zeros((0, 6)), kdbuild.config()),                                    # empty scene: one empty leaf
        (np.array([[0, 0, 0, 1, 1, 1.0]]), kdbuild.config()),                     # fewer items than target_max_nodes
        (np.tile(np.array([[0, 0, 0, 1, 1, 1.0]]), (50, 1)), kdbuild.config(6)),  # identical boxes: every split shares everything
        (random_boxes(64, 1), kdbuild.config(0)),                                # max_depth 0: one leaf with everything
    ]:
        t = kdbuild.KdTree.build(bounds, cfg)
        assert_same_tree(t, kdbuild.host_build(bounds, cfg))


def random_boxes(n, seed, spread=100.0, size=3.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-spread, spread, (n, 3))
    h = rng.uniform(0.0, size, (n, 3))
    return np.concatenate([c - h, c + h], axis=1)


@pytest.mark.gpu
@pytest.mark.parametrize("n,depth,seed", [(10, 10, 1), (1000, 10, 2), (5000, 14, 3), (200_000, 17, 4)])
def test_random_boxes_device(gpu_ready, n, depth, seed):
    bounds = random_boxes(n, seed)
    cfg = kdbuild.config(depth)
    t = kdbuild.KdTree.build(bounds, cfg)
    ref = kdbuild.host_build(bounds, cfg)
    assert_same_tree(t, ref)
    ms, launches = t.build_stats()
    print(f"n={n} depth={depth}: device {ms:.3f} ms ({launches} launches), host {ref.seconds * 1e3:.1f} ms, {t.node_count} nodes, {t.item_count} leaf members")


@pytest.mark.gpu
@pytest.mark.parametrize("merit,tries,tmn", [(0, 10, 3), (3, 1, 3), (3, 0, 3), (10, 10, 1), (3, 30, 8)])
def test_partition_config_variants_device(gpu_ready, merit, tries, tmn):
    bounds = random_boxes(3000, 7, spread=20.0, size=4.0)  # heavy overlap: many shared members
    cfg = kdbuild.config(9, tmn, merit, tries)
    assert_same_tree(kdbuild.KdTree.build(bounds, cfg), kdbuild.host_build(bounds, cfg))


# the scene trees of the example scenes (KDTreeScene::from, kdscene.rs:19-44): the device tree equals the tree the host
# packed into the blob, record for record
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["nonhier", "primitives", "big-scene", "glossy-reflection"])
def test_scene_tree_equals_blob_tree(gpu_ready, name):
    scene = pt.Scene.example(name)
    bounds = scene.item_bounds()
    assert bounds.shape[0] == scene.header.n_instances
    t = kdbuild.KdTree.build(bounds, kdbuild.config(10))
    nodes, items = t.download()
    blob_nodes = np.frombuffer(scene.blob, dtype=kdbuild.KD_NODE_DTYPE, count=scene.header.n_tlas_nodes, offset=scene.header.off_tlas_nodes)
    blob_items = scene.section("tlas_items", np.uint32, 1)
    assert np.array_equal(nodes["a"], blob_nodes["a"]) and np.array_equal(nodes["b"], blob_nodes["b"])
    assert np.array_equal(nodes["split"], blob_nodes["split"]) and np.array_equal(items, blob_items)
    assert t.depth == scene.header.tlas_depth and t.root_bounds()[1] == scene.header.tlas_extent


# KDMesh::new (kdmesh.rs:37-58): the tree over the triangles of a mesh; graphics-castle carries 13 of them
@pytest.mark.gpu
@pytest.mark.skipif(not has_reference_assets(), reason="reference assets not synced")
def test_kdmesh_trees_equal_blob_trees(gpu_ready):
    import ctypes as C

    from portrayer_b200._ffi import PtBlobHeader  # noqa: F401

    scene = pt.Scene.example("graphics-castle")
    h = scene.header
    tri = scene.section("tri_pos", np.float64, 9)
    meshes = np.frombuffer(scene.blob, dtype=np.uint8, count=h.n_meshes * 160, offset=h.off_meshes).reshape(h.n_meshes, 160)
    blas_nodes = np.frombuffer(scene.blob, dtype=kdbuild.KD_NODE_DTYPE, count=h.n_blas_nodes, offset=h.off_blas_nodes)
    blas_items = scene.section("blas_items", np.uint32, 1)
    checked = 0
    for m in range(h.n_meshes):
        rec = meshes[m]
        extent = rec[96:104].view(np.float64)[0]
        kind, _flags, tri_first, tri_count, node_first, node_count, item_first, item_count = rec[104:136].view(np.uint32)
        if kind != 1:  # PT_MESH_KD
            continue
        bounds = kdbuild.triangle_bounds(tri[tri_first:tri_first + tri_count])
        t = kdbuild.KdTree.build(bounds, kdbuild.config(10))
        nodes, items = t.download()
        ref_nodes = blas_nodes[node_first:node_first + node_count]
        assert len(nodes) == node_count and len(items) == item_count
        assert np.array_equal(nodes["a"], ref_nodes["a"]) and np.array_equal(nodes["b"], ref_nodes["b"])
        assert np.array_equal(nodes["split"], ref_nodes["split"])
        assert np.array_equal(items, blas_items[item_first:item_first + item_count])
        assert t.root_bounds()[1] == extent
        checked += 1
    assert checked >= 10


# a device-built scene tree spliced into an uploaded scene renders the same image, bit for bit
@pytest.mark.gpu
def test_render_with_device_built_tree(gpu_ready):
    import parity

    scene = pt.Scene.synthetic_instances(20_000, kd_depth=13)
    kw = dict(samples=1, rng="hash", size=(330, 170))
    a, _ = parity.render_gpu(scene, **kw)
    # same scene, but uploaded with a linear (single-leaf) scene tree, then given the tree built on the device
    flat = pt.Scene.synthetic_instances(20_000, kd_depth=0)
    ds = pt.DeviceScene(flat.blob)
    tree = kdbuild.KdTree.build(flat.item_bounds(), kdbuild.config(13))
    ds.set_tlas(tree)
    tree.close()
    b, _ = parity.render_gpu(flat, dscene=ds, **kw)
    assert np.array_equal(a.buffer, b.buffer) and np.array_equal(a.hit_id, b.hit_id) and np.array_equal(a.hit_t, b.hit_t)
