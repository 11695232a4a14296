"""Scene flattening on the device (pt_flatten) against the host mirror of FlatScene::from /
FlatSceneNode::new / FlatSceneNode::bounds (src/flat_scene.rs:18-46,63-69,101-108): same instances in the same
breadth-first order, bit-identical matrices, inverses and world bounds."""
import math

import numpy as np
import pytest

import portrayer_b200 as pt
from portrayer_b200 import flatten, kdbuild

from conftest import has_reference_assets

NEEDS_ASSETS = {"texture-mapping", "cube-mapping", "water-glass", "transmission-refraction", "robot-alarm-clock", "graphics-castle", "fish"}
SCENES = ["nonhier", "hier", "instance", "macho-cows", "simple-cows", "single-triangle", "primitives", "big-scene",
          "entering-the-mirror-dimension", "edge-empty", "edge-degenerate", "cube-mapping", "robot-alarm-clock", "graphics-castle"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def blob_instances(scene):
    h = scene.header
    inst = np.frombuffer(scene.blob, dtype=flatten.INSTANCE_DTYPE, count=h.n_instances, offset=h.off_instances)
    trans = np.frombuffer(scene.blob, dtype=np.float64, count=h.n_instances * 12, offset=h.off_instance_trans).reshape(-1, 12)
    return inst, trans


def test_hierarchy_export_shapes(native_libraries):
    """instancing: a shared SceneNode is exported once and referenced from several child lists (examples/instance.rs)"""
    scene = pt.Scene.example("instance")
    hier = flatten.hierarchy_of(scene)
    refs = np.bincount(hier.children, minlength=len(hier.nodes))
    assert refs.max() == 6  # the arc is instanced six times
    assert scene.header.n_instances == 20 and int((hier.nodes["geometry"] != 0xFFFFFFFF).sum()) == 5
    with pytest.raises(ValueError):
        flatten.hierarchy_of(pt.Scene.example("kat-edge-case"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_device_flatten_equals_host_flatten(gpu_ready, name):
    if name in NEEDS_ASSETS and not has_reference_assets():
        pytest.skip("reference assets not synced")
    scene = pt.Scene.example(name)
    flat = flatten.FlatScene.build(flatten.hierarchy_of(scene))
    inst, trans, bounds = flat.download()
    ref_inst, ref_trans = blob_instances(scene)
    assert flat.instance_count == scene.header.n_instances
    for field in ("prim", "mesh", "material"):
        assert np.array_equal(inst[field], ref_inst[field]), field
    assert np.array_equal(bits(trans), bits(ref_trans)), "total_trans = parent * node (flat_scene.rs:30)"
    assert np.array_equal(bits(inst["invtrans"]), bits(ref_inst["invtrans"])), "trans.inverted() (flat_scene.rs:103)"
    ref_bounds = scene.item_bounds()
    assert np.array_equal(bounds, ref_bounds), "trans * primitive.bounds() (flat_scene.rs:63-69)"


# graph -> flatten -> tree build -> render without the instances or the tree ever leaving the device
@pytest.mark.gpu
def test_flatten_build_render_chain(gpu_ready):
    import parity

    n, depth = 30_000, 14
    scene = pt.Scene.synthetic_instances(n, kd_depth=depth)
    kw = dict(samples=1, rng="hash", size=(330, 170))
    a, _ = parity.render_gpu(scene, **kw)

    ds = pt.DeviceScene(scene.blob)  # meshes / materials / lights come from the blob; instances and tree are replaced
    flat = flatten.FlatScene.build(flatten.hierarchy_of(scene))
    tree = kdbuild.KdTree.build_device(flat.bounds_device_ptr, flat.instance_count, kdbuild.config(depth))
    ds.set_instances(flat, tree)
    ms_f, _ = flat.build_stats()
    ms_t, _ = tree.build_stats()
    print(f"{n} instances: flatten {ms_f:.3f} ms + tree {ms_t:.3f} ms on the device; host prepare {scene.prepare_seconds * 1e3:.0f} ms")
    flat.close()
    tree.close()
    b, _ = parity.render_gpu(scene, dscene=ds, **kw)
    assert np.array_equal(a.buffer, b.buffer) and np.array_equal(a.hit_id, b.hit_id) and np.array_equal(a.hit_t, b.hit_t)
