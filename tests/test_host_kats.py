"""The reference's own unit tests for the code that PREPARES the hot path's inputs, restated against the C++
host mirror (tests/cpp/test_host.cpp): src/bounding_box.rs:171-195, src/kdtree/leaf.rs:248-361, plus
flatten order, blob round trip and the rand-0.7 StdRng stream.  CPU only."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_known_answer_tests(native_libraries):
    exe = os.path.join(REPO, "tests", "cpp", "test_host")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", REPO, "hosttest"], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "all host checks passed" in res.stdout


def test_every_reference_example_used_by_the_configs_builds(native_libraries):
    """BASELINE.json configs name these scene programs; each must build (scene API -> flatten -> kd -> blob)."""
    import portrayer_b200 as pt
    from conftest import has_reference_assets

    names = set(pt.example_names())
    need = {"nonhier", "primitives", "big-scene", "glossy-reflection", "soft-shadows"}
    if has_reference_assets():
        need |= {"texture-mapping", "normal-mapping", "normal-mapping-left", "normal-mapping-right", "water-glass"}
    assert need <= names, need - names
    for name in sorted(need):
        sc = pt.Scene.example(name)
        assert sc.header.n_instances > 0 and sc.header.n_tlas_nodes > 0 and sc.width > 0


# "every examples/*.rs renders unchanged" -> every one of the 28 has a scene program here
REFERENCE_EXAMPLES = [
    "antialiasing", "big-scene", "cube-mapping", "entering-the-mirror-dimension", "fish", "four-shapes", "glossy-reflection",
    "graphics-castle", "graphics-poster", "graphics-temple", "hier", "instance", "macho-cows", "monkeys-making-monkeys", "nonhier",
    "nonhier2", "normal-mapping", "primitives-simple", "primitives", "robot-alarm-clock", "simple-cows", "simple", "single-triangle",
    "smooth-shading", "soft-shadows", "texture-mapping", "transmission-refraction", "water-glass",
]


def test_all_28_reference_examples_are_registered(native_libraries):
    import portrayer_b200 as pt

    assert len(REFERENCE_EXAMPLES) == 28
    missing = set(REFERENCE_EXAMPLES) - set(pt.example_names())
    assert not missing, missing


def test_last_two_examples_build_with_the_expected_instance_counts(native_libraries):
    """examples/monkeys-making-monkeys.rs: 3 + 3 + 5 + 5 + 6 + 1 + 3 geometry nodes; examples/graphics-temple.rs: 1 + 1 + 2 + 0
    (its unfinished maze floor emits nothing) + (16 columns x 5 + ceiling + 5 idol cubes) + (ceiling + 3 puppets) + 3."""
    import portrayer_b200 as pt
    from conftest import has_reference_assets
    import pytest

    if not has_reference_assets():
        pytest.skip("reference meshes not synced (tools/sync_assets.py)")
    sc = pt.Scene.example("monkeys-making-monkeys")
    assert (sc.width, sc.height, sc.header.n_instances, sc.header.n_lights) == (1920, 1080, 26, 2)
    sc = pt.Scene.example("graphics-temple")
    assert (sc.width, sc.height, sc.header.n_instances, sc.header.n_lights) == (533, 300, 97, 1)


def test_scene_programs_agree_with_the_reference_sources(native_libraries):
    """Where the reference checkout is present (this container; not the GPU box): every scene program's image size and
    light count against what its examples/*.rs says (`Image::new(.., W, H)`, `Light {` literals, comments skipped)."""
    import re

    import pytest

    import portrayer_b200 as pt
    from conftest import has_reference_assets

    examples = "/root/reference/examples"
    if not os.path.isdir(examples) or not has_reference_assets():
        pytest.skip("reference checkout not present")
    for name in REFERENCE_EXAMPLES:
        code = "\n".join(l for l in open(os.path.join(examples, name + ".rs")).read().splitlines() if not l.strip().startswith("//"))
        sizes = re.findall(r"Image::new\([^;]*?,\s*(\d+)\s*,\s*(\d+)\s*\)", code)
        assert sizes, name
        scene = pt.Scene.example(name)
        assert (int(sizes[-1][0]), int(sizes[-1][1])) == (scene.width, scene.height), (name, sizes, scene.width, scene.height)
        assert len(re.findall(r"\bLight\s*\{", code)) == scene.header.n_lights, name
