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
