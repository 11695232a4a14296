"""bench.py's reference arm (`--impl reference`: the CPU port of the reference's render loop on the host cores) runs
without a GPU, so its side of the JSON contract is checked here: one line, the keys the driver reads, the same metric /
unit / workload naming as the GPU arm, and under torchrun only rank 0 prints.  CPU only."""
import json
import os
import subprocess
import sys

import pytest

from conftest import has_reference_assets

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "cpu_baseline", "e2e"]


def _run(args, env=None):
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                         env={**os.environ, **(env or {})})
    assert res.returncode == 0, res.stderr[-2000:]
    return [l for l in res.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_the_contract_line():
    workload = "configs1" if has_reference_assets() else "nonhier"
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", workload])
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    """under torchrun the reference arm runs on rank 0 alone; the other ranks exit 0 without work or output"""
    lines = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "nonhier"],
                 env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == []


def test_gpu_arm_refuses_to_run_without_a_gpu():
    """no CPU fallback: without a CUDA device the GPU arm fails loudly instead of timing the oracle"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stderr + res.stdout)
