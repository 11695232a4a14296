#!/bin/bash
# one GPU round trip: parity tests, then device-only bench lines of the workloads named on the command line
#   tools/gpu_round.sh tag "castle-hd" "configs1" ...
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.txt 2>&1; echo "tests rc $?"; tail -5 gpurun_out/${TAG}_tests.txt
for w in "$@"; do
  n=$(echo $w | tr ' ' '_' | tr -d '-')
  timeout 600 python bench.py --device-only --workload $w > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$n.json"))
    print("$w", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],3), "ms", d["roofline"]["kernel_ms_per_step"], d["gpu_launches"]); print("   levels", d["roofline"]["kernel_ms_per_level"])
except Exception as e: print("$w", "FAILED", e); print(open("gpurun_out/${TAG}_$n.err").read()[-1500:])
PY
done
