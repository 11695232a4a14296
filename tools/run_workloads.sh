mkdir -p gpurun_out
for w in "configs1" "configs1 --samples 4" "configs1 --streams 1" "nonhier" "big-scene" "secondary" "castle-hd" "synthetic-instances-1e5"; do
  n=$(echo $w | tr ' ' '_' | tr -d '-')
  timeout 300 python bench.py --device-only --workload $w > gpurun_out/base_$n.json 2> gpurun_out/base_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/base_$n.json"))
    print("$w", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],3), "ms", d["roofline"]["kernel_ms_per_step"], d["gpu_launches"])
except Exception as e: print("$w", "FAILED", e)
PY
done
