#!/bin/bash
# The strong-scaling record of the default workload (graphics-castle 3840x2160 x 64, fixed work) on ONE box:
# bench.py at N = 8, 4, 2, 1 (peer-store exchange), the NCCL-gather exchange at N = 8, and the single-process device
# group (tests/cpp/test_group) over all GPUs.   tools/run_strong_scaling.sh ["8 4 2 1"]
NS=${1:-"8 4 2 1"}
mkdir -p gpurun_out
: > gpurun_out/strong_scaling.jsonl
line() { python - "$1" "$2" <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = [json.loads(l) for l in open(path) if l.startswith("{")][-1]
    e = d.get("e2e") or {}
    print(tag, round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 2), "ms/step (median", round(d["ms_per_step_median"], 2), "p95", round(d["ms_per_step_p95"], 2), ") e2e",
          round(e.get("value", 0), 1), "exchange ok", d["config"].get("exchange_verified_against_nccl_gather"), "broadcast ms", round(d["config"].get("scene_broadcast_ms", 0), 1))
    open("gpurun_out/strong_scaling.jsonl", "a").write(json.dumps(d) + "\n")
except Exception as ex:
    print(tag, "FAILED", ex)
    print(open(path.replace(".json", ".err")).read()[-1500:])
PY
}
for n in $NS; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --no-table > gpurun_out/ss_1.json 2> gpurun_out/ss_1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --no-table > gpurun_out/ss_$n.json 2> gpurun_out/ss_$n.err
  fi
  line "N=$n peer" gpurun_out/ss_$n.json
done
NMAX=$(echo $NS | awk '{print $1}')
if [ "$NMAX" != 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $NMAX --no-table --exchange nccl --device-only > gpurun_out/ss_${NMAX}_nccl.json 2> gpurun_out/ss_${NMAX}_nccl.err
  line "N=$NMAX nccl-gather" gpurun_out/ss_${NMAX}_nccl.json
fi
# one process, all GPUs, the unchanged Image::render of the host mirror
PORTRAYER_WRITE_DECODED=1 python -c "import portrayer_b200 as pt; pt.Scene.example('graphics-castle')" > /dev/null 2>&1
PORTRAYER_ASSETS=assets SAMPLES=16 tests/cpp/test_group $NMAX graphics-castle 16 | tee gpurun_out/ss_group.txt
