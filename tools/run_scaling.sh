#!/bin/bash
# bench.py at N GPUs of one box, both exchange modes:  tools/run_scaling.sh "2 4 8" [extra bench args]
NS=${1:-2}; shift
mkdir -p gpurun_out
for n in $NS; do
  for ex in peer nccl; do
    if [ "$n" = 1 ]; then
      [ $ex = nccl ] && continue
      python bench.py --gpus 1 "$@" > gpurun_out/scale_${n}_$ex.json 2> gpurun_out/scale_${n}_$ex.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n --exchange $ex "$@" > gpurun_out/scale_${n}_$ex.json 2> gpurun_out/scale_${n}_$ex.err
    fi
    python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/scale_${n}_$ex.json") if l.startswith("{")][-1]
    print("N=$n $ex", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms/step e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, d["config"].get("exchange_verified_against_nccl_gather"))
except Exception as e:
    print("N=$n $ex FAILED", e); print(open("gpurun_out/scale_${n}_$ex.err").read()[-1500:])
PY
  done
done
