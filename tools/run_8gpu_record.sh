#!/bin/bash
# the multi-GPU record kept under profiles/: configs[1] at N=8 (fused peer exchange, full line with e2e; NCCL gather for
# comparison) and configs[4] proper (graphics-castle 3840x2160, SAMPLES=64) at N=8
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n "$@" > gpurun_out/rec_$tag.json 2> gpurun_out/rec_$tag.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/rec_$tag.json") if l.startswith("{")][-1]
    print("$tag", round(d["value"],1), "Mrays/s", round(d["ms_per_step"],4), "ms/step e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, d["config"].get("exchange_verified_against_nccl_gather"))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/rec_$tag.err").read()[-1500:])
PY
}
run 8 n8_peer
run 8 n8_nccl --exchange nccl --device-only
run 8 n8_castle --workload castle --steps 2 --warmup 1 --device-only
run 4 n4_peer --device-only
