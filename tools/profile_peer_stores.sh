#!/bin/bash
# Evidence for the fused exchange (resolve kernels storing RGB8 tiles into the primary's image over NVLink), on 2 GPUs:
#  (a) ncu on the SINGLE-PROCESS device group (tests/cpp/test_group, stream path): NVLink bytes sent / received by the
#      resolve kernel launches of both devices;
#  (b) nvidia-smi NVLink data counters around a 2-rank bench run of castle-hd (one process per GPU, CUDA IPC mapping).
mkdir -p gpurun_out
PORTRAYER_WRITE_DECODED=1 python -c "import portrayer_b200 as pt; pt.Scene.example('graphics-castle')" > /dev/null 2>&1
PT_DISABLE_GRAPHS=1 PORTRAYER_ASSETS=assets timeout 600 ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,dram__bytes_write.sum \
  --clock-control none -k regex:resolve_kernel -c 12 --csv --log-file gpurun_out/peer_resolve.csv tests/cpp/test_group 2 graphics-castle 4 > gpurun_out/peer_resolve.log 2>&1
tail -2 gpurun_out/peer_resolve.log
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/peer_resolve.csv")) if len(r) > 10]
if rows:
    h = rows[0]; dev = h.index("Device") if "Device" in h else None
    for r in rows[1:]:
        print(r[h.index("ID")], (r[dev] if dev is not None else ""), r[h.index("Kernel Name")][:40], r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
nvidia-smi nvlink -gt d > gpurun_out/nvlink_before.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload castle-hd --device-only --steps 20 --warmup 3 > gpurun_out/peer_bench_n2.json 2> gpurun_out/peer_bench_n2.err
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after.txt 2>&1
python - <<'PY'
import re
def parse(p):
    out = {}; gpu = None
    for line in open(p):
        m = re.match(r"GPU (\d+):", line)
        if m: gpu = int(m.group(1)); out[gpu] = [0, 0]
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None: out[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
    return out
a, b = parse("gpurun_out/nvlink_before.txt"), parse("gpurun_out/nvlink_after.txt")
for g in sorted(b):
    print("GPU", g, "NVLink Tx +%.1f MiB  Rx +%.1f MiB" % ((b[g][0] - a.get(g, [0, 0])[0]) / 1024, (b[g][1] - a.get(g, [0, 0])[1]) / 1024))
PY
tail -c 300 gpurun_out/peer_bench_n2.json
