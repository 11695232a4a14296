#!/bin/bash
# The profile passes of the default bench command kept under profiles/ (stream path: ncu cannot see kernels inside
# conditional graph nodes):  tools/profile_bench.sh <tag>
#   1. launch list + DRAM bytes per launch of `bench.py --device-only --steps 2 --warmup 1`
#   2. --set full of the first extend / shadow / shade launches of every frame of the step
TAG=${1:-r1}
mkdir -p gpurun_out
PT_DISABLE_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 \
  --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --device-only --steps 2 --warmup 1 --streams 1 > gpurun_out/${TAG}_launches.log 2>&1
PT_DISABLE_GRAPHS=1 timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k regex:'(extend|shadow)_kernel<\(bool\)0, \(bool\)0>|shade_kernel' -c 9 -f -o gpurun_out/${TAG}_configs1 \
  python bench.py --device-only --steps 1 --warmup 1 --streams 1 > gpurun_out/${TAG}_configs1.log 2>&1
ls -la gpurun_out/${TAG}_*
