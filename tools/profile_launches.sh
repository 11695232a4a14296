#!/bin/bash
# The ncu launch list of the default bench command (graphics-castle 3840x2160 x 64 on the kernel-by-kernel stream path:
# ncu cannot see kernel nodes of graphs with conditional nodes), first COUNT launches: duration, DRAM bytes, warp and
# thread instructions, FP64-pipe instructions per launch -> gpurun_out/<tag>_launches.csv
#   tools/profile_launches.sh [tag] [count] [bench args...]
TAG=${1:-launches}; COUNT=${2:-1500}; shift 2
mkdir -p gpurun_out
PT_DISABLE_GRAPHS=1 timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_pipe_fp64.sum \
  --clock-control none -c $COUNT --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --device-only --steps 2 --warmup 1 --streams 1 "$@" > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log | cut -c1-300
ls -la gpurun_out/${TAG}_launches.csv
