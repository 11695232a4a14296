#!/bin/bash
# ncu --set full of consecutive extend / shadow launches of graphics-castle (1080p x 1 spp) on the stream path
# (ncu cannot profile kernels inside conditional graph nodes); report comes back in gpurun_out/.
#   tools/profile_castle.sh [workload] [tag] [count] [skip]
# With the stream path the first matching launches are warm-up step 1: extend L0, shadow L0, extend L1, shadow L1, ...
W=${1:-castle-hd}; TAG=${2:-castle}; N=${3:-12}; SKIP=${4:-0}
mkdir -p gpurun_out
PT_DISABLE_GRAPHS=1 timeout 1200 ncu --set full --import-source on --clock-control none \
  --kernel-name-base demangled -k regex:'(extend|shadow)_kernel<\(bool\)0, \(bool\)0, \(bool\)1>' -s $SKIP -c $N -f -o gpurun_out/prof_$TAG \
  python bench.py --device-only --workload $W --samples ${PROFILE_SAMPLES:-1} --steps 1 --warmup 1 --streams 1 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
