#!/bin/bash
# Build A/B variants of the native GPU library with different -D tuning macros into build/variants/<name>/
# (each gets a copy of the host library so PORTRAYER_LIB_DIR can point at it):
#   tools/build_variants.sh name1 "-DX=1 -DY=2" name2 "-DX=3" ...
set -e
cd "$(dirname "$0")/.."
make -s gpu host >/dev/null
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=build/variants/$name
  mkdir -p $d
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude $flags \
    -shared portrayer_b200/csrc/*.cu build/scene_blob.o build/tiles.o -o $d/libportrayer_gpu.so &
  cp portrayer_b200/lib/libportrayer_host.so portrayer_b200/lib/libportrayer_blob.so portrayer_b200/lib/libportrayer_render.so $d/
done
wait
ls build/variants
