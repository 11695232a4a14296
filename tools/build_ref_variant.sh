#!/bin/bash
# Build the GPU library of a git revision (default HEAD) into build/variants/<name>/ so that tools/ab_bench.sh can A/B the
# working tree against it:   tools/build_ref_variant.sh [name] [rev]
set -e
cd "$(dirname "$0")/.."
name=${1:-base}; rev=${2:-HEAD}
wt=/tmp/pt_ref_worktree
git worktree remove --force $wt >/dev/null 2>&1 || true
git worktree add -f --detach $wt $rev >/dev/null 2>&1
mkdir -p build/variants/$name
[ -f build/scene_blob.o ] || make -s build/scene_blob.o build/tiles.o
(cd $wt && /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Iinclude \
  -shared portrayer_b200/csrc/*.cu $OLDPWD/build/scene_blob.o $OLDPWD/build/tiles.o -o $OLDPWD/build/variants/$name/libportrayer_gpu.so)
cp portrayer_b200/lib/libportrayer_host.so portrayer_b200/lib/libportrayer_blob.so portrayer_b200/lib/libportrayer_render.so build/variants/$name/
git worktree remove --force $wt >/dev/null 2>&1 || true
ls -la build/variants/$name
