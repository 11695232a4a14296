#!/bin/bash
# A/B every library variant (tools/build_variants.sh) over several workloads on the GPU box:
#   tools/ab_workloads.sh "configs1" "castle-hd" "configs1 --samples 4" ...
for w in "$@"; do
  echo "== $w"
  tools/ab_bench.sh --workload $w
done
