#!/bin/bash
# A/B the native-library variants under build/variants/* (built with different -D tuning macros) on the GPU box:
#   tools/ab_bench.sh [bench.py args...]   -> one summary line per variant
for d in portrayer_b200/lib build/variants/*; do
  [ -f "$d/libportrayer_gpu.so" ] || continue
  PORTRAYER_LIB_DIR=$(realpath $d) python bench.py --device-only "$@" 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    try: j=json.loads(line)
    except Exception: continue
    r=j['roofline']
    print('$d', 'Mrays/s %.1f ms/step %.3f'%(j['value'], j['ms_per_step']), 'kernels', {k: round(v,3) for k,v in r['kernel_ms_per_step'].items()}, 'frac %.3f'%r['frac'])
"
done
