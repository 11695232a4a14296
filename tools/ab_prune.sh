for w in castle-hd configs1 secondary big-scene "castle --samples 16" synthetic-instances-1e5 synthetic-triangles-1e6; do
  echo "== $w"
  for mode in 0 1; do
    PT_EXACT_WALK=$mode python bench.py --device-only --workload $w 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    try: j=json.loads(line)
    except Exception: continue
    r=j['roofline']
    print('exact_walk=$mode', 'Mrays/s %.1f ms/step %.3f'%(j['value'], j['ms_per_step']), 'kernels', {k: round(v,3) for k,v in r['kernel_ms_per_step'].items()})
"
  done
done
