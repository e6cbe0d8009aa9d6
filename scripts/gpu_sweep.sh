# usage: bash scripts/gpu_sweep.sh "<bench args 1>" "<bench args 2>" ...   (each: one short bench run, compact summary)
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
for args in "$@"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gt-queries 500 $args 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['config']
    print('ARGS $args | value',d['value'],'e2e',d['e2e']['value'],'recall',c['recall_at_10'],'frac',r['frac'],'kms',r['kernel_ms_per_launch'],'hops',c['mean_hops'],'vis',c['mean_visited'],'B/q',r['algorithmic_bytes_per_query'])
except Exception as e:
    print('ARGS $args | FAILED', e)
"
done
