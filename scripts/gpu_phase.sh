# per-phase cycle accounting of the fast search kernel (needs a DR_PHASE_TIMING=1 build): usage gpu_phase.sh "<bench args>" ...
for args in "$@"; do
  echo "ARGS $args"
  timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --gt-queries 200 $args 2>&1 | grep -E "phase|value" | tail -2 | python -c "
import sys, json, re
for l in sys.stdin:
    if l.startswith('[phase]'):
        print('  ' + ' | '.join(re.findall(r'([A-Za-z0-9+/ ]+ [0-9.]+% \([0-9]+ cyc/query\))', l)))
    else:
        try:
            d = json.loads(l); print('  value', d['value'], 'recall', d['config']['recall_at_10'], 'frac', d['roofline']['frac'])
        except Exception as e: print('  ?', l[:200])
"
done
