# per-kernel device times of one bench step (ncu launch list): usage gpu_launchtimes.sh "<bench args>"
for args in "$@"; do
  echo "ARGS $args"
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --queries 50000 --steps 1 --warmup 1 --gt-queries 100 --no-cpu-baseline --cuda-profile $args 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
for r in rows: print('  %-70s %10.3f ms' % (r[4][:70], float(r[-1])/1e6))
"
done
