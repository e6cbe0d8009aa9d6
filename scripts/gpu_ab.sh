# A/B of kernel variants in one gpurun call.  usage: gpu_ab.sh "<common bench args>" name[:extra bench args] ...
#   name = default -> the regular library; anything else -> diskrag_b200/variants/lib_<name>.so (scripts/build_variants.py)
common="$1"; shift
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
for spec in "$@"; do
  name="${spec%%:*}"; extra=""; [ "$name" != "$spec" ] && extra="${spec#*:}"
  libenv=""; [ "$name" != "default" ] && libenv="$PWD/diskrag_b200/variants/lib_$name.so"
  DISKRAG_B200_LIB=$libenv timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --gt-queries 300 --no-points $common $extra 2> gpurun_out/ab_$name.err | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['config']
    print('VARIANT $name | value',d['value'],'recall',c['recall_at_10'],'frac',r['frac'],'kms',r['kernel_ms_per_launch'],'hops',c['mean_hops'],'vis',c['mean_visited'])
except Exception as e:
    print('VARIANT $name | FAILED', e)
"
done
