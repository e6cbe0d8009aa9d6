# usage: gpu_variants.sh "<bench args>" name1 name2 ...   (default library first, then diskrag_b200/variants/lib_<name>.so)
args="$1"; shift
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
run() {
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gt-queries 300 $args 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['config']
    print('VARIANT $1 | value',d['value'],'recall',c['recall_at_10'],'frac',r['frac'],'kms',r['kernel_ms_per_launch'])
except Exception as e:
    print('VARIANT $1 | FAILED', e)
"
}
run default
for v in "$@"; do DISKRAG_B200_LIB=$PWD/diskrag_b200/variants/lib_$v.so run $v; done
