set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 1200 python -m pytest tests/test_search_gpu.py -q -m gpu -x 2>&1 | tail -3
for W in 5 6 8; do
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --W $W 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('W',$W,'value',d['value'],'e2e',d['e2e']['value'],'recall',d['config']['recall_at_10'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms_per_launch'],'kshare',d['roofline']['kernel_share_of_step'])"
done
