set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'recall',d['config']['recall_at_10'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms_per_launch'],'kshare',d['roofline']['kernel_share_of_step'])"
