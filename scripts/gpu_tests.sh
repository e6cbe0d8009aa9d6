set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 1200 python -m pytest tests/test_search_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --W 8 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'recall',d['config']['recall_at_10'],'roofline',d['roofline'])"
bash scripts/profile.sh r01c --W 8
