set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 900 python -m pytest tests/test_build_gpu.py tests/test_build_index_gpu.py tests/test_dataset_benchmark.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-points --gt-queries 1000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'recall',d['config']['recall_at_10'],'hops',d['config']['mean_hops'],'setup',d['config']['setup'])"
timeout 600 python tests/tools/build_config4.py 2>&1 | tail -3
