set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_golden_config0.py tests/test_lut_tc_gpu.py tests/test_build_index_gpu.py -q -m gpu -x 2>&1 | tail -6
cat gpurun_out/codebook_ab_config0.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-points --gt-queries 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'recall',d['config']['recall_at_10'],'setup',d['config']['setup'])"
