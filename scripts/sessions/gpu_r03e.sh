set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 900 python -m pytest tests/test_lut_tc_gpu.py tests/test_search_gpu.py tests/test_golden_config0.py -q -m gpu -x 2>&1 | tail -4
bash scripts/gpu_ab.sh "" default default
DR_LUT_QUERY_MAJOR=1 bash scripts/gpu_ab.sh "" default
