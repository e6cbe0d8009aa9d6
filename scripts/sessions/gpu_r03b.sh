set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_golden_config0.py tests/test_config1_parity_gpu.py tests/test_build_gpu.py -q -m gpu -x 2>&1 | tail -3
bash scripts/gpu_ab.sh "" default w16:"--W2 16" default
