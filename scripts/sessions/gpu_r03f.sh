set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_build_gpu.py tests/test_golden_inmem.py -q -m gpu -x 2>&1 | tail -4
