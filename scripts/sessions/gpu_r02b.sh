set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 900 python -m pytest tests/test_search_gpu.py -q -m gpu -x 2>&1 | tail -4
bash scripts/gpu_ab.sh "" default pf13:"--prefetch 13" pf21:"--prefetch 21" pf29:"--prefetch 29" default
