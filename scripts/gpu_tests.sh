set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --W 8 2>&1 | tail -2
