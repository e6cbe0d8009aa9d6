set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 1200 python -m pytest tests/test_build_gpu.py -x -q -m gpu -s 2>&1 | tail -40
timeout 600 python scripts/build_probe.py 100000 128 32 64 32 2>&1 | tail -20
