set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 1500 python -m pytest tests -q -m gpu -x --durations=5 2>&1 | tail -25
cat gpurun_out/insert_replay.json
