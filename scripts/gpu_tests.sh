set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import oracle" 2>/dev/null
cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm && cd ..
timeout 1200 python -m pytest tests/test_search_gpu.py -q -m gpu 2>&1 | tail -40
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -40
