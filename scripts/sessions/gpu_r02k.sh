set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 600 python tests/tools/exp_adaptive_w.py 2>&1 | grep w_after
timeout 600 python -m pytest tests/test_dataset_benchmark.py tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python tests/tools/dataset_benchmark_ab.py 20000 200 > gpurun_out/r02k_dataset_benchmark_ab.json 2> gpurun_out/r02k_ab.err; tail -c 300 gpurun_out/r02k_ab.err; head -c 1500 gpurun_out/r02k_dataset_benchmark_ab.json
