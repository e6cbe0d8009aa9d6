set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
ARGS="--queries 20000 --steps 1 --warmup 1 --gt-queries 200 --no-cpu-baseline --no-points --cuda-profile"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_fast -c 1 -f \
    -o gpurun_out/r02c_search python bench.py $ARGS > gpurun_out/r02c_full_bench.log 2>&1
ls -la gpurun_out/
DISKRAG_B200_LIB=$PWD/diskrag_b200/variants/lib_pt.so timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-points --gt-queries 200 2>&1 | grep "\[phase\]" | tail -2
