# usage: gpu_ncu.sh <tag> [bench args]: one --set full capture of the search kernel (20k-query launch) -> gpurun_out/<tag>_search.ncu-rep
TAG=$1; shift
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
ARGS="--queries 20000 --steps 1 --warmup 1 --gt-queries 200 --no-cpu-baseline --no-points --cuda-profile $*"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_ -c 1 -f \
    -o gpurun_out/${TAG}_search python bench.py $ARGS > gpurun_out/${TAG}_full_bench.log 2>&1
ls -la gpurun_out/${TAG}_search.ncu-rep
