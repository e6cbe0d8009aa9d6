# ncu evidence for the round (B200_PROFILING.md recipe).  Usage: bash scripts/profile.sh <tag>
set -x
TAG=${1:-r01}
ARGS="--queries 20000 --steps 1 --warmup 1 --gt-queries 200 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_search \
    python bench.py $ARGS > gpurun_out/${TAG}_full_bench.log 2>&1
ls -la gpurun_out/
