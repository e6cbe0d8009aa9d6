# ncu evidence for the round (B200_PROFILING.md recipe).  Usage: bash scripts/profile.sh <tag> [extra bench args]
set -x
TAG=${1:-r01}; shift
ARGS="--queries 20000 --steps 1 --warmup 1 --gt-queries 200 --no-cpu-baseline --no-points --cuda-profile $*"
# launch list of one profiled step (every kernel of the step with its device time)
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_launches_bench.log 2>&1
# the dominant kernel, full set with source
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:search_fast -c 1 -f \
    -o gpurun_out/${TAG}_search python bench.py $ARGS > gpurun_out/${TAG}_full_bench.log 2>&1
ls -la gpurun_out/
