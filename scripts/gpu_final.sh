# final evidence of the round: device suite, bench (all legs), reference arm, launch list + full ncu capture of the search kernel
TAG=$1
set -x
bash scripts/gpu_full.sh $TAG 2>&1 | grep -v "^+" | tail -12
bash scripts/profile.sh ${TAG}p > /dev/null 2>&1
ls -la gpurun_out/${TAG}p_search.ncu-rep gpurun_out/${TAG}p_launches.csv
