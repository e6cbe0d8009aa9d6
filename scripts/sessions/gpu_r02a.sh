# round 2, first GPU call: the whole device suite (no xfails any more, configs[1] parity at 100k included) + a baseline bench line
set -x
mkdir -p gpurun_out
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tee gpurun_out/r02a_pytest.log | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 600 gpurun_out/r02a_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02a_bench.json')); r=d['roofline']
print('value',d['value'],'e2e',d['e2e']['value'],'recall',d['config']['recall_at_10'],'frac',r['frac'],'kms',r['kernel_ms_per_launch'],'kshare',r['kernel_share_of_step'])"
