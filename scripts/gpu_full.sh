# usage: gpu_full.sh <tag>: whole device suite, bench (our arm, all legs), reference arm
TAG=$1
set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); r=d['roofline']; c=d['config']
print('value',d['value'],'e2e',d['e2e']['value'],'recall',c['recall_at_10'],'frac',r['frac'],'kms',r['kernel_ms_per_launch'],'kshare',r['kernel_share_of_step'],'hops',c['mean_hops'],'vis',c['mean_visited'])
print('parity',c['parity_mode']); print('points',c['other_operating_points']); print('cpu',d['cpu_baseline'])"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference.json 2> gpurun_out/${TAG}_reference.err; tail -c 300 gpurun_out/${TAG}_reference.err; cat gpurun_out/${TAG}_reference.json | head -c 600
