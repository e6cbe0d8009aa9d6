set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for ex in nccl p2p; do
  timeout 900 $TR bench.py --gpus 2 --mode index-sharded --exchange $ex --steps 4 --warmup 3 > gpurun_out/r03h_is_$ex.json 2> gpurun_out/r03h_is_$ex.err; tail -c 400 gpurun_out/r03h_is_$ex.err
  python -c "
import json; d=json.load(open('gpurun_out/r03h_is_$ex.json')); print('IS $ex value',d['value'],'e2e',d['e2e'],'recall',d['config']['recall_at_10'],'ms',d['ms_per_step'])"
done
timeout 300 $TR tests/tools/multi_gpu_check.py 2>&1 | grep "world="
