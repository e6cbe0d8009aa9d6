# 2 GPUs: exchange correctness (NCCL packed vs peer-routed), index-sharded bench with both exchanges, strong scaling
set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "packed or topk" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR tests/tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\[W\|Warning" | tail -8
for ex in nccl p2p; do
  timeout 900 $TR bench.py --gpus 2 --mode index-sharded --exchange $ex --steps 4 --warmup 3 > gpurun_out/r02f_is_$ex.json 2> gpurun_out/r02f_is_$ex.err; tail -c 400 gpurun_out/r02f_is_$ex.err
  python -c "
import json; d=json.load(open('gpurun_out/r02f_is_$ex.json')); print('IS $ex value',d['value'],'e2e',d['e2e']['value'],'recall',d['config']['recall_at_10'],'ms',d['ms_per_step'],'kshare',d['roofline']['kernel_share_of_step'],'frac',d['roofline']['frac'])"
done
timeout 900 $TR bench.py --gpus 2 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_strong2.json 2> gpurun_out/r02f_strong2.err; tail -c 300 gpurun_out/r02f_strong2.err
python -c "
import json; d=json.load(open('gpurun_out/r02f_strong2.json')); print('STRONG2 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'aff',d['config']['cpu_affinity'])"
