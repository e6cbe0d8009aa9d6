# 8 GPUs: weak (driver's SCALE shape), strong (configs[2] literally), index-sharded configs[4] with both exchanges
set -x
N=${1:-8}
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
run() { timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline $2 > gpurun_out/r02_scale${N}_$1.json 2> gpurun_out/r02_scale${N}_$1.err; tail -c 200 gpurun_out/r02_scale${N}_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_scale${N}_$1.json')); print('SCALE$N $1 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'recall',d['config']['recall_at_10'],'frac',d['roofline']['frac'],'aff',d['config'].get('cpu_affinity'))"; }
run weak ""
run strong "--scaling strong"
run is_nccl "--mode index-sharded --exchange nccl"
run is_p2p "--mode index-sharded --exchange p2p"
