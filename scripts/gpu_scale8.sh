# N GPUs: weak (driver's SCALE shape), strong (configs[2] literally), index-sharded configs[4] with both exchanges
# usage: gpu_scale8.sh [N=8] [TAG=r03] [legs="weak strong is_nccl is_p2p"]
set -x
N=${1:-8}; TAG=${2:-r03}; LEGS=${3:-"weak strong is_nccl is_p2p"}
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
run() { timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline $2 > gpurun_out/${TAG}_scale${N}_$1.json 2> gpurun_out/${TAG}_scale${N}_$1.err; tail -c 200 gpurun_out/${TAG}_scale${N}_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/${TAG}_scale${N}_$1.json')); print('SCALE$N $1 value',d['value'],'e2e',d['e2e'],'ms',d['ms_per_step'],'recall',d['config']['recall_at_10'],'frac',d['roofline']['frac'],'aff',d['config'].get('cpu_affinity'))"; }
for leg in $LEGS; do
  case $leg in
    weak) run weak "";;
    strong) run strong "--scaling strong";;
    is_nccl) run is_nccl "--mode index-sharded --exchange nccl";;
    is_p2p) run is_p2p "--mode index-sharded --exchange p2p";;
  esac
done
