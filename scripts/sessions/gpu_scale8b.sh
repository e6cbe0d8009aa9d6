set -x
N=8
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
nvidia-smi topo -m | head -14
lscpu | grep -E "NUMA|Socket|^CPU\(s\)"
cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale8b_weak.json 2> gpurun_out/r02_scale8b_weak.err; tail -c 200 gpurun_out/r02_scale8b_weak.err
python -c "
import json; d=json.load(open('gpurun_out/r02_scale8b_weak.json')); print('SCALE8b weak value',d['value'],'e2e',d['e2e']['value'],'aff',d['config'].get('cpu_affinity'))"
