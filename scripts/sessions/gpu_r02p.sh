bash scripts/gpu_ab.sh "" default ml96 ml128
for f in 0.3 0.6 1.0; do
DR_L2_PERSIST=$f timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --gt-queries 300 --no-points 2> gpurun_out/l2p.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('L2PERSIST $f | value',d['value'],'frac',r['frac'],'kms',r['kernel_ms_per_launch'])"
grep "l2 persist" gpurun_out/l2p.err | head -1
done
