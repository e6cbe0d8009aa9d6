set -x
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
timeout 900 python -m pytest tests/test_lut_tc_gpu.py tests/test_search_gpu.py -q -m gpu -x 2>&1 | tail -4
bash scripts/gpu_ab.sh "" default default
python -c "
import json
" 
tail -3 gpurun_out/ab_default.err
