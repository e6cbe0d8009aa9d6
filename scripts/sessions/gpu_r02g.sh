bash scripts/gpu_ab.sh "" default nt128 nt192 nt224 nt128:"--W 4" nt192:"--W 6" default
