set -x
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -3
bash scripts/profile.sh r01a
