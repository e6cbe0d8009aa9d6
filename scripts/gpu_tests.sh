set -x
bash scripts/profile.sh r01a
