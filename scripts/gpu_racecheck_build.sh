cat > /tmp/san_build.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from diskrag_b200 import ops
from diskrag_b200.synth import synth_numpy
X = synth_numpy(1500, 64, seed=1)
adj, deg = ops.vamana_build(X, 16, 32, 1.2, 0, seed=1)
print("SAN_BUILD ok")
PY
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/san_build.py 2>&1 | grep -v "^=========     Saved host backtrace\|^=========     Host Frame\|^=========         Host Frame" | head -60
