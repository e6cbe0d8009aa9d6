# compute-sanitizer over the smoke test (every search mode + the table kernels + a dynamic update) and over a small build / PQ train:
# memcheck (out-of-bounds, misaligned), then racecheck (shared-memory hazards) and synccheck on the smoke test.
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
cat > /tmp/san_small.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from diskrag_b200 import ops
from diskrag_b200.engine import GpuIndex
from diskrag_b200.pq.fast_pq import DiskANNPQ
from diskrag_b200.synth import synth_numpy
X = synth_numpy(3000, 64, seed=1); Q = synth_numpy(40, 64, seed=1, sample_seed=5)
pq = DiskANNPQ(8); pq.train_iters = 3; pq.fit(X); codes = pq.encode(X)
adj, deg = ops.vamana_build(X, 16, 32, 1.2, 0, seed=1)
cb = np.stack([km.cluster_centers_ for km in pq.kmeans_list]).astype(np.float32)
with GpuIndex.from_arrays(X, adj, codes=codes, codebook=cb, medoid=0) as idx:
    for kw in (dict(W=1, dist="exact", rerank=False), dict(W=4, dist="exact", rerank=False), dict(W=1, dist="pq", adc_order="seq", rerank=True),
               dict(W=8, dist="pq", rerank=True, lut_fmt="u8tc", prefetch=5, w2=20), dict(W=8, dist="pq", rerank=True, lut_fmt="u8")):
        r = idx.search(Q, k=10, L=50, **kw)
        assert r.ids.shape == (40, 10)
print("SAN_SMALL ok")
PY
for tool in memcheck; do
  echo "== $tool smoke"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE ok')" 2>&1 | grep -E "ERROR SUMMARY|SMOKE ok|Invalid|misaligned|out of bounds|hazard" | head -8
  echo "== $tool small build + searches"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_small.py 2>&1 | grep -E "ERROR SUMMARY|SAN_SMALL ok|Invalid|misaligned|out of bounds" | head -8
done
for tool in racecheck synccheck; do
  echo "== $tool small build + searches"; timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_small.py 2>&1 | grep -E "ERROR SUMMARY|SAN_SMALL ok|hazard|RACECHECK SUMMARY|divergent|Barrier error" | head -12
done
