# racecheck over a small build + every search mode: one line per (kind, kernel, source line) with its count
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
    rc = idx.beam_search_c(Q[:8], k=5, beam_width=8, dist="pq")
print("SAN_SMALL ok")
PY
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python /tmp/san_small.py > gpurun_out/racecheck_analysis.txt 2>&1
grep -E "SAN_SMALL ok|RACECHECK SUMMARY" gpurun_out/racecheck_analysis.txt | head
grep -E "Race reported|and (Read|Write) access" gpurun_out/racecheck_analysis.txt | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -60
