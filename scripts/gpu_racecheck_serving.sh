# racecheck + memcheck over the SERVING specialisation of the throughput kernel (D = 1536, M = 192, R = 32, L = 100, W = 8, w2 = 20:
# search_fast_kernel<48,3,4>) and the ds = 8 table kernels, on a small index
cat > /tmp/san_serving.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from diskrag_b200 import ops
from diskrag_b200.engine import GpuIndex
from diskrag_b200.pq.fast_pq import DiskANNPQ
from diskrag_b200.synth import synth_numpy
X = synth_numpy(6000, 1536, seed=1); Q = synth_numpy(96, 1536, seed=1, sample_seed=5)
pq = DiskANNPQ(192); pq.train_iters = 2; pq.fit(X); codes = pq.encode(X)
adj, deg = ops.vamana_build(X, 32, 64, 1.2, 0, seed=1)
cb = np.stack([km.cluster_centers_ for km in pq.kmeans_list]).astype(np.float32)
with GpuIndex.from_arrays(X, adj, codes=codes, codebook=cb, medoid=0) as idx:
    r = idx.search(Q, k=10, L=100, W=8, dist="pq", rerank=True, lut_fmt="u8tc", prefetch=5, w2=20)
    assert r.ids.shape == (96, 10) and (r.ids >= 0).all()
print("SAN_SERVING ok")
PY
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --racecheck-report analysis python /tmp/san_serving.py > gpurun_out/serving_$tool.txt 2>&1
  echo "== $tool"; grep -E "SAN_SERVING ok|ERROR SUMMARY|RACECHECK SUMMARY|Race reported|and (Read|Write) access|Invalid" gpurun_out/serving_$tool.txt | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head -20
  grep -c "search_fast_kernel" gpurun_out/serving_$tool.txt
done
