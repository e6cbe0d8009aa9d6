#!/usr/bin/env python3
"""Generate tests/golden/ref_config0_e.npz: what the REAL reference's served search returns at BASELINE configs[0] — the two seam
methods of search_engine.SearchEngineCorrect on the ref_config0.npz index (10k x 1536, sklearn PQ M = 64, reference-built graph):

  _pq_accelerated_graph_search(q, k=10, L=100, beam_width=8)   search_engine.py:398-506  (variant E, stochastic: np.random.seed(0))
  _exact_graph_search(q, k=10, L=100)                          search_engine.py:508-528  (variant D with beam_width = 8)

for the fixture's 64 queries: (squared-L2, id) results and the stats dicts.  Run in the build container only:
    python tests/golden/make_golden_config0_e.py
Everything under `exp_*` was computed by reference code, nothing by ours."""
import sys
import tempfile
import threading
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import ref_loader  # noqa: E402
from test_oracle_vs_reference import _load_reference_search_engine  # noqa: E402
from diskrag_b200.pq.fast_pq import _wrap_kmeans  # noqa: E402
from diskrag_b200.synth import synth_numpy  # noqa: E402

m = ref_loader.load()
vg, fp, dp = m["vamana_graph"], m["fast_pq"], m["diskann_persist"]
se = _load_reference_search_engine()
z = np.load(ROOT / "tests" / "golden" / "ref_config0.npz")
N, D, M, R, med, seed = (int(z[k]) for k in ("N", "D", "M", "R", "medoid", "seed"))
X = synth_numpy(N, D, seed=seed)
adj = z["adj16"].astype(np.uint32); codes = z["codes"]; cb = z["codebook"]; Q = z["Q"]
tmp = tempfile.mkdtemp()
rec = np.zeros((N, D + R), np.uint32)
rec[:, :D] = X.view(np.uint32); rec[:, D:] = adj
rec.tofile(tmp + "/index.dat")                                         # the byte layout of DiskANNPersist.save_index
pq = fp.DiskANNPQ(M, 256)
pq.sub_dim = D // M; pq.is_fitted = True
pq.kmeans_list = [_wrap_kmeans(cb[i], 42 + i) for i in range(M)]
eng = object.__new__(se.SearchEngineCorrect)
eng.reader = dp.MMapNodeReader(tmp + "/index.dat", dim=D, R=R)
eng.pq_model = pq; eng.pq_codes = codes; eng.n_subvectors = M; eng.sub_dim = D // M; eng.num_centroids = 256
eng.meta = {"N": N}; eng.medoid_idx = med; eng.use_pq = True
eng.search_stats = {"total_searches": 0, "total_exact_computations": 0, "total_pq_computations": 0, "total_search_time": 0.0}
eng.use_thread_safe_stats = True; eng._stats_lock = threading.Lock()
k, L, bw = 10, 100, 8
nq = len(Q)
out = {}
np.random.seed(0)
ids = np.full((nq, k), -1, np.int32); dd = np.full((nq, k), np.inf, np.float32); st = np.zeros((nq, 4), np.int32)
keys = None
for qi, q in enumerate(Q):
    r, s = eng._pq_accelerated_graph_search(q, k=k, L=L, beam_width=bw)
    for j, (d, i) in enumerate(r):
        ids[qi, j] = i; dd[qi, j] = d
    st[qi] = [s["nodes_visited"], s["exact_distance_computations"], s["pq_distance_computations"], s["search_steps"]]
    keys = sorted(s.keys())
out["exp_E_ids"] = ids; out["exp_E_d2"] = dd; out["exp_E_stats"] = st
out["exp_E_stat_keys"] = np.array(keys)
out["exp_E_result_types"] = np.array([type(r[0][0]).__name__, type(r[0][1]).__name__])
ids = np.full((nq, k), -1, np.int64); dd = np.full((nq, k), np.inf, np.float32); lens = np.zeros(nq, np.int32)
for qi, q in enumerate(Q):
    r, s = eng._exact_graph_search(q, k=k, L=L)
    lens[qi] = len(r)
    for j, (d, i) in enumerate(r):
        ids[qi, j] = i; dd[qi, j] = d
    xkeys = sorted(s.keys()); xs = s
out["exp_X_ids"] = ids; out["exp_X_dist"] = dd; out["exp_X_len"] = lens
out["exp_X_stat_keys"] = np.array(xkeys); out["exp_X_search_type"] = np.array(xs["search_type"])
out["exp_X_exact_computations"] = np.int32(xs["exact_distance_computations"])
gt = z["gt"]
rec_ = lambda a: float(np.mean([len(set(a[i].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
out["recall_E"] = rec_(out["exp_E_ids"]); out["recall_X"] = rec_(out["exp_X_ids"])
print("recall@10 served PQ search (E)", out["recall_E"], "exact fallback (bw=8)", out["recall_X"], "mean exact computations", st[:, 1].mean(),
      "stat keys", keys, xkeys, "result types", out["exp_E_result_types"])
np.savez_compressed(ROOT / "tests" / "golden" / "ref_config0_e.npz", **out)
