#!/usr/bin/env python3
"""Generate tests/golden/ref_config0.npz: BASELINE configs[0] — the REAL reference (oracle/_ref, compiled from /root/reference by
oracle/build_ref.py) on 10k x 1536 synthetic vectors: sklearn-trained DiskANNPQ with M = 64 (the adaptive default at D = 1536,
adaptive_pq.py:81-108), build_vamana_with_pq R = 32, L = 64, alpha = 1.2, the graph written by DiskANNPersist.save_index, then the
reference's own searches on 64 held-out queries.  Run in the build container only (about 8 minutes of CPU):

    python tests/golden/make_golden_config0.py

The 61 MB of vectors are NOT stored: everything the PQ traversal needs is (adjacency as index.dat holds it, codes, codebook, queries),
and the vectors are regenerated from the seed by the tests, which compare their SHA-256 with the one stored here and run the checks
that need full vectors (variant D, rerank) only when it matches (BLAS / LAPACK builds may differ in the last bit across machines).
Everything stored under `exp_*` was computed by reference code, nothing by ours."""
import contextlib
import hashlib
import io as _io
import os
import random
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import ref_loader  # noqa: E402
from diskrag_b200.synth import synth_numpy  # noqa: E402

m = ref_loader.load()
cu, vg, fp, dp = m["cython_utils"], m["vamana_graph"], m["fast_pq"], m["diskann_persist"]
quiet = contextlib.redirect_stdout(_io.StringIO())

N, D, M, R, LB, NQ = 10000, 1536, 64, 32, 64, 64
SEED = 20240                                       # base seed 20240 + config index 0
X = synth_numpy(N, D, seed=SEED)
Q = synth_numpy(NQ, D, seed=SEED, sample_seed=1000)
t0 = time.time()
with quiet:
    pq = fp.DiskANNPQ(M, 256)
    pq.fit(X)
codes = pq.encode(X)
codebook = np.stack([k.cluster_centers_ for k in pq.kmeans_list]).astype(np.float32)
print(f"PQ fit + encode {time.time() - t0:.0f} s", flush=True)

random.seed(5)
t0 = time.time()
with quiet:
    g = vg.build_vamana_with_pq(X, pq, R=R, L=LB, alpha=1.2)
print(f"build_vamana_with_pq {time.time() - t0:.0f} s", flush=True)
medoid = int(g.medoid_idx)
tmp = tempfile.mkdtemp()
idx_path = os.path.join(tmp, "index.dat")
dp.DiskANNPersist(dim=D, R=R).save_index(idx_path, g)
rec32 = np.fromfile(idx_path, dtype=np.uint32).reshape(N, D + R)
vec = rec32[:, :D].copy().view(np.float32)
adj = rec32[:, D:].copy()
assert np.array_equal(vec, X) and adj.max() < 65536

out = dict(N=N, D=D, M=M, R=R, LB=LB, seed=SEED, medoid=medoid, adj16=adj.astype(np.uint16), codes=codes, codebook=codebook, Q=Q,
           x_sha256=np.frombuffer(hashlib.sha256(X.tobytes()).digest(), np.uint8))

# variant A (greedy_search_cython + compute_query_distance with PQ on) on the graph as index.dat holds it
gf = vg.VamanaGraphWithPQ(R, pq)
for i in range(N):
    gf.add_node(i, vec[i], codes[i])
    gf.nodes[i].neighbors = [int(x) for x in adj[i]]
gf.medoid_idx = medoid
gf.use_pq_for_search = True
t0 = time.time()
for L in (64, 100):
    ids = np.full((NQ, L), -1, np.int32)
    dist = np.full((NQ, L), np.inf, np.float32)
    for qi in range(NQ):
        gf._distance_table_cache.clear()
        r = cu.greedy_search_cython(gf, medoid, Q[qi], L, vg.compute_query_distance)
        T = pq.compute_distance_table(Q[qi])
        ids[qi, :len(r)] = r
        dist[qi, :len(r)] = pq.asymmetric_distance_sq(codes[r], T)
    out[f"exp_A_ids_L{L}"] = ids
    out[f"exp_A_dist_L{L}"] = dist
print(f"variant A {time.time() - t0:.0f} s", flush=True)
out["exp_lut0"] = pq.compute_distance_table(Q[0])

# rerank composition (search_engine.py:374-379): ids from A (L = 100), exact d2, stable sort, top 10
rr_ids = np.empty((NQ, 10), np.int32); rr_d = np.empty((NQ, 10), np.float32)
for qi in range(NQ):
    ids = out["exp_A_ids_L100"][qi]
    ids = ids[ids >= 0]
    d2 = np.array([np.sum((vec[i] - Q[qi]) * (vec[i] - Q[qi])) for i in ids], np.float32)
    o = np.argsort(d2, kind="stable")[:10]
    rr_ids[qi] = ids[o]; rr_d[qi] = d2[o]
out["exp_rerank_ids"] = rr_ids; out["exp_rerank_d2"] = rr_d

# variant D (beam_search_from_disk, beam_width = L = 64)
reader = dp.MMapNodeReader(idx_path, dim=D, R=R)
d_ids = np.empty((NQ, 10), np.int32); d_dist = np.empty((NQ, 10), np.float32)
for qi in range(NQ):
    r = vg.beam_search_from_disk(reader, Q[qi], medoid, beam_width=64, k=10)
    d_ids[qi] = [int(i) for _, i in r]; d_dist[qi] = [float(d) for d, _ in r]
out["exp_D_ids"] = d_ids; out["exp_D_dist"] = d_dist

# brute-force ground truth as dataset_benchmark.compute_ground_truth does (dataset_benchmark.py:62-73) -> the reference's recall@10
gt = np.array([np.argsort(np.linalg.norm(X - q, axis=1))[:10] for q in Q], np.int32)
out["gt"] = gt
rec = lambda ids: float(np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(NQ)]))
out["recall_D"] = rec(d_ids); out["recall_rerank"] = rec(rr_ids); out["recall_A_top10"] = rec(out["exp_A_ids_L100"][:, :10])
print("recall@10: variant D", out["recall_D"], "A+rerank", out["recall_rerank"], "A (PQ order)", out["recall_A_top10"], flush=True)

p = ROOT / "tests" / "golden" / "ref_config0.npz"
np.savez_compressed(p, **out)
print("wrote", p, os.path.getsize(p), "bytes")
