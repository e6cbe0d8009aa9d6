#!/usr/bin/env python3
"""Generate tests/golden/ref_small.npz by running the REAL reference (oracle/_ref, built from
/root/reference by oracle/build_ref.py).  Run in the build container only; the GPU box has no
/root/reference and only reads the committed .npz.

    python tests/golden/make_golden.py

Everything stored under `exp_*` was computed by reference code (pydiskann.*), nothing by ours.
"""
import contextlib
import io as _io
import os
import random
import sys
import tempfile
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

N, D, M, R, LB, NQ = 2000, 64, 8, 16, 32, 48
X = synth_numpy(N, D, seed=11, K=48, r=16)
Q = synth_numpy(NQ, D, seed=11, sample_seed=1, K=48, r=16)
# a few exact duplicate rows: identical codes and distances -> exercises the tie rules
X[1500:1510] = X[100:110]

with quiet:
    pq = fp.DiskANNPQ(M, 256)
    pq.fit(X)
codes = pq.encode(X)
codebook = np.stack([k.cluster_centers_ for k in pq.kmeans_list]).astype(np.float32)

random.seed(3)
with quiet:
    g = vg.build_vamana_with_pq(X, pq, R=R, L=LB, alpha=1.2)
medoid = int(g.medoid_idx)
tmp = tempfile.mkdtemp()
idx_path = os.path.join(tmp, "index.dat")
dp.DiskANNPersist(dim=D, R=R).save_index(idx_path, g)
records = np.fromfile(idx_path, dtype=np.uint8)
assert records.size == N * 4 * (D + R)
rec32 = records.view(np.uint32).reshape(N, D + R)
vec = rec32[:, :D].copy().view(np.float32)
adj = rec32[:, D:].copy()

out = dict(N=N, D=D, M=M, R=R, medoid=medoid, records=records, codes=codes, codebook=codebook, Q=Q)

# K2: ADC tables
out["exp_lut"] = np.stack([pq.compute_distance_table(Q[i]) for i in range(8)])

# variant A (greedy_search_cython + compute_query_distance, PQ on); graph rebuilt with file-ordered rows
gf = vg.VamanaGraphWithPQ(R, pq)
for i in range(N):
    gf.add_node(i, vec[i], codes[i])
    gf.nodes[i].neighbors = [int(x) for x in adj[i]]
gf.medoid_idx = medoid
gf.use_pq_for_search = True
for L in (10, 40):
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

# rerank composition: ids from A (L=40), exact d2 as search_engine.py:374-379, stable sort, top 10
rr_ids = np.empty((NQ, 10), np.int32); rr_d = np.empty((NQ, 10), np.float32)
for qi in range(NQ):
    ids = out["exp_A_ids_L40"][qi]
    ids = ids[ids >= 0]
    d2 = np.array([np.sum((vec[i] - Q[qi]) * (vec[i] - Q[qi])) for i in ids], np.float32)
    o = np.argsort(d2, kind="stable")[:10]
    rr_ids[qi] = ids[o]; rr_d[qi] = d2[o]
out["exp_rerank_ids"] = rr_ids
out["exp_rerank_d2"] = rr_d

# variant D (beam_search_from_disk) and B (greedy_search, exact)
reader = dp.MMapNodeReader(idx_path, dim=D, R=R)
d_ids = np.empty((NQ, 10), np.int32); d_dist = np.empty((NQ, 10), np.float32)
for qi in range(NQ):
    r = vg.beam_search_from_disk(reader, Q[qi], medoid, beam_width=40, k=10)
    d_ids[qi] = [int(i) for _, i in r]; d_dist[qi] = [float(d) for d, _ in r]
out["exp_D_ids"] = d_ids; out["exp_D_dist"] = d_dist
gf.use_pq_for_search = False
b_ids = np.full((NQ, 40), -1, np.int32)
for qi in range(NQ):
    r = vg.greedy_search(gf, medoid, Q[qi], 40)
    b_ids[qi, :len(r)] = r
out["exp_B_ids"] = b_ids

# known-answer distances (scripts/test_pydiskann_cython.sh:36-56 uses RandomState(0).randn(128))
rs = np.random.RandomState(0)
ka_x = rs.randn(16, 128).astype(np.float32); ka_y = rs.randn(16, 128).astype(np.float32)
out["ka_x"] = ka_x; out["ka_y"] = ka_y
out["exp_l2"] = np.array([cu.l2_distance_fast_cython(ka_x[i], ka_y[i]) for i in range(16)], np.float64)
out["exp_cos"] = np.array([cu.cosine_similarity_cython(ka_x[i], ka_y[i]) for i in range(16)], np.float64)
out["exp_sdc"] = np.array([cu.pq_distance_fast_cython(pq, codes[i], codes[i + 1]) for i in range(16)], np.float64)

# medoid, deterministic branch (n <= sample_size)
out["exp_medoid_500"] = int(cu.compute_approximate_medoid_cython(X[:500], sample_size=1000))

# sequential build with fixed permutations
NB, RB, LBB = 600, 8, 16
random.seed(7)
st = random.getstate()
adj_ref = cu.build_vamana_index_cython(X[:NB], RB, LBB, 1.2, 5, False)
random.setstate(st)
s0 = list(range(NB)); random.shuffle(s0)
s1 = list(range(NB)); random.shuffle(s1)
ba = np.full((NB, RB), -1, np.int32)
for i, row in enumerate(adj_ref):
    ba[i, :len(row)] = row
out["build_sigma0"] = np.array(s0, np.int32); out["build_sigma1"] = np.array(s1, np.int32)
out["exp_build_adj"] = ba

np.savez_compressed(ROOT / "tests" / "golden" / "ref_small.npz", **out)
print("wrote", ROOT / "tests" / "golden" / "ref_small.npz", os.path.getsize(ROOT / "tests" / "golden" / "ref_small.npz"), "bytes")
