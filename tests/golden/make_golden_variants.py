#!/usr/bin/env python3
"""Generate tests/golden/ref_variants_ce.npz: outputs of the REAL reference for the two search functions the product does not copy
by default — variant C (pydiskann.vamana_graph.beam_search_with_pq, with its inverted frontier truncation) and variant E
(search_engine.SearchEngineCorrect._pq_accelerated_graph_search, the stochastic served search, under np.random.seed) — on a small
reference-built index.  Run in the build container only (needs oracle/_ref and /root/reference/search_engine.py):

    python tests/golden/make_golden_variants.py

Everything under `exp_*` was computed by reference code, nothing by ours.  The index reuses ref_small.npz (vectors, reference-built
graph, sklearn codes / codebook), so the fixture only adds queries' outputs (a few KB)."""
import contextlib
import io as _io
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

m = ref_loader.load()
vg, fp, dp = m["vamana_graph"], m["fast_pq"], m["diskann_persist"]
se = _load_reference_search_engine()
quiet = contextlib.redirect_stdout(_io.StringIO())

z = np.load(ROOT / "tests" / "golden" / "ref_small.npz")
N, D, R, M, med = int(z["N"]), int(z["D"]), int(z["R"]), int(z["M"]), int(z["medoid"])
rec32 = z["records"].view(np.uint32).reshape(N, D + R)
X = rec32[:, :D].copy().view(np.float32); adj = rec32[:, D:].copy()
codes, cb, Q = z["codes"], z["codebook"], z["Q"][:24]

pq = fp.DiskANNPQ(M, 256)
pq.sub_dim = D // M; pq.is_fitted = True
pq.kmeans_list = [_wrap_kmeans(cb[i], 42 + i) for i in range(M)]
g = vg.VamanaGraphWithPQ(R, pq)
for i in range(N):
    node = vg.Node(i, X[i], codes[i])
    node.neighbors = [int(x) for x in adj[i]]            # index.dat order, 0-padded
    g.nodes[i] = node
g.medoid_idx = med

out = {}
# ---- variant C -------------------------------------------------------------------------------------------------------------
SHAPES_C = [(5, 3), (8, 5), (2, 10), (16, 10), (64, 10)]
deleted = np.zeros(N, np.uint8)
deleted[[i for i in (4, 9, 100, 101, 102, 777, 1503, 1999) if i != med]] = 1
for tag, dead in (("live", np.zeros(N, np.uint8)), ("del", deleted)):
    for i in range(N):
        g.nodes[i].is_deleted = bool(dead[i])
    for bw, k in SHAPES_C:
        for use_pq in (True, False):
            ids = np.full((len(Q), k), -1, np.int32); dd = np.full((len(Q), k), np.inf, np.float64)
            for qi, q in enumerate(Q):
                with quiet:
                    r = vg.beam_search_with_pq(g, q, med, bw, k, use_pq=use_pq)
                for j, (d, i) in enumerate(r):
                    ids[qi, j] = i; dd[qi, j] = d
            key = f"exp_C_{tag}_{'pq' if use_pq else 'l2'}_bw{bw}_k{k}"
            out[key + "_ids"] = ids; out[key + "_d"] = dd
for i in range(N):
    g.nodes[i].is_deleted = False
out["deleted"] = deleted
out["shapes_c"] = np.array(SHAPES_C, np.int32)

# ---- variant E -------------------------------------------------------------------------------------------------------------
tmp = tempfile.mkdtemp()
for node in g.nodes.values():
    row = [int(x) for x in adj[node.idx]]
    node.neighbors = row
dp.DiskANNPersist(dim=D, R=R).save_index(tmp + "/index.dat", g)
assert np.array_equal(np.fromfile(tmp + "/index.dat", np.uint8), z["records"])
eng = object.__new__(se.SearchEngineCorrect)
eng.reader = dp.MMapNodeReader(tmp + "/index.dat", dim=D, R=R)
eng.pq_model = pq; eng.pq_codes = codes; eng.n_subvectors = M; eng.sub_dim = D // M; eng.num_centroids = 256
eng.meta = {"N": N}; eng.medoid_idx = med; eng.use_pq = True
eng.search_stats = {"total_searches": 0, "total_exact_computations": 0, "total_pq_computations": 0, "total_search_time": 0.0}
eng.use_thread_safe_stats = True; eng._stats_lock = threading.Lock()
SHAPES_E = [(40, 10, 0), (40, 10, 8), (100, 10, 8), (10, 5, 4)]          # L, k, beam_width (0 = None)
for seed, (L, k, bw) in enumerate(SHAPES_E):
    np.random.seed(seed)
    ids = np.full((len(Q), k), -1, np.int32); dd = np.full((len(Q), k), np.inf, np.float32); st = np.zeros((len(Q), 4), np.int32)
    for qi, q in enumerate(Q):
        r, s = eng._pq_accelerated_graph_search(q, k=k, L=L, beam_width=bw or None)
        for j, (d, i) in enumerate(r):
            ids[qi, j] = i; dd[qi, j] = d
        st[qi] = [s["nodes_visited"], s["exact_distance_computations"], s["pq_distance_computations"], s["search_steps"]]
    out[f"exp_E_seed{seed}_ids"] = ids; out[f"exp_E_seed{seed}_d2"] = dd; out[f"exp_E_seed{seed}_stats"] = st
    out[f"exp_E_seed{seed}_next_draw"] = np.float64(np.random.random())     # the generator's position after the 24 queries
out["shapes_e"] = np.array(SHAPES_E, np.int32)
out["nq"] = np.int32(len(Q))
dst = ROOT / "tests" / "golden" / "ref_variants_ce.npz"
np.savez_compressed(dst, **out)
print(dst, dst.stat().st_size, "bytes;", len(out), "arrays")
