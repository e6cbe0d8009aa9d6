#!/usr/bin/env python3
"""Generate tests/golden/ref_inmem.npz by running the REAL reference (oracle/_ref) on a LIVE in-memory graph:

  * in-memory search semantics (vamana_graph.py:607-640, cython_utils.pyx:72-122 iterate `node.neighbors`: true degree, the
    set's iteration order, NO padding) on a reference-built graph with short rows — variant B (`greedy_search`, exact) and
    variant A (`greedy_search_cython` + ADC callback);
  * dynamic updates (vamana_graph.py:58-125): a scripted sequence of `insert_node` (1000 new points) and `delete_node`, then
    searches on the mutated graph — the final neighbour sets (rows in the sets' iteration order, so longer than R where
    reverse edges piled up), delete flags and search results.  Re-enabling a deleted id is not scripted: the reference raises
    UnboundLocalError there (`new_node` is only bound on the new-id path, vamana_graph.py:100), checked at the end.

Run in the build container only:   python tests/golden/make_golden_inmem.py
Everything stored under `exp_*` / `rows_*` was computed by reference code, nothing by ours.  Vectors are regenerated from seeds."""
import contextlib
import io as _io
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import ref_loader  # noqa: E402
from diskrag_b200.synth import synth_numpy  # noqa: E402

m = ref_loader.load()
cu, vg, fp = m["cython_utils"], m["vamana_graph"], m["fast_pq"]
quiet = contextlib.redirect_stdout(_io.StringIO())

N0, NI, D, M, R, LB, NQ = 1000, 1000, 48, 8, 12, 24, 40       # N0 <= 1000: the reference's medoid is deterministic (all points sampled)
SEED = 23
X = synth_numpy(N0 + NI, D, seed=SEED, K=32, r=12)
Q = synth_numpy(NQ, D, seed=SEED, sample_seed=1, K=32, r=12)
X[N0 + 5] = X[17]; X[N0 + 6] = X[17]                            # exact duplicates among the inserted points: distance ties

with quiet:
    pq = fp.DiskANNPQ(M, 256)
    pq.fit(X[:N0])
codebook = np.stack([k.cluster_centers_ for k in pq.kmeans_list]).astype(np.float32)
random.seed(5)
with quiet:
    g = vg.build_vamana_with_pq(X[:N0], pq, R=R, L=LB, alpha=1.0)    # alpha = 1: sparse rows, many shorter than R
medoid = int(g.medoid_idx)


def rows_of(graph, n):
    w = max(len(graph.nodes[i].neighbors) for i in range(n))
    rows = np.full((n, w), 0xFFFFFFFF, np.uint32); deg = np.zeros(n, np.int32)
    for i in range(n):
        nb = list(graph.nodes[i].neighbors)                      # the order the reference's searches iterate in
        rows[i, :len(nb)] = nb; deg[i] = len(nb)
    return rows, deg


def searches(graph, n, tag, out):
    for L in (10, 32):
        b = np.full((NQ, L), -1, np.int32)
        a = np.full((NQ, L), -1, np.int32); ad = np.full((NQ, L), np.inf, np.float32)
        for qi in range(NQ):
            graph.use_pq_for_search = False
            r = vg.greedy_search(graph, graph.medoid_idx, Q[qi], L)                       # variant B, in memory
            b[qi, :len(r)] = r
            graph.use_pq_for_search = True
            graph._distance_table_cache.clear()
            r = cu.greedy_search_cython(graph, graph.medoid_idx, Q[qi], L, vg.compute_query_distance)   # variant A, in memory
            T = pq.compute_distance_table(Q[qi])
            a[qi, :len(r)] = r
            ad[qi, :len(r)] = pq.asymmetric_distance_sq(np.stack([graph.nodes[i].pq_code for i in r]), T)
        graph.use_pq_for_search = False
        out[f"exp_B_ids_L{L}_{tag}"] = b; out[f"exp_A_ids_L{L}_{tag}"] = a; out[f"exp_A_dist_L{L}_{tag}"] = ad


out = dict(N0=N0, NI=NI, D=D, M=M, R=R, seed=SEED, medoid=medoid, codebook=codebook, Q=Q,
           codes0=np.stack([g.nodes[i].pq_code for i in range(N0)]))
rows0, deg0 = rows_of(g, N0)
out["rows_built"] = rows0; out["deg_built"] = deg0
assert deg0.min() < R, "fixture needs rows shorter than R"
searches(g, N0, "built", out)

# ---- dynamic updates: the script (op, id): 0 = insert next new point, 1 = delete id ----
rng = np.random.default_rng(99)
script = []
nxt = N0
for step in range(NI):
    script.append((0, nxt)); nxt += 1
    if step % 25 == 24:
        victim = int(rng.integers(0, nxt))
        if victim != medoid:
            script.append((1, victim))
script = np.array(script, np.int64)
alive_again = {}
with quiet:
    for op, i in script:
        i = int(i)
        if op == 0:
            g.insert_node(i, X[i])
        else:
            g.delete_node(i)
n_all = len(g.nodes)
rows1, deg1 = rows_of(g, n_all)
out["script"] = script
out["rows_final"] = rows1; out["deg_final"] = deg1
out["deleted_final"] = np.array([g.nodes[i].is_deleted for i in range(n_all)], np.uint8)
out["vec_src_final"] = np.array([alive_again.get(i, i) for i in range(n_all)], np.int64)   # row of X each node's vector is
out["codes_final"] = np.stack([g.nodes[i].pq_code for i in range(n_all)])
out["medoid_final"] = int(g.medoid_idx)
searches(g, n_all, "final", out)
# the re-enable path of the reference (on a scratch copy of one deleted node's state)
dead = int(np.flatnonzero(out["deleted_final"])[0])
try:
    with quiet:
        g.insert_node(dead, X[dead])
    out["reenable_raises"] = 0
except UnboundLocalError:
    out["reenable_raises"] = 1
np.savez_compressed(ROOT / "tests" / "golden" / "ref_inmem.npz", **out)
print("rows built: width", rows0.shape[1], "min/mean deg", deg0.min(), deg0.mean(), "| final: width", rows1.shape[1], "max deg", deg1.max(),
      "deleted", int(out["deleted_final"].sum()), "re-enabled", len(alive_again))
