"""One-off check (about 2 minutes of one CPU core): the oracle's sequential Vamana build (oracle.c:orc_vamana_build, the restatement of
build_vamana_index_cython, cython_utils.pyx:269-492) against the graph the REAL reference built for the BASELINE configs[0] fixture
(tests/golden/ref_config0.npz: 10k x 1536, R = 32, L = 64, alpha = 1.2, Python `random` seeded with 5 before build_vamana_with_pq, which
draws nothing from it before the builder's two shuffles).  Row-for-row comparison with the index.dat adjacency AS SETS: the reference
keeps `Node.neighbors` in a Python set (vamana_graph.py:8-16), so save_index writes each row in set-iteration order, 0-padded.  This is what the
"reference-equivalent graph" of the configs[1] / configs[3] parity runs rests on, at the real dimension.

    python tests/tools/check_build_config0.py          -> one JSON line"""
import hashlib
import json
import random
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O  # noqa: E402
from diskrag_b200.synth import synth_numpy  # noqa: E402

z = np.load(ROOT / "tests" / "golden" / "ref_config0.npz")
N, D, R, LB, seed, med = (int(z[k]) for k in ("N", "D", "R", "LB", "seed", "medoid"))
X = synth_numpy(N, D, seed=seed)
assert hashlib.sha256(X.tobytes()).digest() == z["x_sha256"].tobytes(), "regenerated vectors differ from the generator's"
random.seed(5)
s0 = list(range(N)); random.shuffle(s0)
s1 = list(range(N)); random.shuffle(s1)
O.build()
t0 = time.time()
rows = O.vamana_build(X, R, LB, 1.2, med, np.array(s0, np.int32), np.array(s1, np.int32))
dt = time.time() - t0
adj = z["adj16"].astype(np.int64)
same = 0
jac = []
for i, row in enumerate(rows):
    want = set(row[:R]) | ({0} if len(row) < R else set())          # short rows are 0-padded in the file
    got = set(adj[i].tolist())
    same += want == got
    jac.append(len(want & got) / len(want | got))
# the same exact search (variant D, L = 64) over both graphs, against the fixture's brute-force ground truth
oadj = np.zeros((N, R), np.uint32)
for i, row in enumerate(rows):
    oadj[i, :min(len(row), R)] = row[:R]
Q, gt = z["Q"], z["gt"]


def recall(a):
    hit = 0
    for qi in range(Q.shape[0]):
        r = O.search_heap(a, med, 64, vec=X, q=Q[qi], dist_mode=O.DIST_L2_SQRT, flavor=O.FLAVOR_DOUBLE, truncate_frontier=True)
        hit += len(set(r["ids"][:10].tolist()) & set(gt[qi].tolist()))
    return hit / (10 * Q.shape[0])


radj = z["adj16"].astype(np.uint32)
print(json.dumps({"config": f"{N}x{D} R={R} L={LB} alpha=1.2, permutations from random.seed(5)", "rows_identical_as_sets": int(same), "rows": N,
                  "mean_jaccard_of_rows": round(float(np.mean(jac)), 4), "oracle_build_seconds": round(dt, 1),
                  "mean_degree_oracle": round(float(np.mean([len(r) for r in rows])), 2),
                  "mean_degree_reference": round(float(np.mean([(len(set(r.tolist()) - {0})) for r in radj])), 2),
                  "recall_at_10_exact_L64": {"reference_graph": round(recall(radj), 4), "oracle_graph": round(recall(oadj), 4)}}))
