"""BASELINE configs[1] input: a REFERENCE-EQUIVALENT Vamana graph over 100k x 1536 synthetic vectors, built on the CPU by
oracle/oracle.c:orc_vamana_build — the restatement of build_vamana_index_cython (cython_utils.pyx:269-492) that
tests/test_golden_oracle.py pins row for row against the real reference build at 600 points and
tests/tools/check_build_config0.py at 10k x 1536 (10000 / 10000 rows, with the compiled summation order of oracle.c:l2sq_refbuild;
graphs cached before that fix used plain sequential sums: same algorithm, statistically the same graph, not edge-identical to what the
reference binary would output).  Sequential by nature (~25 min on one core); the adjacency (N x R u32, 0-padded like DiskANNPersist.save_index) is cached under .cache/ and the vectors are
regenerated from the seed wherever it is used (tests/tools/parity_config2.py).
usage: python tests/tools/build_config2_graph.py [N [D R L tag]]     (tag names the cache file: config2 | config4)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np
import oracle as O
from diskrag_b200.synth import synth_numpy

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
D, R, L, ALPHA, SEED = 1536, 32, 64, 1.2, 20241
TAG = "config2"
if len(sys.argv) > 5:
    D, R, L, TAG = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    SEED = 20243
O.build()
X = synth_numpy(N, D, seed=SEED)
rng = np.random.default_rng(SEED)
s0 = rng.permutation(N).astype(np.int32); s1 = rng.permutation(N).astype(np.int32)
med = O.medoid(X, rng.choice(N, 1000, replace=False).astype(np.int32))
t = time.time()
rows = O.vamana_build(X, R, L, ALPHA, med, s0, s1)
adj = np.zeros((N, R), np.uint32)
for i, row in enumerate(rows):
    adj[i, :len(row)] = row[:R]
out = ROOT / ".cache" / f"{TAG}_adj_{N}.npz"
np.savez_compressed(out, adj=adj, medoid=np.int64(med), N=N, D=D, R=R, L=L, alpha=ALPHA, seed=SEED,
                    deg=np.array([len(r) for r in rows], np.int32), build_s=time.time() - t)
print(out, "built in", round(time.time() - t, 1), "s; mean degree", float(np.mean([len(r) for r in rows])))
