"""Experiment (oracle on the bench index, CPU of the GPU box): how many steps of the W = 8 throughput search end without a survivor,
and what expanding more entries after such a step would change (steps, hops, rows evaluated, recall)."""
import ctypes as C, sys, argparse
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np, torch
import bench, oracle as O
a = argparse.Namespace(n=1_000_000, dim=1536, R=32, M=192, Lbuild=64)
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
X, adj, deg, codes, cb, med, info = bench.build_index(a, dev)
from diskrag_b200.synth import synth_torch
nq = 200
Q = synth_torch(nq, a.dim, seed=20242, sample_seed=1000, device=dev)
gt = bench.ground_truth(X, Q, 10)
Xh = X.cpu().numpy(); adjh = adj.cpu().numpy().view(np.uint32); ch = codes.cpu().numpy(); cbh = cb.cpu().numpy(); Qh = Q.cpu().numpy()
O.build()
for w2 in (0, 16, 24, 32):
    O.lib().orc_set_w_after_empty(C.c_int(w2))
    t = C.c_int(0); e = C.c_int(0); O.lib().orc_step_stats(C.byref(t), C.byref(e), C.c_int(1))
    hops = vis = rec = 0
    for qi in range(nq):
        t8, _, _ = O.lut_u8(cbh, Qh[qi])
        l = O.search_list(adjh, med, 100, codes=ch, lut_=t8, dist_mode=O.DIST_ADC_U8, W=8, strict_ties=False)
        oi, _ = O.rerank(Xh, Qh[qi], l["ids"], 10, flavor=O.FLAVOR_WARP)
        hops += l["hops"]; vis += l["visited"]; rec += len(set(oi.tolist()) & set(gt[qi].tolist()))
    O.lib().orc_step_stats(C.byref(t), C.byref(e), C.c_int(1))
    print(f"w_after_empty {w2}: steps {t.value / nq:.2f} (empty {e.value / nq:.2f}) hops {hops / nq:.1f} visited {vis / nq:.1f} recall {rec / nq / 10:.4f}", flush=True)
