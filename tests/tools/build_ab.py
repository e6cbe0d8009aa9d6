"""Device A/B of graph-build variants (DISKRAG_B200_LIB selects the library): builds N x D synthetic points with
dr_vamana_build, prints seconds, mean degree and recall@10 of an exact search (L = 64 / 100) against brute force."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from diskrag_b200._lib import check, lib
    from diskrag_b200.engine import GpuIndex
    from diskrag_b200.synth import synth_torch
    N, D, R, L = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1_000_000, 768, 64, 100)))
    dev = torch.device("cuda", 0)
    X = synth_torch(N, D, seed=20243, device=dev)
    Q = synth_torch(1000, D, seed=20243, sample_seed=1000, device=dev)
    prev = torch.backends.cuda.matmul.allow_tf32; torch.backends.cuda.matmul.allow_tf32 = False
    gt = torch.topk((X * X).sum(1)[None, :] - 2.0 * (Q @ X.T), 10, dim=1, largest=False).indices.cpu().numpy()
    torch.backends.cuda.matmul.allow_tf32 = prev
    adj = torch.empty((N, R), dtype=torch.int32, device=dev); deg = torch.empty(N, dtype=torch.int32, device=dev)
    times = []
    import os
    profile_only = bool(os.environ.get("BUILD_AB_PROFILE"))     # one build, no recall leg (under ncu)
    for rep in range(1 if profile_only else 2):
        torch.cuda.synchronize(); t = time.time()
        check(lib().dr_vamana_build_dev(X.data_ptr(), N, D, R, L, 1.2, 0, 1, adj.data_ptr(), deg.data_ptr(), 0, 0), "dr_vamana_build_dev")
        torch.cuda.synchronize(); times.append(round(time.time() - t, 3))
    out = {"N": N, "D": D, "R": R, "L": L, "build_s": times, "mean_degree": round(float(deg.float().mean()), 2),
           "truncated": int(lib().dr_vamana_build_last_truncated())}
    if profile_only:
        print("BUILD_AB", json.dumps(out)); return
    Xh = X.cpu().numpy(); Qh = Q.cpu().numpy(); adjh = adj.cpu().numpy().view(np.uint32)
    del X
    with GpuIndex.from_arrays(Xh, adjh, medoid=0) as idx:
        for Ls in (64, 100):
            ids = idx.search(Qh, k=10, L=Ls, W=1, dist="exact", rerank=False).ids
            out[f"recall_at_10_L{Ls}"] = float(np.mean([len(set(ids[i, :10].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(gt))]))
    print("BUILD_AB", json.dumps(out))


if __name__ == "__main__":
    main()
