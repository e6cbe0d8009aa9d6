"""BASELINE configs[3]: GPU Vamana build (batched greedy search + RobustPrune alpha = 1.2, R = 64) on 1M x 768 synthetic
vectors; graph quality judged by recall@10 against a reference-built graph, as north_star asks (within 0.5 points).

The reference's builder is strictly sequential (cython_utils.pyx:269-369: ~hours at 1M), so the comparison graph is a
reference-EQUIVALENT one (oracle/oracle.c:orc_vamana_build, pinned row for row to the real builder in
tests/test_golden_oracle.py) over a 50k subsample, built once on the CPU by
    python tests/tools/build_config2_graph.py 50000 768 64 100 config4
and cached under .cache/.  On the GPU box this script
  (i)  builds the SAME 50k subset with dr_vamana_build (same R, L, alpha), searches both graphs with the same exact search
       (L = 100, k = 10, same queries) and reports both recalls;
  (ii) builds the full 1M x 768, R = 64 graph on the GPU: wall time, degree statistics, recall@10 against brute force.
Prints one JSON line (kept in profiles/)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))


def recall(ids, gt, k=10):
    return float(np.mean([len(set(ids[i][:k].tolist()) & set(gt[i][:k].tolist())) / k for i in range(len(gt))]))


def main():
    import torch
    from diskrag_b200 import ops
    from diskrag_b200._lib import check, lib
    from diskrag_b200.engine import GpuIndex
    from diskrag_b200.synth import synth_numpy, synth_torch
    out = {}
    cached = sorted(list((ROOT / ".cache").glob("config4_adj_*.npz")) + list((ROOT / "tests" / "golden").glob("config3_adj_*.npz")), key=lambda p: int(p.stem.split("_")[-1]))
    if cached:
        z = np.load(cached[-1])
        adj_ref, med, N, D, R, L, seed = z["adj"], int(z["medoid"]), int(z["N"]), int(z["D"]), int(z["R"]), int(z["L"]), int(z["seed"])
        X = synth_numpy(N, D, seed=seed)
        Q = synth_numpy(2000, D, seed=seed, sample_seed=1000)
        gt = np.argsort(-2.0 * Q @ X.T + (X * X).sum(1)[None, :], axis=1)[:, :10]
        t = time.time()
        adj_gpu, deg_gpu = ops.vamana_build(X, R, L, 1.2, med, seed=1)
        t_gpu = time.time() - t
        res = {}
        for name, adj in (("reference_equivalent", adj_ref), ("gpu", adj_gpu)):
            with GpuIndex.from_arrays(X, adj, medoid=med) as idx:
                res[name] = {Ls: recall(idx.search(Q, k=10, L=Ls, W=1, dist="exact", rerank=False).ids, gt) for Ls in (64, 100, 200)}
        out["subsample"] = {"N": N, "D": D, "R": R, "L_build": L, "alpha": 1.2,
                            "reference_equivalent_build_s_one_core": float(z["build_s"]), "gpu_build_s": round(t_gpu, 2),
                            "mean_degree": {"reference_equivalent": float(np.mean(z["deg"])), "gpu": float(deg_gpu.mean())},
                            "recall_at_10_by_search_L": res,
                            "max_recall_gap_points": round(100 * max(res["reference_equivalent"][l] - res["gpu"][l] for l in (64, 100, 200)), 3)}
    # ---- the full configuration on the device ----------------------------------------------------------------------
    dev = torch.device("cuda", 0)
    N, D, R, L = 1_000_000, 768, 64, 100
    X = synth_torch(N, D, seed=20243, device=dev)
    Q = synth_torch(1000, D, seed=20243, sample_seed=1000, device=dev)
    xn = (X * X).sum(1)
    prev = torch.backends.cuda.matmul.allow_tf32; torch.backends.cuda.matmul.allow_tf32 = False
    gt = torch.cat([(xn[None, :] - 2.0 * (Q[s:s + 250] @ X.T)).topk(10, largest=False).indices for s in range(0, 1000, 250)]).cpu().numpy()
    torch.backends.cuda.matmul.allow_tf32 = prev
    smp = torch.randperm(N, device=dev)[:256]
    sums = torch.zeros(256, dtype=torch.float64, device=dev)
    for c0 in range(0, N, 262144):
        d2 = (X[smp] * X[smp]).sum(1)[:, None] + xn[c0:c0 + 262144][None, :] - 2.0 * (X[smp] @ X[c0:c0 + 262144].T)
        sums += d2.clamp_min(0).sqrt().double().sum(1)
    med = int(smp[int(sums.argmin().item())].item())
    adj = torch.empty((N, R), dtype=torch.int32, device=dev); deg = torch.empty(N, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    torch.cuda.synchronize()
    t = time.time()
    check(lib().dr_vamana_build_dev(X.data_ptr(), N, D, R, L, 1.2, med, 1234, adj.data_ptr(), deg.data_ptr(), 0, st), "dr_vamana_build_dev")
    torch.cuda.synchronize()
    t_build = time.time() - t
    idx = GpuIndex.from_device_ptrs(X.data_ptr(), adj.data_ptr(), 0, 0, N, D, R, 0, med, 0, keepalive=(X, adj))
    Qh = Q.cpu().numpy()
    with idx:
        rec = {Ls: recall(idx.search(Qh, k=10, L=Ls, W=1, dist="exact", rerank=False).ids, gt) for Ls in (64, 100, 200)}
    d = deg.cpu().numpy()
    out["full"] = {"N": N, "D": D, "R": R, "L_build": L, "alpha": 1.2, "gpu_build_s": round(t_build, 2),
                   "degree": {"mean": float(d.mean()), "min": int(d.min()), "max": int(d.max())}, "recall_at_10_by_search_L": rec}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
