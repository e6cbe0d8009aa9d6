"""BASELINE.json configs[4]: a corpus too big for one GPU (default 8 x 1.25M x 3072 = 10M x 3072), index-sharded.
Every rank owns one shard with its OWN Vamana graph / medoid / PQ codes, searches ALL queries on it (throughput kernel,
fused exact rerank), then one NCCL all-to-all hands rank g the G partial top-k lists of its query slice and the k-way
merge kernel (dr_topk_merge_dev) reduces them (diskrag_b200/dist.py:index_sharded_topk).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/bench_index_sharded.py [--shard-n 1250000 --dim 3072 --queries 100000]
Rank 0 prints one JSON line (QPS over the whole sharded corpus, recall@10 against the GLOBAL exact ground truth)."""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shard-n", type=int, default=1_250_000)
    ap.add_argument("--dim", type=int, default=3072)
    ap.add_argument("--M", type=int, default=192)
    ap.add_argument("--R", type=int, default=32)
    ap.add_argument("--Lbuild", type=int, default=64)
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--W", type=int, default=8)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=100_000)
    ap.add_argument("--gt-queries", type=int, default=500)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from diskrag_b200 import dist as DD, engine
    from diskrag_b200._lib import check, lib
    from diskrag_b200.synth import synth_torch
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local); torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    N, D, M, R, B, k = a.shard_n, a.dim, a.M, a.R, a.queries, a.k
    t0 = time.time()
    X = synth_torch(N, D, seed=20245, sample_seed=rank, device=dev)            # this rank's shard of the corpus
    Q = synth_torch(B, D, seed=20245, sample_seed=100_000, device=dev)         # the same query batch on every rank
    # shard-local medoid (sampled, cython_utils.pyx:210-263), PQ and graph: all on this GPU
    g = torch.Generator(device=dev); g.manual_seed(77 + rank)
    smp = torch.randperm(N, generator=g, device=dev)[:256]
    xs = X[smp]; sums = torch.zeros(256, dtype=torch.float64, device=dev)
    xn = (X * X).sum(1)
    for c0 in range(0, N, 131072):
        d2 = (xs * xs).sum(1)[:, None] + xn[c0:c0 + 131072][None, :] - 2.0 * (xs @ X[c0:c0 + 131072].T)
        sums += d2.clamp_min(0).sqrt().double().sum(1)
    med = int(smp[int(sums.argmin().item())].item())
    cb = torch.empty((M, 256, D // M), dtype=torch.float32, device=dev)
    codes = torch.empty((N, M), dtype=torch.uint8, device=dev)
    mse = C.c_double(0)
    check(lib().dr_pq_train_dev(X.data_ptr(), N, D, M, 25, 42 + rank, cb.data_ptr(), C.byref(mse), local, st), "dr_pq_train_dev")
    check(lib().dr_pq_encode_dev(cb.data_ptr(), X.data_ptr(), N, D, M, codes.data_ptr(), local, st), "dr_pq_encode_dev")
    adj = torch.empty((N, R), dtype=torch.int32, device=dev); deg = torch.empty(N, dtype=torch.int32, device=dev)
    check(lib().dr_vamana_build_dev(X.data_ptr(), N, D, R, a.Lbuild, 1.2, med, 1234 + rank, adj.data_ptr(), deg.data_ptr(), local, st),
          "dr_vamana_build_dev")
    torch.cuda.synchronize(dev)
    setup_s = time.time() - t0
    idx = engine.GpuIndex.from_device_ptrs(X.data_ptr(), adj.data_ptr(), codes.data_ptr(), cb.data_ptr(), N, D, R, M, med, local,
                                           keepalive=(X, adj, codes, cb))
    p = engine.make_params(k=k, L=a.L, W=a.W, dist="pq", adc_order="tree", rerank=True, lut="u8tc" if (D // M) % 8 == 0 else "u8",
                           prefetch=5)
    ids = torch.empty((B, k), dtype=torch.int32, device=dev); dd = torch.empty((B, k), dtype=torch.float32, device=dev)
    stat = torch.empty(B, dtype=torch.int32, device=dev)
    offset = rank * N

    def step():
        idx.search_dev(Q.data_ptr(), B, p, ids.data_ptr(), dd.data_ptr(), d_status=stat.data_ptr(), stream=st)
        return DD.index_sharded_topk(ids, dd, offset, gather=False)            # this rank's slice of the global top-k

    gi, gd = step(); torch.cuda.synchronize(dev)
    assert int(stat.abs().sum().item()) == 0
    # global exact ground truth for the first gt-queries queries: per-shard exact top-k, then the same exchange
    ng = min(a.gt_queries, B)
    prev = torch.backends.cuda.matmul.allow_tf32; torch.backends.cuda.matmul.allow_tf32 = False
    ei = torch.empty((ng, k), dtype=torch.int32, device=dev); ed = torch.empty((ng, k), dtype=torch.float32, device=dev)
    for s in range(0, ng, 100):
        d = xn[None, :] - 2.0 * (Q[s:s + 100] @ X.T) + 1.0
        t = d.topk(k, largest=False)
        ei[s:s + 100] = t.indices.to(torch.int32); ed[s:s + 100] = t.values
    torch.backends.cuda.matmul.allow_tf32 = prev
    ti, _ = DD.index_sharded_topk(ei, ed, offset, gather=True)
    ai, _ = DD.index_sharded_topk(ids[:ng].contiguous(), dd[:ng].contiguous(), offset, gather=True)
    rec = float(np.mean([len(set(ai[i].tolist()) & set(ti[i].tolist())) / k for i in range(ng)]))
    for _ in range(a.warmup):
        step()
    dist.barrier(); torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "QPS@recall10>=0.95 (index-sharded)", "value": round(B * a.steps / (ms.item() / 1e3), 1),
                          "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": round(ms.item() / a.steps, 3), "scaling": "index-sharded: every GPU searches all queries on its shard",
                          "config": {"workload": f"{world} shards x {N} x {D} synthetic unit-norm (corpus {world * N}), per-shard Vamana R={R} + PQ M={M}, "
                                                 f"L={a.L}, W={a.W}, {B}-query batch, NCCL all-to-all of the per-shard top-{k} + k-way merge kernel",
                                     "recall_at_10_global": round(rec, 4), "recall_queries": ng, "setup_s_per_shard": round(setup_s, 1),
                                     "pq_mse": mse.value, "exchange_bytes_per_rank_per_step": B * k * 8}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
