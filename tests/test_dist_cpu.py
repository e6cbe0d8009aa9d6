"""Host-side logic of the multi-GPU partitionings with world_size 2 over gloo on the CPU: slicing, the
all-to-all layout of the partial top-k lists, id offsets and the gather.  The per-shard searches are stood
in by the oracle's exact brute force (the real ones run on GPUs; tests/tools/multi_gpu_check.py covers NCCL)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_slices_cover_everything():
    from diskrag_b200.dist import query_slice
    for B in (0, 1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            sl = [query_slice(B, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == B
            assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmp):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
    import torch
    import torch.distributed as dist
    import oracle as O
    from diskrag_b200 import dist as D
    from diskrag_b200.synth import synth_numpy
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, Dm, B, k = 999, 16, 37, 5                      # ragged on purpose: neither divides by 2
    X = synth_numpy(N, Dm, seed=3, K=16, r=8); Q = synth_numpy(B, Dm, seed=3, sample_seed=1, K=16, r=8)
    gt = O.ground_truth(X, Q, k)
    merge = lambda keys: tuple(torch.from_numpy(a) for a in D.merge_keys_numpy(keys.numpy()))
    pack = lambda i, d, off, w: torch.from_numpy(D.pack_topk_numpy(i.numpy(), d.numpy(), off, w))
    # ---- index-sharded: every rank searches all queries on its rows, one exchange, k-way merge ----------
    lo, hi = D.shard_rows(N, rank, world)
    loc = O.ground_truth(X[lo:hi], Q, k)                                  # shard-local row numbers
    dd = np.stack([((X[lo:hi][loc[b]] - Q[b]) ** 2).sum(1) for b in range(B)]).astype(np.float32)
    ids, dists = D.index_sharded_topk(torch.from_numpy(loc.astype(np.int32)), torch.from_numpy(dd), lo, pack=pack, merge=merge)
    assert ids.shape == (B, k)
    for b in range(B):
        assert set(ids[b].tolist()) == set(gt[b].tolist()), (rank, b)
    assert bool((dists[:, 1:] >= dists[:, :-1]).all())
    # a shard that returns fewer than k hits (padding) must not break the merge
    loc2 = loc.copy().astype(np.int32); dd2 = dd.copy()
    if rank == 1:
        loc2[:, 2:] = -1; dd2[:, 2:] = np.inf
    ids2, _ = D.index_sharded_topk(torch.from_numpy(loc2), torch.from_numpy(dd2), lo, pack=pack, merge=merge)
    assert (ids2 >= 0).all()
    # ---- query-sharded: replicated index, each rank answers its slice, results all-gathered ---------------
    qlo, qhi = D.query_slice(B, rank, world)
    mine = torch.from_numpy(gt[qlo:qhi].astype(np.int32))
    md = torch.zeros((qhi - qlo, k), dtype=torch.float32)
    allids, _ = D.gather_query_sharded(mine, md, B)
    assert np.array_equal(allids.numpy(), gt.astype(np.int32))
    dist.destroy_process_group()
    Path(tmp, f"ok{rank}").write_text("ok")


def test_index_and_query_sharding_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
