"""Multi-GPU partitioning of the search path.  One process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).  Nothing in the reference
corresponds to this module.

  query-sharded  : every rank holds the whole index; the batch is split into contiguous slices; there is no
                   collective on the data path, only an optional all-gather of the [B/G, k] results.
  index-sharded  : rank s owns rows [offset_s, offset_s + N_s) with its own graph / medoid / codes; every rank
                   searches ALL queries on its shard, then ONE exchange step: an all-to-all hands rank g the G
                   partial top-k lists of its query slice, which a k-way merge kernel (dr_topk_merge_dev)
                   reduces to the global top-k.  Payload is 8*B*k bytes per rank.
"""
import math

import numpy as np


def query_slice(B: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of a B-query batch owned by `rank` (the last ranks may get one query fewer)."""
    per, extra = divmod(B, world)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def shard_rows(N: int, rank: int, world: int):
    """Row range [lo, hi) of an N-row corpus owned by index shard `rank`."""
    return query_slice(N, rank, world)


def padded_slice_len(B: int, world: int) -> int:
    return math.ceil(B / world)


def exchange_partial_topk(ids, dists, group=None):
    """ids/dists: torch tensors [B, k] — this rank's shard-local top-k for ALL B queries (global ids, -1 / +inf pad).
    Returns ([G, Bq, k] ids, [G, Bq, k] dists, lo, hi) for the query slice this rank reduces; Bq = ceil(B/G) rows,
    rows past hi - lo are padding."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, k = ids.shape
    bq = padded_slice_len(B, world)
    send_i = torch.full((world, bq, k), -1, dtype=ids.dtype, device=ids.device)
    send_d = torch.full((world, bq, k), float("inf"), dtype=dists.dtype, device=dists.device)
    for g in range(world):
        lo, hi = query_slice(B, g, world)
        send_i[g, :hi - lo] = ids[lo:hi]
        send_d[g, :hi - lo] = dists[lo:hi]
    recv_i = torch.empty_like(send_i)
    recv_d = torch.empty_like(send_d)
    dist.all_to_all_single(recv_i.view(-1), send_i.view(-1), group=group)
    dist.all_to_all_single(recv_d.view(-1), send_d.view(-1), group=group)
    lo, hi = query_slice(B, rank, world)
    return recv_i, recv_d, lo, hi


def _merge_gpu(ids, dists):
    from . import ops
    return ops.topk_merge(ids, dists)


def index_sharded_topk(local_ids, local_dists, id_offset: int, group=None, gather=True, merge=None):
    """local_ids [B, k] are shard-LOCAL row numbers (-1 = empty); id_offset maps them to global ids.
    -> (ids [B, k], dists [B, k]) global top-k for every query when gather=True, else this rank's slice."""
    import torch
    import torch.distributed as dist
    merge = merge or _merge_gpu
    gids = torch.where(local_ids >= 0, local_ids + id_offset, local_ids)
    ri, rd, lo, hi = exchange_partial_topk(gids, local_dists, group)
    mi, md = merge(ri, rd)                                    # [Bq, k]
    if not gather:
        return mi[:hi - lo], md[:hi - lo]
    world = dist.get_world_size(group)
    out_i = [torch.empty_like(mi) for _ in range(world)]
    out_d = [torch.empty_like(md) for _ in range(world)]
    dist.all_gather(out_i, mi.contiguous(), group=group)
    dist.all_gather(out_d, md.contiguous(), group=group)
    B = local_ids.shape[0]
    parts_i, parts_d = [], []
    for g in range(world):
        glo, ghi = query_slice(B, g, world)
        parts_i.append(out_i[g][:ghi - glo]); parts_d.append(out_d[g][:ghi - glo])
    return torch.cat(parts_i), torch.cat(parts_d)


def gather_query_sharded(ids, dists, B: int, group=None):
    """All-gather the per-rank result slices of a query-sharded search back into [B, k] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bq = padded_slice_len(B, world)
    k = ids.shape[1]
    pad_i = torch.full((bq, k), -1, dtype=ids.dtype, device=ids.device); pad_i[:ids.shape[0]] = ids
    pad_d = torch.full((bq, k), float("inf"), dtype=dists.dtype, device=dists.device); pad_d[:dists.shape[0]] = dists
    out_i = [torch.empty_like(pad_i) for _ in range(world)]
    out_d = [torch.empty_like(pad_d) for _ in range(world)]
    dist.all_gather(out_i, pad_i, group=group)
    dist.all_gather(out_d, pad_d, group=group)
    pi, pd = [], []
    for g in range(world):
        lo, hi = query_slice(B, g, world)
        pi.append(out_i[g][:hi - lo]); pd.append(out_d[g][:hi - lo])
    return torch.cat(pi), torch.cat(pd)


def merge_topk_numpy(ids, dists):
    """Reference semantics of dr_topk_merge_dev for tests: (dist, id) ascending, empties last.  [G,B,k] -> [B,k]."""
    G, B, k = ids.shape
    oi = np.full((B, k), -1, ids.dtype); od = np.full((B, k), np.inf, dists.dtype)
    for b in range(B):
        fi = ids[:, b, :].ravel(); fd = dists[:, b, :].ravel()
        keep = fi >= 0
        o = np.lexsort((fi[keep], fd[keep]))[:k]
        oi[b, :len(o)] = fi[keep][o]; od[b, :len(o)] = fd[keep][o]
    return oi, od
