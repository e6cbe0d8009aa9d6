"""Multi-GPU partitioning of the search path.  One process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).  Nothing in the reference
corresponds to this module.

  query-sharded  : every rank holds the whole index; the batch is split into contiguous slices; there is no
                   collective on the data path, only an optional all-gather of the [B/G, k] results.
  index-sharded  : rank s owns rows [offset_s, offset_s + N_s) with its own graph / medoid / codes; every rank
                   searches ALL queries on its shard, then ONE exchange step of packed 64-bit (distance, global id) keys:
                   either one NCCL all-to-all of a send buffer packed on the device (index_sharded_topk), or no
                   collective at all — the search kernel's epilogue stores every query's top-k straight into the
                   reducing rank's receive buffer over NVLink peer memory (PeerExchange) — then the k-way merge kernel
                   (dr_topk_merge_keys_dev).  Payload is 8*B*k bytes per rank.
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


KEY_EMPTY = -1          # 0xFFFFFFFFFFFFFFFF as int64: the packed key of an empty slot


def _pack_gpu(ids, dists, id_offset, world):
    """[B,k] shard-local lists -> packed send buffer int64[G, Bq, k] (csrc/distance.cu:topk_pack_kernel), built on the device."""
    import torch
    from ._lib import check, lib
    B, k = ids.shape
    bq = padded_slice_len(B, world)
    out = torch.empty((world, bq, k), dtype=torch.int64, device=ids.device)
    st = torch.cuda.current_stream(ids.device).cuda_stream
    check(lib().dr_topk_pack_dev(ids.contiguous().data_ptr(), dists.contiguous().data_ptr(), B, k, int(id_offset), world, out.data_ptr(),
                                 ids.device.index or 0, st), "dr_topk_pack_dev")
    return out


def _merge_keys_gpu(keys):
    """received int64[G, Bq, k] -> (i32[Bq,k], f32[Bq,k]) (csrc/distance.cu:topk_merge_keys_kernel)"""
    import torch
    from ._lib import check, lib
    G, bq, k = keys.shape
    oi = torch.empty((bq, k), dtype=torch.int32, device=keys.device)
    od = torch.empty((bq, k), dtype=torch.float32, device=keys.device)
    st = torch.cuda.current_stream(keys.device).cuda_stream
    check(lib().dr_topk_merge_keys_dev(keys.data_ptr(), G, bq, k, oi.data_ptr(), od.data_ptr(), keys.device.index or 0, st),
          "dr_topk_merge_keys_dev")
    return oi, od


def index_sharded_topk(local_ids, local_dists, id_offset: int, group=None, gather=True, pack=None, merge=None):
    """local_ids [B, k] are shard-LOCAL row numbers (-1 = empty); id_offset maps them to global ids.
    ONE exchange: the lists are packed on the device into 64-bit keys (order-preserving distance bits << 32 | global id) laid out
    [G, Bq, k] by reducing rank, one all_to_all_single moves ids and distances together, the k-way merge kernel reduces the G lists
    of every query of this rank's slice.  -> (ids [B, k], dists [B, k]) for every query when gather=True (one all_gather of the
    merged keys' two halves), else this rank's slice.  `pack` / `merge` replace the device kernels in the CPU tests of this host
    logic (gloo, world_size 2)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, k = local_ids.shape
    send = (pack or _pack_gpu)(local_ids, local_dists, id_offset, world)           # int64 [G, Bq, k]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    mi, md = (merge or _merge_keys_gpu)(recv)                                      # [Bq, k]
    lo, hi = query_slice(B, rank, world)
    if not gather:
        return mi[:hi - lo], md[:hi - lo]
    both = torch.stack([mi.contiguous().view(torch.int32), md.contiguous().view(torch.int32)])     # one gather for both halves
    out = [torch.empty_like(both) for _ in range(world)]
    dist.all_gather(out, both, group=group)
    parts_i, parts_d = [], []
    for g in range(world):
        glo, ghi = query_slice(B, g, world)
        parts_i.append(out[g][0][:ghi - glo]); parts_d.append(out[g][1][:ghi - glo].view(torch.float32))
    return torch.cat(parts_i), torch.cat(parts_d)


class PeerExchange:
    """The exchange fused into the search kernel: every rank owns a receive buffer u64[G, Bq, k]; with the route set on
    its index (dr_index_set_peer_route) the throughput kernel's epilogue writes each query's top-k as packed keys straight into
    the receive buffer of the rank that reduces the query — stores to peer memory over NVLink / NVSwitch, no collective on the data
    path.  What is left per step: a barrier (every rank's kernel has finished, so every buffer is complete) and the merge kernel.
    Buffers are plain cudaMalloc allocations shared through CUDA IPC handles (one node)."""

    def __init__(self, index, B, k, id_offset, device, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._lib import check, lib
        self.group, self.index, self.B, self.k, self.device = group, index, int(B), int(k), int(device)
        self.world = dist.get_world_size(group); self.rank = dist.get_rank(group)
        self.bq = padded_slice_len(self.B, self.world)
        self.bytes = self.world * self.bq * self.k * 8
        self.recv = C.c_void_p()
        handle = C.create_string_buffer(64)
        check(lib().dr_dev_alloc(self.device, self.bytes, C.byref(self.recv), handle), "dr_dev_alloc")
        check(lib().dr_dev_memset(self.device, self.recv, 0xFF, self.bytes, None), "dr_dev_memset")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.peers = []
        ptrs = np.zeros(self.world, np.uint64)
        for g in range(self.world):
            if g == self.rank:
                ptrs[g] = self.recv.value
                self.peers.append(None)
            else:
                p = C.c_void_p()
                check(lib().dr_ipc_open(self.device, C.c_char_p(handles[g]), C.byref(p)), "dr_ipc_open")
                ptrs[g] = p.value
                self.peers.append(p)
        self.table = C.c_void_p()
        check(lib().dr_dev_alloc(self.device, 8 * self.world, C.byref(self.table), None), "dr_dev_alloc")
        check(lib().dr_dev_upload(self.device, self.table, ptrs.ctypes.data_as(C.c_void_p), 8 * self.world), "dr_dev_upload")
        check(lib().dr_index_set_peer_route(index._h, self.table, self.world, self.rank, self.B, int(id_offset)),
              "dr_index_set_peer_route")
        self.out_ids = torch.empty((self.bq, self.k), dtype=torch.int32, device=torch.device("cuda", self.device))
        self.out_dist = torch.empty((self.bq, self.k), dtype=torch.float32, device=torch.device("cuda", self.device))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def merge(self):
        """Call after this rank's search has been enqueued: waits until every rank's kernel is done (their stores into this
        rank's buffer are then complete), merges.  -> this rank's slice (ids [n,k], dists [n,k])."""
        import torch
        import torch.distributed as dist
        from ._lib import check, lib
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group=self.group)
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(lib().dr_topk_merge_keys_dev(self.recv, self.world, self.bq, self.k, self.out_ids.data_ptr(), self.out_dist.data_ptr(),
                                           self.device, st), "dr_topk_merge_keys_dev")
        lo, hi = query_slice(self.B, self.rank, self.world)
        return self.out_ids[:hi - lo], self.out_dist[:hi - lo]

    def close(self):
        from ._lib import lib
        import torch.distributed as dist
        if self.index is not None:
            lib().dr_index_set_peer_route(self.index._h, None, 0, 0, 0, 0)
            self.index = None
            import torch
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)              # nobody writes into a buffer that is about to be freed
            for p in self.peers:
                if p is not None:
                    lib().dr_ipc_close(self.device, p)
            lib().dr_dev_free(self.device, self.table)
            lib().dr_dev_free(self.device, self.recv)


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


def pack_topk_numpy(ids, dists, id_offset, world):
    """Reference semantics of dr_topk_pack_dev for the CPU tests: numpy [B,k] -> int64 [G, Bq, k]."""
    B, k = ids.shape
    bq = padded_slice_len(B, world)
    out = np.full((world, bq, k), -1, np.int64)
    f = np.ascontiguousarray(dists + np.float32(0.0), np.float32).view(np.uint32).astype(np.uint64)
    o = np.where(f & np.uint64(0x80000000), ~f & np.uint64(0xFFFFFFFF), f | np.uint64(0x80000000))       # order-preserving bits
    key = (o << np.uint64(32)) | (ids.astype(np.int64) + id_offset).astype(np.uint64)
    key = np.where(ids >= 0, key, np.uint64(0xFFFFFFFFFFFFFFFF)).view(np.int64)
    for g in range(world):
        lo, hi = query_slice(B, g, world)
        out[g, :hi - lo] = key[lo:hi]
    return out


def merge_keys_numpy(keys):
    """Reference semantics of dr_topk_merge_keys_dev: int64 [G, Bq, k] -> (i32 [Bq,k], f32 [Bq,k])."""
    G, bq, k = keys.shape
    u = keys.view(np.uint64)
    oi = np.full((bq, k), -1, np.int32); od = np.full((bq, k), np.inf, np.float32)
    for b in range(bq):
        flat = np.sort(u[:, b, :].ravel(), kind="stable")[:k]
        live = flat != np.uint64(0xFFFFFFFFFFFFFFFF)
        ob = (flat >> np.uint64(32)).astype(np.uint32)
        bits = np.where(ob & np.uint32(0x80000000), ob & np.uint32(0x7FFFFFFF), ~ob)
        oi[b, live] = (flat[live] & np.uint64(0xFFFFFFFF)).astype(np.int64).astype(np.int32)
        od[b, live] = bits.view(np.float32)[live]
    return oi, od


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
