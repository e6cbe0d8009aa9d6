"""Device-resident index + batched search: the additive batched API of SURVEY §8(b).

`GpuIndex` owns an opaque `dr_index*` (vectors, adjacency, PQ codes, codebook in HBM).  All arrays
crossing this boundary are numpy arrays owned by the caller (nothing is retained) or, for the
`*_dev` methods, raw device pointers (ints) + a CUDA stream handle, e.g. torch `tensor.data_ptr()`.
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np

from . import _lib
from ._lib import SearchParams, as_f32, check, lib, ptr


class SearchResult:
    __slots__ = ("ids", "dists", "hops", "visited", "status", "list_ids", "list_dists", "list_len", "trace")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


def make_params(k=10, L=100, W=1, dist="pq", adc_order="seq", rerank=True, sqrt_out=False, hash_cap=0, chunk=0,
                threads=0, lut="f32", prefetch=False, start=None, ignore_deleted=False, w2=0) -> SearchParams:
    """start: entry point of the call (the reference's start_idx); None = the index's own (medoid)."""
    return SearchParams(k=k, L=L, W=W, dist={"pq": _lib.DR_DIST_PQ, "cosine": _lib.DR_DIST_COSINE}.get(dist, _lib.DR_DIST_EXACT),
                        adc_order=_lib.DR_ADC_TREE if adc_order == "tree" else _lib.DR_ADC_SEQ,
                        rerank=int(bool(rerank)), sqrt_out=int(bool(sqrt_out)), hash_cap=hash_cap, chunk=chunk,
                        threads=threads, lut_fmt={"u8": _lib.DR_LUT_U8, "u8tc": _lib.DR_LUT_U8_TC}.get(lut, _lib.DR_LUT_F32),
                        prefetch=int(prefetch), start_plus1=0 if start is None else int(start) + 1,
                        ignore_deleted=int(bool(ignore_deleted)), w_after_empty=int(w2))


class GpuIndex:
    """Vamana graph + PQ codes resident on one GPU."""

    def __init__(self, handle, N, D, R, M, medoid, device, keepalive=None):
        self._h = handle
        self.N, self.D, self.R, self.M, self.medoid, self.device = N, D, R, M, medoid, device
        self._keepalive = keepalive  # device tensors adopted by dr_index_create_dev

    # ---- constructors -------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, vec, adj, codes=None, codebook=None, medoid=0, device=0):
        """vec f32[N,D], adj u32[N,R] (0-padded like index.dat), codes u8[N,M], codebook f32[M,256,ds]."""
        _lib.require_gpu()
        vec = as_f32(vec)
        adj = np.ascontiguousarray(adj, dtype=np.uint32)
        N, D = vec.shape
        R = adj.shape[1]
        M = 0
        if codes is not None:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            M = codes.shape[1]
        if codebook is not None:
            codebook = as_f32(codebook)
            M = codebook.shape[0]
        h = C.c_void_p()
        check(lib().dr_index_create(ptr(vec), ptr(adj), ptr(codes), ptr(codebook), N, D, R, M, int(medoid), device,
                                    C.byref(h)), "dr_index_create")
        return cls(h, N, D, R, M, int(medoid), device)

    @classmethod
    def from_records(cls, records, N, D, R, codes=None, codebook=None, medoid=0, device=0):
        """records: the raw index.dat image (bytes / uint8 array / np.memmap), N*4*(D+R) bytes."""
        _lib.require_gpu()
        rec = np.ascontiguousarray(np.frombuffer(records, dtype=np.uint8) if not isinstance(records, np.ndarray) else records)
        if rec.nbytes != N * 4 * (D + R):
            raise ValueError(f"index.dat size {rec.nbytes} != N*4*(D+R) = {N * 4 * (D + R)}")
        M = 0
        if codes is not None:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            M = codes.shape[1]
        if codebook is not None:
            codebook = as_f32(codebook)
            M = codebook.shape[0]
        h = C.c_void_p()
        check(lib().dr_index_create_from_records(ptr(rec), N, D, R, ptr(codes), ptr(codebook), M, int(medoid), device,
                                                 C.byref(h)), "dr_index_create_from_records")
        return cls(h, N, D, R, M, int(medoid), device)

    @classmethod
    def from_dir(cls, index_dir, device=0):
        """Load index.dat / meta.json / pq_codes.bin / pq_model.pkl written by the reference or by us."""
        from .io.diskann_persist import DiskANNPersist, codebook_of
        d = Path(index_dir)
        meta = json.loads((d / "meta.json").read_text())
        N, D, R = int(meta["N"]), int(meta["D"]), int(meta["R"])
        rec = np.memmap(d / "index.dat", dtype=np.uint8, mode="r")
        codes = codebook = None
        # search_engine.py:36-70: PQ only when meta says so (default True) and both files are there; anything wrong with them
        # (stale files of another corpus, wrong size, unreadable pickle) means exact search, as in the reference
        M = int(meta.get("n_subvectors", 0) or 0)
        if meta.get("use_pq", True) and M > 0 and (d / "pq_codes.bin").exists() and (d / "pq_model.pkl").exists():
            p = DiskANNPersist(dim=D, R=R)
            try:
                if (d / "pq_codes.bin").stat().st_size != N * M:
                    raise ValueError(f"pq_codes.bin holds {(d / 'pq_codes.bin').stat().st_size} bytes, expected N*M = {N * M}")
                codes = p.load_pq_codes(d / "pq_codes.bin", N, M)
                codebook = codebook_of(p.load_pq_codebook(d / "pq_model.pkl"))
                if codebook.shape != (M, 256, D // M) or D % M:
                    raise ValueError(f"pq_model.pkl codebook {codebook.shape} does not match meta (M={M}, D={D})")
            except _lib.DiskragError:
                raise
            except Exception as e:
                print(f"⚠️  PQ 文件不可用 ({e})，切換到精確搜索模式")
                codes = codebook = None
        return cls.from_records(rec, N, D, R, codes, codebook, int(meta.get("medoid_idx", 0)), device)

    @classmethod
    def from_device_ptrs(cls, d_vec, d_adj, d_codes, d_codebook, N, D, R, M, medoid, device=0, keepalive=None):
        _lib.require_gpu()
        h = C.c_void_p()
        check(lib().dr_index_create_dev(d_vec, d_adj, d_codes or None, d_codebook or None, N, D, R, M, int(medoid),
                                        device, C.byref(h)), "dr_index_create_dev")
        return cls(h, N, D, R, M, int(medoid), device, keepalive)

    # ---- lifecycle ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib().dr_index_destroy(self._h)
            self._h = None
            self._keepalive = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- search ---------------------------------------------------------------------------------
    def search(self, Q, k=10, L=100, W=1, dist="pq", adc_order="seq", rerank=True, sqrt_out=False, lut=None,
               want_list=False, trace=0, hash_cap=0, chunk=0, threads=0, lut_fmt="f32", prefetch=False, start=None,
               ignore_deleted=False, w2=0) -> SearchResult:
        """Batched search of Q f32[B,D] (host).  Returns host numpy arrays."""
        Q = as_f32(np.atleast_2d(Q))
        B, D = Q.shape
        if D != self.D:
            raise ValueError(f"query dimension {D} != index dimension {self.D}")
        if dist == "pq" and self.M == 0:
            raise ValueError("index has no PQ codes; use dist='exact'")
        p = make_params(k, L, W, dist, adc_order, rerank, sqrt_out, hash_cap, chunk, threads, lut_fmt, prefetch, start, ignore_deleted, w2)
        if lut is not None:
            lut = as_f32(lut).reshape(B, self.M, 256)
        ids = np.empty((B, k), np.int32); dd = np.empty((B, k), np.float32)
        hops = np.empty(B, np.int32); vis = np.empty(B, np.int32); st = np.empty(B, np.int32)
        llen = np.empty(B, np.int32)
        lids = np.empty((B, L), np.int32) if want_list else None
        ldist = np.empty((B, L), np.float32) if want_list else None
        tr = np.empty((B, trace), np.int32) if trace else None
        check(lib().dr_search_batch(self._h, ptr(Q), B, C.byref(p), ptr(lut), ptr(ids), ptr(dd), ptr(hops), ptr(vis),
                                    ptr(lids), ptr(ldist), ptr(llen), ptr(tr), trace, ptr(st)), "dr_search_batch")
        if (st != 0).any():
            bad = int(np.flatnonzero(st)[0])
            raise _lib.DiskragError(f"search status {int(st[bad])} for query {bad} "
                                    "(1 = visited-set overflow, 2 = tie-ghost overflow)")
        return SearchResult(ids=ids, dists=dd, hops=hops, visited=vis, status=st, list_ids=lids, list_dists=ldist,
                            list_len=llen, trace=tr)

    def beam_search_c(self, Q, k=3, beam_width=5, dist="pq", sqrt_out=True, start=None) -> SearchResult:
        """Variant C with the reference's own semantics (vamana_graph.py:535-605; csrc/beam_c.cu): the k-capped beam whose
        truncation keeps the beam_width WORST frontier entries.  Results sorted by (dist, id); ids -1 padded."""
        Q = as_f32(np.atleast_2d(Q))
        B, D = Q.shape
        if D != self.D:
            raise ValueError(f"query dimension {D} != index dimension {self.D}")
        if dist not in ("pq", "exact"):
            raise ValueError("beam_search_c: dist must be 'pq' or 'exact'")
        if dist == "pq" and self.M == 0:
            raise ValueError("index has no PQ codes; use dist='exact'")
        ids = np.empty((B, k), np.int32); dd = np.empty((B, k), np.float32)
        hops = np.empty(B, np.int32); vis = np.empty(B, np.int32)
        check(lib().dr_beam_search_c(self._h, ptr(Q), B, int(k), int(beam_width),
                                     _lib.DR_DIST_PQ if dist == "pq" else _lib.DR_DIST_EXACT, int(bool(sqrt_out)),
                                     -1 if start is None else int(start), ptr(ids), ptr(dd), ptr(hops), ptr(vis)), "dr_beam_search_c")
        return SearchResult(ids=ids, dists=dd, hops=hops, visited=vis, status=np.zeros(B, np.int32), list_ids=None,
                            list_dists=None, list_len=None, trace=None)

    def search_host(self, Q, params: SearchParams, ids, dists=None, hops=None, visited=None, status=None):
        """Same call as `search` but into caller-owned host arrays (e.g. pinned buffers): Q f32[B,D] ->
        ids i32[B,k], dists f32[B,k], hops/visited/status i32[B].  H2D and D2H copies happen inside."""
        B = Q.shape[0]
        check(lib().dr_search_batch(self._h, ptr(Q), B, C.byref(params), None, ptr(ids), ptr(dists), ptr(hops), ptr(visited),
                                    None, None, None, None, 0, ptr(status)), "dr_search_batch")

    def search_dev(self, d_Q, B, params: SearchParams, d_ids, d_dist=None, d_hops=None, d_visited=None, d_lut=None,
                   d_list_ids=None, d_list_dist=None, d_list_len=None, d_status=None, stream=None):
        """Enqueue a search on device pointers (ints).  No synchronisation."""
        check(lib().dr_search_batch_dev(self._h, d_Q, B, C.byref(params), d_lut, d_ids, d_dist, d_hops, d_visited,
                                        d_list_ids, d_list_dist, d_list_len, None, 0, d_status, stream),
              "dr_search_batch_dev")

    # ---- dynamic updates: O(rows touched) (vamana_graph.py:58-125) ----------------------------------
    def append(self, vec, codes=None):
        """n new nodes with ids N .. N+n-1 and empty rows."""
        vec = as_f32(np.atleast_2d(vec))
        if codes is not None:
            codes = np.ascontiguousarray(np.atleast_2d(codes), dtype=np.uint8)
        check(lib().dr_index_append(self._h, ptr(vec), ptr(codes), vec.shape[0]), "dr_index_append")
        self.N += vec.shape[0]

    def patch_rows(self, rows, adj):
        """adj u32[n,R] replaces the adjacency rows `rows`; a slot >= N (0xFFFFFFFF) means "no neighbour"."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        adj = np.ascontiguousarray(adj, dtype=np.uint32).reshape(rows.size, self.R)
        check(lib().dr_index_patch_rows(self._h, ptr(rows), rows.size, ptr(adj)), "dr_index_patch_rows")

    def patch_vectors(self, rows, vec, codes=None):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        vec = as_f32(np.atleast_2d(vec))
        if codes is not None:
            codes = np.ascontiguousarray(np.atleast_2d(codes), dtype=np.uint8)
        check(lib().dr_index_patch_vectors(self._h, ptr(rows), rows.size, ptr(vec), ptr(codes)), "dr_index_patch_vectors")

    def set_deleted_rows(self, rows, flags):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        check(lib().dr_index_set_deleted_rows(self._h, ptr(rows), rows.size, ptr(flags)), "dr_index_set_deleted_rows")

    def lut(self, Q):
        Q = as_f32(np.atleast_2d(Q))
        out = np.empty((Q.shape[0], self.M, 256), np.float32)
        check(lib().dr_lut_build(self._h, ptr(Q), Q.shape[0], ptr(out)), "dr_lut_build")
        return out

    def export_records(self):
        out = np.empty(self.N * 4 * (self.D + self.R), np.uint8)
        check(lib().dr_index_export_records(self._h, ptr(out)), "dr_index_export_records")
        return out

    def kernel_timing(self, enable=True):
        ms = C.c_double(0); n = C.c_int64(0)
        check(lib().dr_search_kernel_timing(self._h, int(enable), C.byref(ms), C.byref(n)))
        return ms.value, n.value


def launch_count() -> int:
    return int(lib().dr_launch_count())
