"""Batched array-level entry points over the C ABI (host numpy in/out unless noted)."""
import ctypes as C

import numpy as np

from ._lib import as_f32, check, lib, ptr


def _rowdist(fn, A, B, device):
    A = as_f32(np.atleast_2d(A)); B = as_f32(np.atleast_2d(B))
    n, D = A.shape
    if B.shape[1] != D or B.shape[0] not in (n, 1):
        raise ValueError(f"shape mismatch: A {A.shape} vs B {B.shape}")
    out = np.empty(n, np.float32)
    if n:
        check(fn(ptr(A), ptr(B), n, B.shape[0], D, ptr(out), device), fn.__name__)
    return out


def l2sq_batch(A, B, device=0):
    """Row-wise squared L2 (cython_utils.pyx:18-24, batched).  B may be a single row (broadcast)."""
    return _rowdist(lib().dr_l2sq_batch, A, B, device)


def dot_batch(A, B, device=0):
    return _rowdist(lib().dr_dot_batch, A, B, device)


def cosine_batch(A, B, device=0):
    """Row-wise cosine *distance* 1 - cos (cython_utils.pyx:53-70); 0 where a norm is 0."""
    return _rowdist(lib().dr_cosine_batch, A, B, device)


def pq_sdc_batch(codebook, c1, c2, device=0):
    codebook = as_f32(codebook)
    c1 = np.ascontiguousarray(np.atleast_2d(c1), np.uint8); c2 = np.ascontiguousarray(np.atleast_2d(c2), np.uint8)
    M, _, ds = codebook.shape
    out = np.empty(c1.shape[0], np.float32)
    check(lib().dr_pq_sdc_batch(ptr(codebook), ptr(c1), ptr(c2), c1.shape[0], M, ds, ptr(out), device), "dr_pq_sdc_batch")
    return out


def medoid(X, samples, device=0):
    X = as_f32(X); s = np.ascontiguousarray(samples, np.int32)
    out = C.c_int64(0)
    check(lib().dr_medoid(ptr(X), X.shape[0], X.shape[1], ptr(s), s.size, C.byref(out), device), "dr_medoid")
    return int(out.value)


def vamana_build(X, R, L, alpha, medoid_idx, seed=0, device=0):
    """-> (adj u32[N,R] 0-padded like index.dat, deg i32[N])"""
    X = as_f32(X)
    N, D = X.shape
    adj = np.empty((N, R), np.uint32); deg = np.empty(N, np.int32)
    check(lib().dr_vamana_build(ptr(X), N, D, R, L, float(alpha), int(medoid_idx), int(seed), ptr(adj), ptr(deg), device),
          "dr_vamana_build")
    return adj, deg


def topk_merge(ids_t, dist_t, stream=None):
    """torch CUDA tensors ids i32[G,B,k], dist f32[G,B,k] -> (i32[B,k], f32[B,k]) on the same device."""
    import torch
    G, B, k = ids_t.shape
    ids_t = ids_t.contiguous(); dist_t = dist_t.contiguous()
    oi = torch.empty((B, k), dtype=torch.int32, device=ids_t.device)
    od = torch.empty((B, k), dtype=torch.float32, device=ids_t.device)
    st = stream if stream is not None else torch.cuda.current_stream(ids_t.device).cuda_stream
    check(lib().dr_topk_merge_dev(ids_t.data_ptr(), dist_t.data_ptr(), G, B, k, oi.data_ptr(), od.data_ptr(),
                                  ids_t.device.index or 0, st), "dr_topk_merge_dev")
    return oi, od


def pq_lut_u8(codebook, Q, fmt="u8", device=0):
    """The throughput search's 8-bit ADC table (pq.cu / lut_tc.cu) in plain order:
    -> (table u8[B, M, 256], scale f32[B], offset f32[B]) with ADC^2 ~= offset + scale * sum_m table[b, m, code_m].
    fmt "u8" = CUDA-core build (restated bit-for-bit by oracle.c:orc_lut_u8), "u8tc" = tensor-core build (TF32 products)."""
    from . import _lib
    codebook = as_f32(codebook); Q = as_f32(np.atleast_2d(Q))
    M, _, ds = codebook.shape
    B, D = Q.shape
    if D != M * ds:
        raise ValueError(f"query dimension {D} != M * ds = {M * ds}")
    tab = np.empty((B, M, 256), np.uint8); sc = np.empty(B, np.float32); off = np.empty(B, np.float32)
    check(lib().dr_pq_lut_u8(ptr(codebook), ptr(Q), B, D, M, _lib.DR_LUT_U8_TC if fmt == "u8tc" else _lib.DR_LUT_U8,
                             ptr(tab), ptr(sc), ptr(off), device), "dr_pq_lut_u8")
    return tab, sc, off
