"""Drop-in for the reference's only native module, pydiskann/cython_utils.pyx: the same eight Python-callable
names and argument meaning, each backed by a CUDA kernel through the C ABI (no CPU fallback).

  l2_distance_fast_cython            cython_utils.pyx:18-24    -> dr_l2sq_batch
  pq_distance_fast_cython            :26-51                    -> dr_pq_sdc_batch
  cosine_similarity_cython           :53-70                    -> dr_cosine_batch
  greedy_search_cython               :72-122                   -> dr_search_batch (W = 1: the reference's visit order)
  robust_prune_cython                :124-167                  -> dr_robust_prune
  generate_initial_neighbors_cython  :182-208                  (host; its result is dead work in the reference, §2a)
  compute_approximate_medoid_cython  :210-263                  -> dr_medoid
  build_vamana_index_cython          :269-369                  -> dr_vamana_build

A scalar call launches a kernel for one pair: it keeps the reference's callers working, but throughput lives in
the batched entry points (diskrag_b200.ops, GpuIndex.search).
"""
import numpy as np

from . import ops


def _f32_1d(a, name):
    # the Cython signature is np.ndarray[float32, ndim=1]: a wrong dtype raises ValueError there too
    if not isinstance(a, np.ndarray) or a.dtype != np.float32:
        got = getattr(a, "dtype", type(a).__name__)
        raise ValueError(f"Buffer dtype mismatch, expected 'DTYPE_FLOAT_t' but got '{got}' ({name})")
    if a.ndim != 1:
        raise ValueError(f"Buffer has wrong number of dimensions (expected 1, got {a.ndim})")
    return np.ascontiguousarray(a)


def _u8_1d(a, name):
    if not isinstance(a, np.ndarray) or a.dtype != np.uint8:
        got = getattr(a, "dtype", type(a).__name__)
        raise ValueError(f"Buffer dtype mismatch, expected 'DTYPE_UINT8_t' but got '{got}' ({name})")
    return np.ascontiguousarray(a)


def l2_distance_fast_cython(x, y):
    x, y = _f32_1d(x, "x"), _f32_1d(y, "y")
    return float(ops.l2sq_batch(x[None, :], y[None, :])[0])


def cosine_similarity_cython(x, y):
    x, y = _f32_1d(x, "x"), _f32_1d(y, "y")
    return float(ops.cosine_batch(x[None, :], y[None, :])[0])


def pq_distance_fast_cython(pq_model, code1, code2):
    if not pq_model.is_fitted:
        raise ValueError("PQ 模型未初始化")
    code1, code2 = _u8_1d(code1, "code1"), _u8_1d(code2, "code2")
    from .io.diskann_persist import codebook_of
    cb = pq_model.codebook() if hasattr(pq_model, "codebook") else codebook_of(pq_model)
    return float(ops.pq_sdc_batch(cb, code1[None, :], code2[None, :])[0])


def compute_approximate_medoid_cython(points_array, sample_size=1000):
    pts = np.asarray(points_array)
    if pts.dtype != np.float32 or pts.ndim != 2:
        raise ValueError(f"Buffer dtype mismatch, expected 'float' but got '{pts.dtype}'")
    n = pts.shape[0]
    if n <= sample_size:
        samples = np.arange(n, dtype=np.int32)
    else:
        samples = np.random.default_rng().choice(n, sample_size, replace=False).astype(np.int32)  # reference: time-seeded mt19937
    return ops.medoid(pts, samples)


def generate_initial_neighbors_cython(n_points, R):
    if R >= n_points:
        raise ValueError(f"R={R} must be < n_points={n_points} (the reference loops forever here)")
    rng = np.random.default_rng()
    out = np.empty((n_points, R), np.int32)
    for i in range(n_points):
        c = rng.choice(n_points - 1, R, replace=False)
        out[i] = np.where(c < i, c, c + 1)
    return out


def build_vamana_index_cython(points_array, R, L, alpha, medoid_idx, show_progress=False, seed=None):
    pts = np.asarray(points_array)
    if pts.dtype != np.float32 or pts.ndim != 2:
        raise ValueError(f"Buffer dtype mismatch, expected 'float' but got '{pts.dtype}'")
    if seed is None:
        import random
        seed = random.getrandbits(63)      # the reference draws its permutations from Python's `random` (:303-308)
    adj, deg = ops.vamana_build(pts, int(R), int(L), float(alpha), int(medoid_idx), seed)
    return [adj[i, :deg[i]].astype(np.int64).tolist() for i in range(pts.shape[0])]


def greedy_search_cython(graph, start_idx, query_vector, L, compute_query_distance_fn=None):
    from .vamana_graph import _graph_search
    q = _f32_1d(query_vector, "query_vector")
    return _graph_search(graph, int(start_idx), q, int(L))


def robust_prune_cython(graph, point_idx, candidate_set, alpha, R, compute_distance_fn=None):
    from .vamana_graph import _graph_prune
    _graph_prune(graph, int(point_idx), candidate_set, float(alpha), int(R))
