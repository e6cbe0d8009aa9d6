"""DiskANNPQ on the GPU — same constructor, attributes, methods and error behaviour as
pydiskann/pq/fast_pq.py:162-352 (the live v3 class; `FastPQ` alias as at the end of that file).

  fit      -> dr_pq_train  (Lloyd k-means per subspace on the device; replaces M x sklearn KMeans,
              fast_pq.py:197-243).  The codebook is wrapped in real sklearn KMeans objects so that
              `kmeans_list[i].cluster_centers_ / .n_clusters / .predict` and the pq_model.pkl
              format keep working.
  encode   -> dr_pq_encode (nearest centroid, lowest index on ties; fast_pq.py:245-267)
  decode   -> dr_pq_decode (fast_pq.py:269-292)
  compute_distance_table -> dr_pq_lut, bit-identical to the numpy expression at fast_pq.py:294-318
  asymmetric_distance_sq -> dr_adc, the sequential fp32 sum of fast_pq.py:320-328
No CPU fallback: every method raises if the CUDA library or a device is missing.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .._lib import as_f32, check, lib, ptr


def _wrap_kmeans(centers: np.ndarray, seed: int):
    """A fitted-looking sklearn KMeans around GPU-trained centroids (predict() works)."""
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=centers.shape[0], n_init=1, max_iter=1, random_state=seed, algorithm="lloyd")
    km.cluster_centers_ = np.ascontiguousarray(centers, dtype=np.float32)
    km.n_features_in_ = centers.shape[1]
    km._n_features_out = centers.shape[0]
    km._n_threads = 1
    km.labels_ = np.zeros(0, dtype=np.int32)  # the reference pickles N labels per subspace: dead weight
    km.inertia_ = 0.0
    km.n_iter_ = 0
    return km


class DiskANNPQ:
    def __init__(self, n_subvectors: int = 8, n_centroids: int = 256, device: int = 0):
        if n_centroids != 256:
            print(f"⚠️  警告: n_centroids 已從 {n_centroids} 調整為 256 (最佳實踐)")
            n_centroids = 256
        self.n_subvectors = n_subvectors
        self.n_centroids = n_centroids
        self.kmeans_list = []
        self.sub_dim = 0
        self.is_fitted = False
        self.device = device
        self.train_iters = 25
        self.train_mse_ = None
        self._cb = None

    # -- helpers -------------------------------------------------------------------------------------
    @classmethod
    def from_codebook(cls, codebook, device: int = 0):
        """Additive: a fitted model around an existing codebook f32[M, 256, ds] (e.g. one read from pq_model.pkl)."""
        cb = np.ascontiguousarray(codebook, dtype=np.float32)
        if cb.ndim != 3 or cb.shape[1] != 256:
            raise ValueError(f"codebook must be f32[M, 256, ds], got {cb.shape}")
        pq = cls(cb.shape[0], 256, device)
        pq.sub_dim = cb.shape[2]
        pq.kmeans_list = [_wrap_kmeans(cb[i], 42 + i) for i in range(cb.shape[0])]
        pq._cb = cb
        pq.is_fitted = True
        return pq

    def _invalidate(self):
        self._cb = None

    def _dim(self) -> int:
        return self.sub_dim * self.n_subvectors

    def _check_vectors(self, X, what):
        # the C entry points read 256 * D * 4 bytes of codebook for the D they are told: never pass a caller-derived D that the
        # model was not trained for (the reference fails in sklearn / numpy broadcasting on the same inputs)
        if X.ndim != 2 or X.shape[1] != self._dim():
            raise ValueError(f"{what}: expected vectors of dimension {self._dim()} (= {self.n_subvectors} x {self.sub_dim}), got {X.shape}")

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_cb"] = None
        return st

    def codebook(self) -> np.ndarray:
        """f32[M, 256, ds]"""
        if self._cb is None:
            self._cb = np.ascontiguousarray(
                np.stack([np.asarray(km.cluster_centers_, dtype=np.float32) for km in self.kmeans_list]))
        return self._cb

    def _get_adaptive_kmeans_params(self, n_samples: int) -> dict:
        # fast_pq.py:188-195 (n_init / max_iter of sklearn; kept for callers that read them)
        if n_samples < 10000:
            return {"n_init": 10, "max_iter": 300}
        elif n_samples < 100000:
            return {"n_init": 5, "max_iter": 300}
        return {"n_init": 3, "max_iter": 200}

    # -- API -----------------------------------------------------------------------------------------
    def fit(self, vectors: np.ndarray, show_progress: bool = False) -> None:
        n_vectors, d = vectors.shape
        if d % self.n_subvectors != 0:
            raise ValueError(f"向量維度 {d} 必須能被子向量數量 {self.n_subvectors} 整除")
        self.sub_dim = d // self.n_subvectors
        if n_vectors < self.n_centroids:
            raise ValueError(f"訓練數據量 {n_vectors} 不足，至少需要 {self.n_centroids} 個向量")
        _lib.require_gpu()
        X = as_f32(vectors)
        cb = np.empty((self.n_subvectors, 256, self.sub_dim), np.float32)
        mse = C.c_double(0.0)
        check(lib().dr_pq_train(ptr(X), n_vectors, d, self.n_subvectors, self.train_iters, 42, ptr(cb), C.byref(mse),
                                self.device), "dr_pq_train")
        self.train_mse_ = mse.value
        self.kmeans_list = [_wrap_kmeans(cb[i], 42 + i) for i in range(self.n_subvectors)]
        self._cb = cb
        self.is_fitted = True

    def encode(self, vectors: np.ndarray) -> np.ndarray:
        if not self.is_fitted:
            raise ValueError("模型尚未訓練，請先調用 fit() 方法")
        X = as_f32(np.atleast_2d(vectors))
        self._check_vectors(X, "encode")
        n, d = X.shape
        codes = np.empty((n, self.n_subvectors), np.uint8)
        check(lib().dr_pq_encode(ptr(self.codebook()), ptr(X), n, d, self.n_subvectors, ptr(codes), self.device),
              "dr_pq_encode")
        return codes

    def decode(self, codes: np.ndarray) -> np.ndarray:
        if not self.is_fitted:
            raise ValueError("模型尚未訓練")
        codes = np.ascontiguousarray(np.atleast_2d(codes), dtype=np.uint8)
        if codes.shape[1] != self.n_subvectors:
            raise ValueError(f"decode: expected codes u8[n, {self.n_subvectors}], got {codes.shape}")
        n = codes.shape[0]
        d = self.sub_dim * self.n_subvectors
        out = np.empty((n, d), np.float32)
        check(lib().dr_pq_decode(ptr(self.codebook()), ptr(codes), n, d, self.n_subvectors, ptr(out), self.device),
              "dr_pq_decode")
        return out

    def compute_distance_table(self, query_vector: np.ndarray) -> np.ndarray:
        if not self.is_fitted:
            raise ValueError("模型尚未訓練")
        q = as_f32(query_vector).reshape(1, -1)
        self._check_vectors(q, "compute_distance_table")
        d = self.sub_dim * self.n_subvectors
        out = np.empty((self.n_subvectors, self.n_centroids), np.float32)
        check(lib().dr_pq_lut(ptr(self.codebook()), ptr(q), 1, d, self.n_subvectors, ptr(out), self.device), "dr_pq_lut")
        return out

    def compute_distance_tables(self, queries: np.ndarray) -> np.ndarray:
        """Batched variant (additive): f32[B, D] -> f32[B, M, 256]."""
        if not self.is_fitted:
            raise ValueError("模型尚未訓練")
        Q = as_f32(np.atleast_2d(queries))
        self._check_vectors(Q, "compute_distance_tables")
        out = np.empty((Q.shape[0], self.n_subvectors, self.n_centroids), np.float32)
        check(lib().dr_pq_lut(ptr(self.codebook()), ptr(Q), Q.shape[0], self._dim(), self.n_subvectors, ptr(out),
                              self.device), "dr_pq_lut")
        return out

    def asymmetric_distance_sq(self, codes: np.ndarray, distance_table: np.ndarray) -> np.ndarray:
        codes = np.ascontiguousarray(np.atleast_2d(codes), dtype=np.uint8)
        T = as_f32(distance_table)
        if codes.shape[1] != self.n_subvectors or T.shape != (self.n_subvectors, 256):
            raise ValueError(f"asymmetric_distance_sq: expected codes u8[n, {self.n_subvectors}] and a table f32[{self.n_subvectors}, 256], "
                             f"got {codes.shape} and {T.shape}")
        out = np.empty(codes.shape[0], np.float32)
        check(lib().dr_adc(ptr(codes), ptr(T), codes.shape[0], self.n_subvectors, ptr(out), self.device), "dr_adc")
        return out

    def asymmetric_distance(self, codes: np.ndarray, distance_table: np.ndarray) -> np.ndarray:
        return np.sqrt(self.asymmetric_distance_sq(codes, distance_table))

    def estimate_selectivity(self, vectors: np.ndarray, sample_size: int = 1000) -> float:
        sample_size = min(sample_size, len(vectors))
        idx = np.random.choice(len(vectors), sample_size, replace=False)
        codes = self.encode(vectors[idx])
        total = sum(len(np.unique(codes[:, i])) for i in range(self.n_subvectors))
        return total / (self.n_subvectors * self.n_centroids)

    def get_memory_usage(self) -> dict:
        """Called by build_vamana_with_pq(show_progress=True) (vamana_graph.py:530); in the reference it
        only exists on the dead first class, so that call raises AttributeError there."""
        d = self.sub_dim * self.n_subvectors
        return {"codebook_bytes": 256 * d * 4, "bytes_per_vector": self.n_subvectors,
                "compression_ratio": (d * 4) / max(1, self.n_subvectors)}


FastPQ = DiskANNPQ
