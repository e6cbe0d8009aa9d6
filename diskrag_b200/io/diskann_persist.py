"""On-disk index layout of pydiskann/io (byte-compatible reader/writer).

Mirrors pydiskann/io/diskann_persist.py: `DiskANNPersist` (:11-206) and `MMapNodeReader` (:209-234).
Files:
  index.dat     N fixed records, no header: float32[D] vector || uint32[R] neighbour ids; rows shorter
                than R are padded with 0 (so node 0 is a phantom neighbour — kept verbatim), longer
                rows are truncated.                                         (save_index, :17-24)
  meta.json     json dict                                                   (save_meta, :26-28)
  pq_codes.bin  raw uint8[N, M]                                             (save_pq_codes, :30-31)
  pq_model.pkl  pickle(dict) with a list of sklearn KMeans objects, written to .tmp, re-read,
                renamed                                                     (save_pq_codebook, :33-105)
This module is host-side I/O only; the device index is created from these arrays by
diskrag_b200.engine.GpuIndex.from_dir / from_records.
"""
import json
import mmap
import os
import pickle
from collections import OrderedDict

import numpy as np


def codebook_of(pq_model) -> np.ndarray:
    """f32[M, 256, ds] centroid array of a DiskANNPQ-like object (ours or the reference's)."""
    return np.ascontiguousarray(np.stack([np.asarray(km.cluster_centers_, dtype=np.float32)
                                          for km in pq_model.kmeans_list]))


class DiskANNPersist:
    def __init__(self, dim=128, R=16, record_bytes=None):
        self.D = dim
        self.R = R
        self.record_size = 4 * (dim + R) if record_bytes is None else record_bytes

    # -- index.dat -----------------------------------------------------------------------------------
    def save_index(self, filepath, graph):
        """graph.nodes[i].vector / .neighbors for i in 0..N-1.  Array-backed graphs (ours) are written
        in one vectorised pass; any other object is walked node by node like the reference does."""
        if hasattr(graph, "to_records"):
            rec = graph.to_records(self.R)
            with open(filepath, "wb") as f:
                f.write(memoryview(rec))
            return
        n = len(graph.nodes)
        rec = np.zeros((n, self.D + self.R), dtype=np.uint32)
        for idx in range(n):
            node = graph.nodes[idx]
            rec[idx, :self.D] = np.asarray(node.vector, dtype=np.float32).view(np.uint32)
            nb = list(node.neighbors)[:self.R]
            rec[idx, self.D:self.D + len(nb)] = np.asarray(nb, dtype=np.uint32)
        with open(filepath, "wb") as f:
            f.write(memoryview(rec))

    def save_arrays(self, filepath, vec, adj):
        """vec f32[N,D], adj u32[N,R] (already 0-padded) -> index.dat."""
        vec = np.ascontiguousarray(vec, dtype=np.float32)
        adj = np.ascontiguousarray(adj, dtype=np.uint32)
        rec = np.empty((vec.shape[0], self.D + self.R), dtype=np.uint32)
        rec[:, :self.D] = vec.view(np.uint32)
        rec[:, self.D:] = adj[:, :self.R]
        with open(filepath, "wb") as f:
            f.write(memoryview(rec))

    def load_arrays(self, filepath):
        """index.dat -> (vec f32[N,D], adj u32[N,R]) copies."""
        raw = np.fromfile(filepath, dtype=np.uint32)
        if raw.size % (self.D + self.R):
            raise ValueError(f"{filepath}: size is not a multiple of the record size {self.record_size}")
        rec = raw.reshape(-1, self.D + self.R)
        return rec[:, :self.D].copy().view(np.float32), rec[:, self.D:].copy()

    # -- meta.json -------------------------------------------------------------------------------------
    def save_meta(self, filepath, meta_dict):
        with open(filepath, "w") as f:
            json.dump(meta_dict, f)

    def load_meta(self, filepath):
        with open(filepath, "r") as f:
            return json.load(f)

    # -- pq_codes.bin ------------------------------------------------------------------------------------
    def save_pq_codes(self, filepath, pq_codes):
        np.asarray(pq_codes).astype(np.uint8).tofile(filepath)

    def load_pq_codes(self, filepath, num_nodes, n_subvectors):
        return np.fromfile(filepath, dtype=np.uint8).reshape((num_nodes, n_subvectors))

    # -- pq_model.pkl ------------------------------------------------------------------------------------
    def save_pq_codebook(self, filepath, pq_model):
        if not getattr(pq_model, "is_fitted", False):
            raise ValueError("PQ 模型未訓練完成，無法保存")
        if not getattr(pq_model, "kmeans_list", None):
            raise ValueError("PQ 模型缺少 kmeans_list，無法保存")
        expected = (pq_model.n_centroids, pq_model.sub_dim)
        for i, km in enumerate(pq_model.kmeans_list):
            if not hasattr(km, "cluster_centers_"):
                raise ValueError(f"KMeans 模型 {i} 缺少 cluster_centers_")
            if km.cluster_centers_.shape != expected:
                raise ValueError(f"KMeans 模型 {i} 聚類中心形狀錯誤: {km.cluster_centers_.shape}, 預期: {expected}")
        model_data = {
            "n_subvectors": pq_model.n_subvectors,
            "n_centroids": pq_model.n_centroids,
            "sub_dim": pq_model.sub_dim,
            "is_fitted": pq_model.is_fitted,
            "kmeans_list": pq_model.kmeans_list,
            "means_": getattr(pq_model, "means_", None),
            "stds_": getattr(pq_model, "stds_", None),
            "epsilon": getattr(pq_model, "epsilon", 1e-8),
            "model_type": "DiskANNPQ",
            "version": "2.0",
        }
        tmp = str(filepath) + ".tmp"
        try:
            with open(tmp, "wb") as f:
                pickle.dump(model_data, f, protocol=pickle.HIGHEST_PROTOCOL)
            with open(tmp, "rb") as f:
                back = pickle.load(f)
            assert back["model_type"] == "DiskANNPQ"
            assert back["n_subvectors"] == model_data["n_subvectors"]
            assert len(back["kmeans_list"]) == len(model_data["kmeans_list"])
            os.replace(tmp, filepath)
        except Exception:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise

    def load_pq_codebook(self, filepath):
        with open(filepath, "rb") as f:
            data = pickle.load(f)
        if isinstance(data, dict) and "model_type" in data:
            return self._load_new_format_pq(data)
        return self._load_legacy_format_pq(data)

    def _load_new_format_pq(self, data):
        from ..pq.fast_pq import DiskANNPQ
        for key in ("n_subvectors", "n_centroids", "sub_dim", "is_fitted", "kmeans_list"):
            if key not in data:
                raise ValueError(f"PQ 模型數據缺少必要字段: {key}")
        pq = DiskANNPQ(n_subvectors=data["n_subvectors"], n_centroids=data["n_centroids"])
        pq.sub_dim = data["sub_dim"]
        pq.is_fitted = data["is_fitted"]
        pq.kmeans_list = data["kmeans_list"]
        pq.means_ = data.get("means_")
        pq.stds_ = data.get("stds_")
        pq.epsilon = data.get("epsilon", 1e-8)
        if not pq.kmeans_list:
            raise ValueError("加載的 PQ 模型缺少 kmeans_list")
        if len(pq.kmeans_list) != pq.n_subvectors:
            raise ValueError(f"KMeans 模型數量不匹配: {len(pq.kmeans_list)} != {pq.n_subvectors}")
        expected = (pq.n_centroids, pq.sub_dim)
        for i, km in enumerate(pq.kmeans_list):
            if not hasattr(km, "cluster_centers_"):
                raise ValueError(f"KMeans 模型 {i} 缺少 cluster_centers_")
            if km.cluster_centers_.shape != expected:
                raise ValueError(f"KMeans 模型 {i} 聚類中心形狀錯誤: {km.cluster_centers_.shape}, 預期: {expected}")
        pq._invalidate()
        return pq

    def _load_legacy_format_pq(self, data):
        if getattr(data, "is_fitted", False):
            if getattr(data, "kmeans_list", None):
                return data
            raise ValueError("舊格式 PQ 模型缺少 kmeans_list")
        raise ValueError("舊格式 PQ 模型未訓練或缺少 is_fitted 標記")


class MMapNodeReader:
    """Random access to index.dat records: get_node(id) -> (f32[D] vector, u32[R] neighbours)."""

    def __init__(self, filepath, dim=128, R=16, cache_size=1024):
        self.D = dim
        self.R = R
        self.record_size = 4 * (dim + R)
        self.filepath = str(filepath)
        self.file = open(filepath, "rb")
        self.mmap_obj = mmap.mmap(self.file.fileno(), 0, access=mmap.ACCESS_READ)
        self.cache = OrderedDict()
        self.cache_size = cache_size

    @property
    def num_nodes(self):
        return len(self.mmap_obj) // self.record_size

    def get_node(self, node_id):
        hit = self.cache.get(node_id)
        if hit is not None:
            self.cache.move_to_end(node_id)
            return hit
        off = int(node_id) * self.record_size
        vec = np.frombuffer(self.mmap_obj, dtype=np.float32, count=self.D, offset=off)
        nbr = np.frombuffer(self.mmap_obj, dtype=np.uint32, count=self.R, offset=off + 4 * self.D)
        if len(self.cache) >= self.cache_size:
            self.cache.popitem(last=False)
        self.cache[node_id] = (vec, nbr)
        return vec, nbr

    def records(self):
        """The whole file as a uint8 view (what GpuIndex.from_records uploads)."""
        return np.frombuffer(self.mmap_obj, dtype=np.uint8)

    def close(self):
        self.cache.clear()
        try:
            self.mmap_obj.close()
        except BufferError:
            pass  # numpy views still alive; the OS mapping goes away with them
        self.file.close()
