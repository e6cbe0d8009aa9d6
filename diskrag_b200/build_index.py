"""`diskrag index` on the GPU: the pydiskann half of scripts/tools/build_index.py:66-361, at the index-directory boundary.

The reference's build_index(collection_name, ...) does three things: (1) collection plumbing (vectors.npy lookup,
collection_info.json — preprocessing/collection.py, out of scope: metadata, needs polars), (2) parameter policy
(adaptive R / L / alpha, search L, PQ M), (3) the numeric path: DiskANNPQ.fit / encode, the build-time self-checks,
build_vamana, and the four index files.  This module is (2) + (3) with every numeric step on the GPU:

    meta = build_index_dir(vectors, "collections/<name>/index", target_quality="balanced")

writes index.dat / pq_codes.bin / pq_model.pkl / meta.json in the reference's byte layout, so
SearchEngineCorrect (search_engine.py:18-116), MMapNodeReader and diskrag_b200.search_engine.GpuSearchEngine all open
the result.  With the package swap of INTEGRATION.md §1 the reference's own build_index.py runs the same calls
unmodified; this function exists so that the path is usable (and tested) without the collection layer.
"""
import json
import math
import time
from datetime import datetime
from pathlib import Path

import numpy as np

from ._lib import DiskragError

from . import _lib
from .io.diskann_persist import DiskANNPersist
from .pq.fast_pq import DiskANNPQ
from .vamana_graph import build_vamana

SUBVECTOR_CANDIDATES = (4, 8, 16, 32, 48, 64, 96, 128)      # adaptive_pq.py:29


def adaptive_build_params(n_points: int, target_quality: str = "balanced") -> dict:
    """calculate_adaptive_build_params (scripts/tools/build_index.py:15-48): R / L by corpus size, scaled by the tier."""
    base_R, base_L = ((16, 32) if n_points <= 10_000 else (20, 48) if n_points <= 50_000 else
                      (24, 64) if n_points <= 200_000 else (28, 80))
    if target_quality == "fast":
        return {"R": int(base_R * 0.8), "L": int(base_L * 0.8), "alpha": 1.0, "target_recall": 0.7}
    if target_quality == "high":
        return {"R": int(base_R * 1.2), "L": int(base_L * 1.4), "alpha": 1.2, "target_recall": 0.95}
    return {"R": base_R, "L": base_L, "alpha": 1.2, "target_recall": 0.85}


def adaptive_search_L(n_points: int, target_recall: float = 0.85) -> int:
    """calculate_adaptive_search_L (scripts/tools/build_index.py:50-64)."""
    lg = math.log10(n_points)
    base = 10 * (8 + lg) if n_points <= 10_000 else 10 * (15 + 2 * lg) if n_points <= 100_000 else 10 * (20 + 3 * lg)
    if target_recall >= 0.9:
        base *= 2.0
    elif target_recall >= 0.85:
        base *= 1.5
    return max(20, min(int(base), n_points // 3))


def adaptive_pq_subvectors(n_points: int, dimension: int, target_quality: str = "balanced") -> int:
    """AdaptivePQCalculator.calculate_adaptive_pq_params (pydiskann/pq/adaptive_pq.py:42-150) reduced to its decision:
    candidates dividing D with sub-dimension in [2, 64]; the tier picks max / middle / third / min by corpus size.
    Returns 0 for "brute_force" (N < 1000)."""
    if n_points < 1000:
        return 0
    acc = {"fast": "space_saving", "balanced": "balanced", "high": "high_accuracy"}.get(target_quality, "balanced")
    cand = [m for m in SUBVECTOR_CANDIDATES if dimension % m == 0 and 2 <= dimension // m <= 64] or [8, 16, 32]
    if n_points <= 50_000:
        return max(cand) if acc == "high_accuracy" else cand[len(cand) // 2]
    if n_points <= 500_000:
        return min(cand) if acc == "space_saving" else cand[len(cand) // 2]
    if n_points <= 2_000_000:
        return cand[len(cand) // 3] if acc == "high_accuracy" else min(cand)
    return min(cand)


def build_index_dir(vectors, index_dir, target_quality: str = "balanced", force_rebuild: bool = False, R=None, L=None,
                    alpha=None, n_subvectors=None, device: int = 0, verbose: bool = False):
    """-> meta dict (what meta.json holds), or None when the directory already holds an index and force_rebuild is off
    (build_index.py:137-145)."""
    _lib.require_gpu()                                      # no CPU fallback: fail before touching the directory
    vectors = np.asarray(vectors)
    if vectors.dtype != np.float32:                         # :99-101 (embeddings often arrive float64)
        vectors = vectors.astype(np.float32)
    if vectors.ndim == 1:
        vectors = vectors.reshape(1, -1)
    elif vectors.ndim > 2:
        vectors = vectors.reshape(-1, vectors.shape[-1])
    vectors = np.ascontiguousarray(vectors)
    n_points, dimension = vectors.shape
    if n_points < 16:                                       # :123-125
        raise ValueError(f"向量數量({n_points})不足，至少需要 16 個向量才能建立索引（PQ 訓練需要）")
    index_dir = Path(index_dir)
    if not force_rebuild and index_dir.exists() and any(index_dir.iterdir()):
        return None
    index_dir.mkdir(parents=True, exist_ok=True)

    bp = adaptive_build_params(n_points, target_quality)
    R = bp["R"] if R is None else int(R)
    L = bp["L"] if L is None else int(L)
    alpha = bp["alpha"] if alpha is None else float(alpha)
    target_recall = bp["target_recall"]
    M = adaptive_pq_subvectors(n_points, dimension, target_quality) if n_subvectors is None else int(n_subvectors)
    use_pq = True
    if M == 0:                                              # "brute_force" recommendation (:176-179)
        use_pq, M = False, 8
    if use_pq and n_points < 256:                           # DiskANNPQ.fit needs >= 256 rows (fast_pq.py:212-213)
        use_pq = False
    search_L = adaptive_search_L(n_points, target_recall)
    persist = DiskANNPersist(dim=dimension, R=R)
    t0 = time.time()

    avg_error = selectivity = 0.0
    pq_model = None
    if use_pq:
        try:
            pq_model = DiskANNPQ(n_subvectors=M, n_centroids=256, device=device)
            pq_model.fit(vectors, show_progress=verbose)
            for i, km in enumerate(pq_model.kmeans_list):                      # :221-229
                if km.cluster_centers_.shape != (pq_model.n_centroids, pq_model.sub_dim):
                    raise ValueError(f"KMeans 模型 {i} 聚類中心形狀錯誤")
            pq_codes = pq_model.encode(vectors)
            test = vectors[:5]                                                  # :236-243 decode(encode) error on 5 vectors
            avg_error = float(np.mean(np.linalg.norm(test - pq_model.decode(pq_model.encode(test)), axis=1)))
            selectivity = float(pq_model.estimate_selectivity(vectors, sample_size=min(1000, n_points)))
            persist.save_pq_codebook(str(index_dir / "pq_model.pkl"), pq_model)
            loaded = persist.load_pq_codebook(str(index_dir / "pq_model.pkl"))  # :254-271 save -> load -> re-encode equality
            if not np.array_equal(pq_model.encode(test), loaded.encode(test)):
                raise ValueError("PQ 模型保存/加載驗證失敗，請檢查模型序列化問題")
            persist.save_pq_codes(str(index_dir / "pq_codes.bin"), pq_codes)
        except DiskragError:                                                    # a device failure (out of memory, lost GPU) is not
            raise                                                               # a PQ-quality problem: never hide it
        except Exception:                                                       # :277-282: PQ failure degrades to exact search
            use_pq, pq_model = False, None
    if not use_pq:                                                              # no stale / partial PQ files next to use_pq = false
        for name in ("pq_model.pkl", "pq_codes.bin"):
            (index_dir / name).unlink(missing_ok=True)
    t_pq = time.time() - t0

    graph = build_vamana(vectors, R=R, L=L, alpha=alpha, show_progress=verbose)  # :288 (called WITHOUT the pq model)
    medoid_idx = int(getattr(graph, "medoid_idx", 0))
    persist.save_index(str(index_dir / "index.dat"), graph)
    t_graph = time.time() - t0 - t_pq

    meta = {                                                                    # :299-332, same keys
        "D": int(dimension), "R": int(R), "L": int(L), "alpha": float(alpha), "N": int(n_points),
        "medoid_idx": medoid_idx, "n_subvectors": int(M) if use_pq else 0, "pq_centroids": 256 if use_pq else 0,
        "build_time": datetime.now().isoformat(), "recommended_search_L": int(search_L),
        "target_recall": float(target_recall), "target_quality": str(target_quality), "use_pq": bool(use_pq),
        "vector_stats": {"dtype": str(vectors.dtype), "shape": list(vectors.shape), "min": float(vectors.min()),
                         "max": float(vectors.max()), "mean": float(vectors.mean()), "std": float(vectors.std())},
        "build_seconds": {"pq": round(t_pq, 3), "graph": round(t_graph, 3)},    # additive: wall time on the GPU
    }
    if use_pq and pq_model is not None:
        meta["pq_validation"] = {"avg_reconstruction_error": avg_error, "selectivity": selectivity,
                                 "encoding_consistency_check": "PASSED", "distance_consistency_check": "PASSED"}
    persist.save_meta(str(index_dir / "meta.json"), meta)
    return json.loads(json.dumps(meta))
