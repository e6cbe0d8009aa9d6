"""The reference's recall / latency / QPS protocol (dataset_benchmark.py:75-176) on the GPU path — SURVEY §8(f4).

Same steps, same table: load train / test vectors (parquet with a 'vector' | 'emb' | 'embedding' column, or .npy), build the
Vamana graph, exact ground truth, then (a) in-memory `greedy_search` for L in {50, 100, 200} and (b) `beam_search_from_disk`
over the written index file for beam widths {24, 32, 48, 64}, reporting recall@k, mean latency and QPS = 1000 / ms.
Each row is measured twice: "per-query" = one shim call per query exactly as the reference's loop does (the latency a
single /search request sees), and "batched" = all test queries in one launch (what the GPU is for).

  python -m diskrag_b200.dataset_benchmark --train-file train.parquet --test-file test.parquet
  python -m diskrag_b200.dataset_benchmark --synthetic 100000 128          # no dataset at hand
"""
import argparse
import os
import tempfile
import time

import numpy as np

from . import ops
from .io.diskann_persist import DiskANNPersist, MMapNodeReader
from .vamana_graph import beam_search_from_disk, build_vamana, greedy_search


def load_vectors(filepath, max_points=None):
    """dataset_benchmark.py:27-60: parquet with a vector column (lists or arrays) -> float32 [n, D]; .npy accepted too."""
    if str(filepath).endswith(".npy"):
        v = np.load(filepath).astype(np.float32)
        if max_points and len(v) > max_points:
            v = v[np.random.RandomState(42).choice(len(v), max_points, replace=False)]
        return np.ascontiguousarray(v)
    import pandas as pd
    df = pd.read_parquet(filepath)
    if max_points and len(df) > max_points:
        df = df.sample(n=max_points, random_state=42).reset_index(drop=True)
    col = next((c for c in df.columns if c in ("vector", "emb", "embedding")), None)
    if col is None:
        col = next((c for c in df.columns if df[c].dtype == "object" and isinstance(df[c].iloc[0], (list, np.ndarray))), None)
    if col is None:
        raise ValueError(f"No vector column found in {filepath}")
    first = df[col].iloc[0]
    v = np.stack(df[col].tolist()) if isinstance(first, (list, np.ndarray)) else df[col].values
    return np.ascontiguousarray(v.astype(np.float32))


def compute_ground_truth(train_vecs, test_vecs, k=10, device=0):
    """dataset_benchmark.py:62-73 (np.linalg.norm + argsort per query), evaluated with the library's batched L2 kernel."""
    gt = np.empty((len(test_vecs), k), np.int64)
    for i, q in enumerate(test_vecs):
        d = ops.l2sq_batch(train_vecs, q[None, :], device=device)
        gt[i] = np.argsort(d, kind="stable")[:k]
    return gt


def _recall(pred, gt, k):
    return float(np.mean([len(set(list(pred[i])[:k]) & set(gt[i][:k].tolist())) / k for i in range(len(gt))]))


def run_benchmark(train_vecs, test_vecs, R=32, L=64, alpha=1.2, k=10, search_Ls=(50, 100, 200), beam_widths=(24, 32, 48, 64),
                  device=0, verbose=True):
    say = print if verbose else (lambda *a, **kw: None)
    t0 = time.time()
    graph = build_vamana(train_vecs, R=R, L=L, alpha=alpha)
    build_time = time.time() - t0
    avg_degree = float(np.mean(graph._deg)) if hasattr(graph, "_deg") else float(
        sum(len(n.neighbors) for n in graph.nodes.values()) / len(graph.nodes))
    say(f"Build complete in {build_time:.2f}s, avg degree {avg_degree:.2f}")
    gt = compute_ground_truth(train_vecs, test_vecs, k, device)
    start = int(getattr(graph, "medoid_idx", 0))
    out = {"build_time_s": build_time, "avg_degree": avg_degree, "in_memory": [], "disk": []}

    say(f"\nIn-memory search (k={k})\n{'L':<5} {'Recall':<10} {'Avg Time (ms)':<15} {'QPS':<10} {'batched QPS':<12}")
    for Ls in search_Ls:
        lat, res = [], []
        for q in test_vecs:                                   # the reference's loop: one call per query
            t = time.perf_counter()
            r = greedy_search(graph, start, q, Ls)
            lat.append((time.perf_counter() - t) * 1e3)
            res.append(r[:k])
        rec, ms = _recall(res, gt, k), float(np.mean(lat))
        idx = graph.gpu_index()                               # the device mirror greedy_search itself uses
        idx.search(test_vecs[:8], k=k, L=Ls, dist="exact", rerank=False)         # warm-up
        t = time.perf_counter()
        rb = idx.search(test_vecs, k=k, L=Ls, W=1, dist="exact", rerank=False)
        bq = len(test_vecs) / (time.perf_counter() - t)
        assert _recall(rb.ids, gt, k) == rec                  # the batch is the same search
        out["in_memory"].append({"L": Ls, "recall": rec, "avg_ms": ms, "qps": 1000 / ms, "batched_qps": bq})
        say(f"{Ls:<5} {rec:<10.4f} {ms:<15.2f} {1000 / ms:<10.0f} {bq:<12.0f}")

    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "vamana_index.bin")
        DiskANNPersist(dim=train_vecs.shape[1], R=R).save_index(path, graph)
        reader = MMapNodeReader(path, dim=train_vecs.shape[1], R=R)
        say(f"\nDisk-layout search\n{'Beam':<5} {'Recall':<10} {'Avg Time (ms)':<15} {'QPS':<10}")
        for bw in beam_widths:
            lat, res = [], []
            for q in test_vecs:
                t = time.perf_counter()
                r = beam_search_from_disk(reader, q, start, beam_width=bw, k=k)
                lat.append((time.perf_counter() - t) * 1e3)
                res.append([i for _, i in r][:k])
            rec, ms = _recall(res, gt, k), float(np.mean(lat))
            out["disk"].append({"beam": bw, "recall": rec, "avg_ms": ms, "qps": 1000 / ms})
            say(f"{bw:<5} {rec:<10.4f} {ms:<15.2f} {1000 / ms:<10.0f}")
        reader.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train-file", default="dataset/sift_small_500k/train_fixed.parquet")
    ap.add_argument("--test-file", default="dataset/sift_small_500k/test_fixed.parquet")
    ap.add_argument("--synthetic", nargs=2, type=int, metavar=("N", "D"), help="structured synthetic corpus instead of files")
    ap.add_argument("--max-train-points", type=int, default=None)
    ap.add_argument("--max-test-points", type=int, default=1000)
    ap.add_argument("--R", type=int, default=32)
    ap.add_argument("--L", type=int, default=64)
    ap.add_argument("--alpha", type=float, default=1.2)
    ap.add_argument("--search-L", type=int, default=None)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    if a.synthetic:
        from .synth import synth_numpy
        train = synth_numpy(a.synthetic[0], a.synthetic[1], seed=20240)
        test = synth_numpy(a.max_test_points, a.synthetic[1], seed=20240, sample_seed=1000)
    else:
        train = load_vectors(a.train_file, a.max_train_points)
        test = load_vectors(a.test_file, a.max_test_points)
    run_benchmark(train, test, a.R, a.L, a.alpha, a.k, [a.search_L] if a.search_L else (50, 100, 200))


if __name__ == "__main__":
    main()
