"""the reference's dataset benchmark (dataset_benchmark.py:75-176: recall@k / mean latency / QPS of the in-memory
`greedy_search` for L in {50,100,200} and of `beam_search_from_disk` for beam widths {24,32,48,64}) run on the SAME parquet files by
  (A) the REAL reference — its own run_benchmark(args), from oracle/_ref — on the box's CPU, and
  (B) this package (diskrag_b200.dataset_benchmark, same protocol, per-query and batched) on the GPU.
No public dataset can be fetched here (SIFT-small parquet is not shipped with the reference either), so the files are written in
the reference's layout ('id', 'emb' list column) from a SIFT-like generator: 128-d, non-negative, integer-valued descriptors drawn
around cluster prototypes.  One JSON line: both tables side by side.

  python tests/tools/dataset_benchmark_ab.py [n_train] [n_test]"""
import contextlib
import io
import json
import re
import sys
import tempfile
import time
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))


def sift_like(n, seed, D=128, K=512):
    """SIFT-style descriptors: non-negative integers in [0, 218], heavy mass near 0, clustered."""
    rng = np.random.default_rng(1234)
    protos = rng.gamma(0.6, 28.0, size=(K, D))
    r = np.random.default_rng(seed)
    c = r.integers(0, K, n)
    x = protos[c] + r.normal(0, 9.0, size=(n, D)) + r.gamma(0.5, 6.0, size=(n, D))
    return np.clip(np.rint(x), 0, 218).astype(np.float32)


def write_parquet(path, X):
    import pandas as pd
    pd.DataFrame({"id": np.arange(len(X)), "emb": [row for row in X]}).to_parquet(path)


def parse_reference_tables(text):
    mem, disk, mode = [], [], None
    for line in text.splitlines():
        if "In-Memory" in line:
            mode = mem
        elif "Disk-Based" in line:
            mode = disk
        m = re.match(r"^(\d+)\s+([0-9.]+)\s+([0-9.]+)\s+(\d+)\s*$", line.strip())
        if m and mode is not None:
            mode.append({"param": int(m.group(1)), "recall": float(m.group(2)), "avg_ms": float(m.group(3)), "qps": float(m.group(4))})
    b = re.search(r"Build Complete in ([0-9.]+)s", text)
    return mem, disk, (float(b.group(1)) if b else None)


def run(n_train=20000, n_test=200, R=32, L=64, alpha=1.2, k=10):
    import random
    import ref_loader
    from diskrag_b200 import dataset_benchmark as ours
    out = {"dataset": f"SIFT-like synthetic parquet ('id','emb'), {n_train} x 128 train / {n_test} test, R={R} L={L} alpha={alpha} k={k}"}
    with tempfile.TemporaryDirectory() as tmp:
        tr, te = Path(tmp) / "train_fixed.parquet", Path(tmp) / "test_fixed.parquet"
        write_parquet(tr, sift_like(n_train, 1)); write_parquet(te, sift_like(n_test, 2))
        # (B) this package, through the parquet loader
        train = ours.load_vectors(tr); test = ours.load_vectors(te, n_test)
        assert train.shape == (n_train, 128) and train.dtype == np.float32
        t0 = time.time()
        res = ours.run_benchmark(train, test, R=R, L=L, alpha=alpha, k=k, verbose=False)
        out["gpu"] = {"build_time_s": res["build_time_s"], "avg_degree": res["avg_degree"], "in_memory": res["in_memory"], "disk": res["disk"],
                      "wall_s": time.time() - t0}
        # (A) the reference's own driver on the same files
        ref = ref_loader.load_dataset_benchmark() if ref_loader.available() else None
        if ref is None:
            out["reference"] = None
        else:
            args = types.SimpleNamespace(train_file=str(tr), test_file=str(te), max_train_points=None, max_test_points=n_test, R=R, L=L,
                                         alpha=alpha, search_L=None, k=k)
            random.seed(11)
            buf = io.StringIO()
            t0 = time.time()
            cwd = Path.cwd()
            import os
            os.chdir(tmp)                                    # it writes vamana_index.bin into the working directory
            try:
                with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
                    ref.run_benchmark(args)
            finally:
                os.chdir(cwd)
            mem, disk, build_s = parse_reference_tables(buf.getvalue())
            out["reference"] = {"build_time_s": build_s, "in_memory": mem, "disk": disk, "wall_s": time.time() - t0,
                                "what": "the reference's own dataset_benchmark.run_benchmark (oracle/_ref), one process"}
    return out


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 20000, int(sys.argv[2]) if len(sys.argv) > 2 else 200)))
