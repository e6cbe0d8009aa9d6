"""the reference's dataset_benchmark.py protocol (dataset_benchmark.py:75-176) on the GPU path."""
from pathlib import Path

import numpy as np
import pytest


def test_load_vectors_reads_the_reference_parquet_shapes(tmp_path):
    """the loader accepts what dataset_benchmark.py:27-60 accepts: a 'vector' / 'emb' / 'embedding' column of lists, or any
    object column of arrays; sub-sampling is seeded (random_state=42)"""
    import pandas as pd
    from diskrag_b200.dataset_benchmark import load_vectors
    v = np.random.default_rng(0).standard_normal((50, 8)).astype(np.float32)
    pd.DataFrame({"id": range(50), "emb": [r.tolist() for r in v]}).to_parquet(tmp_path / "a.parquet")
    pd.DataFrame({"id": range(50), "payload": [r for r in v]}).to_parquet(tmp_path / "b.parquet")
    np.save(tmp_path / "c.npy", v.astype(np.float64))
    assert np.array_equal(load_vectors(tmp_path / "a.parquet"), v)
    assert np.array_equal(load_vectors(tmp_path / "b.parquet"), v)
    assert np.array_equal(load_vectors(str(tmp_path / "c.npy")), v)
    s1, s2 = load_vectors(tmp_path / "a.parquet", 10), load_vectors(tmp_path / "a.parquet", 10)
    assert s1.shape == (10, 8) and np.array_equal(s1, s2)
    pd.DataFrame({"id": range(3)}).to_parquet(tmp_path / "d.parquet")
    with pytest.raises(ValueError):
        load_vectors(tmp_path / "d.parquet")


@pytest.mark.gpu
def test_benchmark_table_on_sift_like_vectors():
    from diskrag_b200.dataset_benchmark import run_benchmark
    from diskrag_b200.synth import synth_numpy
    train = synth_numpy(20000, 128, seed=4)               # SIFT-small is 128-d
    test = synth_numpy(100, 128, seed=4, sample_seed=77)
    out = run_benchmark(train, test, R=32, L=64, alpha=1.2, k=10, search_Ls=(50, 100), beam_widths=(24, 64), verbose=False)
    rec = [r["recall"] for r in out["in_memory"]]
    assert rec[0] >= 0.9 and rec[1] >= rec[0] - 1e-9       # recall grows with L (dataset_benchmark.py:102-130)
    assert out["disk"][1]["recall"] >= out["disk"][0]["recall"] - 1e-9 and out["disk"][1]["recall"] >= 0.9
    # beam_search_from_disk with beam 64 is greedy_search with L = 64 on the written file: between L = 50 and 100
    assert rec[0] - 0.02 <= out["disk"][1]["recall"] <= rec[1] + 0.02
    assert all(r["batched_qps"] > r["qps"] for r in out["in_memory"])
    assert 16 <= out["avg_degree"] <= 32


def test_dimension_gate_admits_3072_and_the_patch_applies(tmp_path):
    """f4: text-embedding-3-large collections.  The drop-in gate accepts what the reference accepts plus 3072, every admitted
    dimension gets a usable PQ sub-vector count, and the shipped patch turns the reference's own gate (when the tree is here)."""
    import importlib.util
    import subprocess
    from diskrag_b200 import config_gate as gate
    assert all(gate.validate_vector_dimension(d) for d in (128, 256, 768, 960, 1536, 3072)) and not gate.validate_vector_dimension(100)
    for d in sorted(gate.SUPPORTED_DIMENSIONS):
        m = gate.pq_subvectors_for(d)
        assert m > 0 and d % m == 0
    src = Path("/root/reference/preprocessing/config.py")
    if not src.exists():
        return
    work = tmp_path / "preprocessing"
    work.mkdir()
    (work / "config.py").write_text(src.read_text())
    patch = Path(__file__).resolve().parents[1] / "integration" / "0001-accept-3072-dimensional-collections.patch"
    subprocess.check_call(["patch", "-p1", "-s", "-d", str(tmp_path), "-i", str(patch)])
    text = (work / "config.py").read_text()
    ns = {}
    exec(compile("SUPPORTED" + text.split("SUPPORTED", 1)[1].split("def get_text_hash", 1)[0], "patched_config", "exec"), ns)
    assert ns["validate_vector_dimension"](3072) and ns["validate_vector_dimension"](1536) and not ns["validate_vector_dimension"](100)


@pytest.mark.gpu
def test_parquet_protocol_against_the_reference_driver():
    """f4: the reference's own benchmark driver and this package on the same parquet files (SIFT-like, reference column layout):
    same table shape, GPU recall within 2 points of the reference's on every row (its graph is built by another algorithm)."""
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent / "tools"))
    import dataset_benchmark_ab as ab
    out = ab.run(3000, 100, R=16, L=32)
    g = out["gpu"]
    assert [r["L"] for r in g["in_memory"]] == [50, 100, 200] and [r["beam"] for r in g["disk"]] == [24, 32, 48, 64]
    if out["reference"] is not None:
        r = out["reference"]
        assert [x["param"] for x in r["in_memory"]] == [50, 100, 200] and [x["param"] for x in r["disk"]] == [24, 32, 48, 64]
        for a, b in zip(g["in_memory"] + g["disk"], r["in_memory"] + r["disk"]):
            assert a["recall"] >= b["recall"] - 0.02, (a, b)
