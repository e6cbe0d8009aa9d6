"""Parameter policy of `diskrag index` (scripts/tools/build_index.py:15-64, pydiskann/pq/adaptive_pq.py:42-150) as restated in
diskrag_b200/build_index.py: pure host logic, no GPU."""
import pytest


def test_policy_tables_match_the_reference():
    from diskrag_b200.build_index import adaptive_build_params, adaptive_pq_subvectors, adaptive_search_L
    assert adaptive_build_params(10_000) == {"R": 16, "L": 32, "alpha": 1.2, "target_recall": 0.85}       # build_index.py:17-18
    assert adaptive_build_params(1_000_000, "high") == {"R": 33, "L": 112, "alpha": 1.2, "target_recall": 0.95}
    assert adaptive_build_params(30_000, "fast") == {"R": 16, "L": 38, "alpha": 1.0, "target_recall": 0.7}
    assert adaptive_pq_subvectors(10_000, 1536) == 64         # the adaptive default at D = 1536 (adaptive_pq.py:81-108)
    assert adaptive_pq_subvectors(500, 1536) == 0             # brute_force below 1000 points
    assert adaptive_search_L(10_000, 0.85) == 180 and adaptive_search_L(1_000_000, 0.95) == 760


def test_build_index_dir_fails_loudly_without_a_gpu(tmp_path):
    """no CPU fallback: without a CUDA device the index build raises before it touches the directory"""
    import numpy as np
    from diskrag_b200 import _lib
    from diskrag_b200.build_index import build_index_dir
    try:
        _lib.require_gpu()
    except Exception:
        with pytest.raises(Exception):
            build_index_dir(np.zeros((300, 16), np.float32), tmp_path / "idx")
        assert not (tmp_path / "idx").exists()
    else:
        pytest.skip("a GPU is present")
