"""`diskrag index` end to end on the GPU (diskrag_b200/build_index.py mirrors the pydiskann half of
scripts/tools/build_index.py:66-361).  The directory it writes must be what the reference's own readers expect:
file sizes (verify_disk_index.py:18-138), meta.json keys (build_index.py:299-332), pq_model.pkl loadable by the
REAL reference loader and re-encoding identically (:263-271), index.dat searchable by the REAL beam_search_from_disk."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REF_META_KEYS = {"D", "R", "L", "alpha", "N", "medoid_idx", "n_subvectors", "pq_centroids", "build_time", "recommended_search_L",
                 "target_recall", "target_quality", "use_pq", "vector_stats", "pq_validation"}


def _recall(ids, gt, k=10):
    return float(np.mean([len(set(ids[i][:k]) & set(gt[i][:k])) / k for i in range(len(gt))]))


def test_build_index_dir_writes_a_reference_compatible_index(tmp_path):
    from diskrag_b200.build_index import build_index_dir
    from diskrag_b200.search_engine import GpuSearchEngine
    from diskrag_b200.synth import synth_numpy
    N, D = 12_000, 256
    X = synth_numpy(N, D, seed=21).astype(np.float64)          # embeddings arrive float64 in the CLI path (diskrag.py:194)
    Q = synth_numpy(64, D, seed=21, sample_seed=5)
    d = tmp_path / "collections" / "demo" / "index"
    meta = build_index_dir(X, d, target_quality="balanced")
    assert set(meta) >= REF_META_KEYS and meta["use_pq"] and meta["pq_validation"]["encoding_consistency_check"] == "PASSED"
    R, M = meta["R"], meta["n_subvectors"]
    assert (meta["D"], meta["N"], R, meta["L"]) == (D, N, 20, 48) and M == 32
    assert json.loads((d / "meta.json").read_text())["medoid_idx"] == meta["medoid_idx"]
    assert (d / "index.dat").stat().st_size == N * 4 * (D + R)                  # verify_disk_index.py
    assert (d / "pq_codes.bin").stat().st_size == N * M
    assert build_index_dir(X, d) is None                                        # non-empty index dir: skipped (:137-145)
    X32 = X.astype(np.float32)
    gt = np.argsort(-2.0 * Q @ X32.T + (X32 * X32).sum(1)[None, :], axis=1)[:, :10]

    eng = GpuSearchEngine(d)
    try:
        r = eng.search_vectors(Q, k=10, L=meta["recommended_search_L"])
        assert _recall(r.ids, gt) >= meta["target_recall"]
        res, stats = eng._pq_accelerated_graph_search(Q[0], k=10, L=100)
        assert len(res) == 10 and {"search_time", "nodes_visited", "pq_distance_computations"} <= set(stats)
    finally:
        eng.close()

    # the REAL reference reads what we wrote (oracle/_ref = pydiskann compiled from /root/reference; test infrastructure)
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    ref = ref_loader.load()
    persist = ref["diskann_persist"].DiskANNPersist(dim=D, R=R)
    pq_ref = persist.load_pq_codebook(str(d / "pq_model.pkl"))                  # diskann_persist.py:107-199
    codes = np.fromfile(d / "pq_codes.bin", dtype=np.uint8).reshape(N, M)
    assert np.array_equal(pq_ref.encode(X32[:200]), codes[:200])                # sklearn predict == our encode on our codebook
    reader = ref["diskann_persist"].MMapNodeReader(str(d / "index.dat"), dim=D, R=R)
    ids = [[int(i) for _, i in ref["vamana_graph"].beam_search_from_disk(reader, Q[i], meta["medoid_idx"], beam_width=64, k=10)]
           for i in range(20)]
    reader.close()
    assert _recall(ids, gt[:20]) >= 0.9


def test_brute_force_tier_and_force_rebuild(tmp_path):
    from diskrag_b200.build_index import build_index_dir
    from diskrag_b200.synth import synth_numpy
    X = synth_numpy(600, 64, seed=2)
    d = tmp_path / "idx"
    meta = build_index_dir(X, d)
    assert meta["use_pq"] is False and meta["n_subvectors"] == 0 and not (d / "pq_codes.bin").exists()   # < 1000 points
    assert (d / "index.dat").stat().st_size == 600 * 4 * (64 + meta["R"])
    meta2 = build_index_dir(X, d, force_rebuild=True, n_subvectors=8)
    assert meta2["use_pq"] and (d / "pq_codes.bin").stat().st_size == 600 * 8
    with pytest.raises(ValueError):
        build_index_dir(X[:10], tmp_path / "tiny")


@pytest.mark.gpu
def test_index_dir_with_unusable_pq_files_falls_back_to_exact_search(tmp_path):
    """search_engine.py:36-70: use_pq = false in meta, a missing file, or PQ files that do not belong to this index (stale codes of
    another corpus, wrong size) all mean exact search — never a crash, never a traversal over foreign codes."""
    import json
    from diskrag_b200.build_index import build_index_dir
    from diskrag_b200.engine import GpuIndex
    from diskrag_b200.search_engine import GpuSearchEngine
    from diskrag_b200.synth import synth_numpy
    X = synth_numpy(600, 64, seed=5, K=16, r=8)
    d = tmp_path / "idx"
    meta = build_index_dir(X, d, target_quality="balanced", n_subvectors=8, verbose=False)
    assert meta["use_pq"] and (d / "pq_codes.bin").exists()
    with GpuIndex.from_dir(d) as idx:
        assert idx.M == 8
    # 1. stale codes of another (larger) corpus
    (d / "pq_codes.bin").write_bytes(b"\0" * (700 * 8))
    with GpuIndex.from_dir(d) as idx:
        assert idx.M == 0
    eng = GpuSearchEngine(d)
    assert not eng.use_pq and len(eng._exact_graph_search(X[3], k=5)[0]) == 5
    eng.close()
    # 2. meta says no PQ although files exist
    (d / "pq_codes.bin").write_bytes(b"\0" * (600 * 8))
    m = json.loads((d / "meta.json").read_text()); m["use_pq"] = False
    (d / "meta.json").write_text(json.dumps(m))
    with GpuIndex.from_dir(d) as idx:
        assert idx.M == 0
    # 3. a rebuild that ends without PQ (too few points) leaves no PQ files behind
    meta = build_index_dir(X[:200], d, target_quality="balanced", verbose=False, force_rebuild=True)
    assert not meta["use_pq"] and not (d / "pq_codes.bin").exists() and not (d / "pq_model.pkl").exists()
