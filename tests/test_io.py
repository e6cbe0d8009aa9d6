"""On-disk layout (pydiskann/io) round trips against the reference-written golden index.dat image.
Mirrors the reference's own layout test (test_disk_write_verify.py:74-83,150-176)."""
import json
import pickle

import numpy as np

from diskrag_b200.io.diskann_persist import DiskANNPersist, MMapNodeReader, codebook_of


def test_index_dat_roundtrip_is_byte_identical(golden, tmp_path):
    g = golden
    p = DiskANNPersist(dim=g["D"], R=g["R"])
    assert p.record_size == 4 * (g["D"] + g["R"])
    f = tmp_path / "index.dat"
    p.save_arrays(f, g["vec"], g["adj"])
    raw = np.fromfile(f, dtype=np.uint8)
    assert raw.size == g["N"] * 4 * (g["D"] + g["R"])          # file size == N*4*(D+R)
    assert np.array_equal(raw, g["records"])                    # byte-for-byte what the reference wrote
    vec, adj = p.load_arrays(f)
    assert np.array_equal(vec, g["vec"]) and np.array_equal(adj, g["adj"])


def test_save_index_from_node_objects_pads_with_zero(golden, tmp_path):
    g = golden

    class Node:
        def __init__(self, v, nb):
            self.vector, self.neighbors = v, nb

    class Graph:
        nodes = {0: Node(g["vec"][0], {3, 1}), 1: Node(g["vec"][1], set()), 2: Node(g["vec"][2], set(range(40)))}

    R = 4
    p = DiskANNPersist(dim=g["D"], R=R)
    f = tmp_path / "i.dat"
    p.save_index(f, Graph())
    vec, adj = p.load_arrays(f)
    assert np.array_equal(vec, g["vec"][:3])
    assert sorted(adj[0][:2]) == [1, 3] and list(adj[0][2:]) == [0, 0]   # short rows 0-padded
    assert list(adj[1]) == [0, 0, 0, 0]
    assert len(adj[2]) == R                                              # long rows truncated


def test_mmap_reader(golden, tmp_path):
    g = golden
    f = tmp_path / "index.dat"
    g["records"].tofile(f)
    r = MMapNodeReader(f, dim=g["D"], R=g["R"], cache_size=4)
    assert (r.D, r.R, r.record_size) == (g["D"], g["R"], 4 * (g["D"] + g["R"]))
    for i in (0, 7, g["N"] - 1, 7):
        v, nb = r.get_node(i)
        assert v.dtype == np.float32 and nb.dtype == np.uint32
        assert np.array_equal(v, g["vec"][i]) and np.array_equal(nb, g["adj"][i])
    for i in range(10):
        r.get_node(i)
    assert len(r.cache) <= 4
    assert r.num_nodes == g["N"]
    r.close()


def test_meta_and_codes_roundtrip(golden, tmp_path):
    g = golden
    p = DiskANNPersist(dim=g["D"], R=g["R"])
    p.save_pq_codes(tmp_path / "pq_codes.bin", g["codes"])
    assert np.array_equal(p.load_pq_codes(tmp_path / "pq_codes.bin", g["N"], g["M"]), g["codes"])
    meta = {"D": g["D"], "R": g["R"], "N": g["N"], "medoid_idx": g["medoid"], "n_subvectors": g["M"]}
    p.save_meta(tmp_path / "meta.json", meta)
    assert p.load_meta(tmp_path / "meta.json") == meta
    assert json.loads((tmp_path / "meta.json").read_text())["N"] == g["N"]


def test_pq_model_pickle_format(golden, tmp_path):
    """pq_model.pkl: a dict with real sklearn KMeans objects, readable by both sides."""
    from diskrag_b200.pq.fast_pq import DiskANNPQ, _wrap_kmeans
    g = golden
    pq = DiskANNPQ(g["M"])
    pq.sub_dim = g["D"] // g["M"]
    pq.kmeans_list = [_wrap_kmeans(g["codebook"][i], 42 + i) for i in range(g["M"])]
    pq.is_fitted = True
    p = DiskANNPersist(dim=g["D"], R=g["R"])
    p.save_pq_codebook(tmp_path / "pq_model.pkl", pq)
    raw = pickle.loads((tmp_path / "pq_model.pkl").read_bytes())
    assert raw["model_type"] == "DiskANNPQ" and raw["version"] == "2.0"
    assert set(raw) >= {"n_subvectors", "n_centroids", "sub_dim", "is_fitted", "kmeans_list", "means_", "stds_", "epsilon"}
    from sklearn.cluster import KMeans
    assert all(isinstance(k, KMeans) for k in raw["kmeans_list"])
    back = p.load_pq_codebook(tmp_path / "pq_model.pkl")
    assert np.array_equal(codebook_of(back), g["codebook"])
    # sklearn's own predict still works on the wrapped objects (search_engine.py:63-64, fast_pq.py:265)
    ds = g["D"] // g["M"]
    pred = back.kmeans_list[0].predict(g["vec"][:50, :ds])
    assert (pred == g["codes"][:50, 0]).mean() > 0.97
    import pytest
    with pytest.raises(ValueError):
        p.save_pq_codebook(tmp_path / "x.pkl", DiskANNPQ(4))   # unfitted
