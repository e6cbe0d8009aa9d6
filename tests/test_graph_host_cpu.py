"""Host side of the VamanaGraphWithPQ drop-in (vamana_graph.py:8-56 Node / graph objects): the array-backed graph with a
lazy `nodes` view must behave like the reference's dict of Node objects for everything callers do WITHOUT a distance
computation — add_node / add_edge / delete_node, reading and mutating `nodes[i].neighbors`, and writing index.dat rows
(DiskANNPersist.save_index, diskann_persist.py:17-24: first R neighbours, 0-padded).  No GPU involved."""
import numpy as np
import pytest

from diskrag_b200.io.diskann_persist import DiskANNPersist, MMapNodeReader
from diskrag_b200.vamana_graph import Node, VamanaGraph, VamanaGraphWithPQ


def small_graph(N=12, D=6, R=4, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype(np.float32)
    g = VamanaGraphWithPQ(R)
    for i in range(N):
        g.add_node(i, X[i])
    return g, X


def test_nodes_view_behaves_like_the_reference_dict():
    g, X = small_graph()
    assert len(g.nodes) == 12 and 11 in g.nodes and 12 not in g.nodes and "x" not in g.nodes
    assert list(g.nodes) == list(range(12)) and list(g.nodes.keys()) == list(range(12))
    assert [n.idx for n in g.nodes.values()] == list(range(12))
    assert np.array_equal(g.nodes[5].vector, X[5]) and g.nodes[5].neighbors == set() and not g.nodes[5].is_deleted
    with pytest.raises(KeyError):
        g.nodes[12]
    with pytest.raises(KeyError):                                  # ids are dense: the next id is 12
        g.nodes[20] = Node(20, X[0])
    g.add_edge(1, 2); g.add_edge(1, 1); g.add_edge(1, 99)          # self loops and unknown ids are ignored (vamana_graph.py:46-49)
    assert g.nodes[1].neighbors == {2}
    g.nodes[1].neighbors.update({3, 4})                            # callers mutate the set in place
    g.nodes[3].neighbors = {0}                                     # or replace it (robust_prune does)
    rec = g.to_records()
    D, R = 6, 4
    assert rec.shape == (12, D + R) and rec.dtype == np.uint32
    assert np.array_equal(rec[:, :D].view(np.float32), X)
    assert set(rec[1, D:D + 3].tolist()) == {2, 3, 4} and rec[1, D + 3] == 0      # 0-padding like save_index
    assert rec[3, D:].tolist() == [0, 0, 0, 0] or rec[3, D] == 0                    # neighbour 0 and the padding are both 0
    assert g._deg[1] == 3 and g._deg[3] == 1


def test_delete_flag_and_reinsert_bookkeeping():
    g, X = small_graph()
    g.delete_node(4)
    assert g.nodes[4].is_deleted
    g._sync()
    assert g._deleted[4] and g._deleted.sum() == 1 and g._dirty
    with pytest.raises(ValueError, match="不存在"):
        g.delete_node(40)                                          # vamana_graph.py:116-125
    with pytest.raises(ValueError, match="PQ"):
        g.enable_pq_search(True)                                   # no fitted PQ model (vamana_graph.py:51-56)
    assert isinstance(VamanaGraph(8), VamanaGraphWithPQ) and VamanaGraph(8).R == 8


def test_rows_longer_than_R_are_cut_like_save_index(tmp_path):
    """The reference writes list(neighbors)[:R]; a Node whose set outgrew R must not corrupt the next record."""
    g, X = small_graph(N=8, R=3)
    g.nodes[0].neighbors = {1, 2, 3, 4, 5}
    g.nodes[7].neighbors = {6}
    p = DiskANNPersist(dim=6, R=3)
    p.save_index(tmp_path / "index.dat", g)
    rd = MMapNodeReader(tmp_path / "index.dat", dim=6, R=3)
    v0, n0 = rd.get_node(0)
    v7, n7 = rd.get_node(7)
    assert np.array_equal(v0, X[0]) and len(n0) == 3 and set(n0.tolist()) <= {1, 2, 3, 4, 5}
    assert np.array_equal(v7, X[7]) and n7.tolist() == [6, 0, 0]
    rd.close()


def test_from_arrays_roundtrip_and_degrees():
    rng = np.random.default_rng(3)
    N, D, R = 20, 5, 4
    X = rng.standard_normal((N, D)).astype(np.float32)
    adj = rng.integers(0, N, (N, R), dtype=np.uint32)
    deg = rng.integers(0, R + 1, N).astype(np.int32)
    g = VamanaGraphWithPQ.from_arrays(X, adj, deg, medoid_idx=7)
    for i in (0, 7, 19):
        assert g.nodes[i].neighbors == set(int(x) for x in adj[i, :deg[i]])
    rec = g.to_records()
    want = adj.copy(); want[np.arange(R)[None, :] >= deg[:, None]] = 0
    assert np.array_equal(rec[:, D:], want) and g.medoid_idx == 7
    # untouched nodes are not rewritten by a flush (neighbour ORDER in the arrays is part of the on-disk contract)
    _ = [g.nodes[i] for i in range(N)]
    g._sync()
    assert np.array_equal(g.to_records()[:, D:], want)
