"""K2 on tensor cores (csrc/lut_tc.cu, tcgen05 TF32) against the exact 8-bit table (csrc/pq.cu, itself pinned bit-for-bit
to oracle.c:orc_lut_u8, which restates DiskANNPQ.compute_distance_table, fast_pq.py:294-318, in the u8 contract).

Tolerance (stated here, as the table is floating-point work rounded to bytes): TF32 carries 10 mantissa bits into
the products, so an entry may land on the other side of a rounding boundary: |tc - exact| <= 1 for >= 99.9 % of the
entries and <= 2 everywhere; scale within 2e-3 relative; the offset (a per-query constant that shifts the reported ADC distances, never their order)
within 2e-3 of max(1, |offset|)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(N, D, M, B, seed):
    rng = np.random.default_rng(seed)
    from diskrag_b200.synth import synth_numpy
    X = synth_numpy(N, D, seed=seed)
    Q = synth_numpy(B, D, seed=seed, sample_seed=seed + 7)
    # a codebook of real data points per subspace: realistic norms and spreads
    ds = D // M
    idx = rng.choice(N, 256, replace=False)
    cb = np.ascontiguousarray(X[idx].reshape(256, M, ds).transpose(1, 0, 2)).astype(np.float32)
    return cb, Q, X


@pytest.mark.parametrize("D,M,B", [(64, 8, 16), (1536, 192, 300), (384, 16, 129), (512, 32, 64)])
def test_exact_u8_table_matches_oracle(orc, D, M, B):
    from diskrag_b200 import ops
    cb, Q, _ = _case(2000, D, M, B, 5)
    tab, sc, off = ops.pq_lut_u8(cb, Q, "u8")
    for qi in range(0, B, max(1, B // 8)):
        t8, s, o = orc.lut_u8(cb, Q[qi])
        assert np.array_equal(tab[qi], t8)
        assert sc[qi] == np.float32(s) and off[qi] == np.float32(o)


@pytest.mark.parametrize("D,M,B", [(64, 8, 16), (1536, 192, 300), (384, 16, 129), (512, 32, 64)])
def test_tensor_core_table_within_one_unit(D, M, B):
    from diskrag_b200 import ops
    cb, Q, _ = _case(2000, D, M, B, 11)
    te, se, oe = ops.pq_lut_u8(cb, Q, "u8")
    tt, st, ot = ops.pq_lut_u8(cb, Q, "u8tc")
    d = np.abs(te.astype(np.int16) - tt.astype(np.int16))
    assert d.max() <= 2, f"max entry difference {d.max()}"
    assert (d <= 1).mean() >= 0.999
    assert (d == 0).mean() >= 0.80, f"only {(d == 0).mean():.3f} of the entries are identical"
    np.testing.assert_allclose(st, se, rtol=2e-3)
    assert np.all(np.abs(ot - oe) <= 2e-3 * np.maximum(1.0, np.abs(oe)))


def test_search_with_tensor_core_table_keeps_recall():
    """same graph, same codes: the tensor-core table may reorder near-ties, recall@10 must not move by more than 0.5 pt"""
    from diskrag_b200 import ops
    from diskrag_b200.engine import GpuIndex
    from diskrag_b200.synth import synth_numpy
    N, D, M, R = 20000, 256, 32, 32
    X = synth_numpy(N, D, seed=3)
    Q = synth_numpy(512, D, seed=3, sample_seed=99)
    from diskrag_b200.pq.fast_pq import DiskANNPQ
    pq = DiskANNPQ(M, 256); pq.fit(X)
    codes = pq.encode(X)
    cb = np.stack([km.cluster_centers_ for km in pq.kmeans_list]).astype(np.float32)
    adj, deg = ops.vamana_build(X, R, 64, 1.2, 0, seed=1)
    gt = np.argsort(-2.0 * Q @ X.T + (X * X).sum(1)[None, :], axis=1)[:, :10]
    with GpuIndex.from_arrays(X, adj, codes, cb, 0, device=0) as idx:
        r0 = idx.search(Q, k=10, L=64, W=4, dist="pq", rerank=True, lut_fmt="u8")
        r1 = idx.search(Q, k=10, L=64, W=4, dist="pq", rerank=True, lut_fmt="u8tc")
    rec = lambda ids: np.mean([len(set(ids[i]) & set(gt[i])) / 10 for i in range(len(gt))])
    assert abs(rec(r0.ids) - rec(r1.ids)) <= 0.005
    assert np.mean(np.all(r0.ids == r1.ids, axis=1)) >= 0.9


@pytest.mark.parametrize("D,M", [(64, 8), (256, 16), (384, 16)])
def test_kmeans_tensor_core_assignment_quality(D, M):
    """K3: the tcgen05 (TF32) assignment step against the exact fp32 one, same seed and iteration count: judged by the
    quantisation error of the trained codebook (final pass and encode are exact in both)."""
    from diskrag_b200._lib import lib
    from diskrag_b200.pq.fast_pq import DiskANNPQ
    from diskrag_b200.synth import synth_numpy
    X = synth_numpy(20000, D, seed=9)
    try:
        lib().dr_pq_train_tensor_cores(0)
        p0 = DiskANNPQ(M); p0.fit(X)
        lib().dr_pq_train_tensor_cores(1)
        p1 = DiskANNPQ(M); p1.fit(X)
    finally:
        lib().dr_pq_train_tensor_cores(1)
    assert p1.train_mse_ <= 1.02 * p0.train_mse_, (p1.train_mse_, p0.train_mse_)
    c1 = p1.encode(X)
    mse1 = float(((p1.decode(c1) - X) ** 2).mean())
    assert abs(mse1 - p1.train_mse_) <= 0.05 * mse1
