"""GPU parity of K2/K3/K4/K6: PQ tables, encode/decode/train, fused distances, top-k merge, medoid."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_distance_helpers_known_answers(golden):
    from diskrag_b200 import cython_utils as cu
    g = golden
    for i in range(16):
        # the reference's own tolerances (scripts/test_pydiskann_cython.sh:36-56)
        np.testing.assert_allclose(cu.l2_distance_fast_cython(g["ka_x"][i], g["ka_y"][i]), g["exp_l2"][i], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(cu.cosine_similarity_cython(g["ka_x"][i], g["ka_y"][i]), g["exp_cos"][i], rtol=1e-5, atol=1e-6)
    z = np.zeros(128, np.float32)
    assert cu.cosine_similarity_cython(z, g["ka_y"][0]) == 0.0          # zero-norm rule (cython_utils.pyx:65-66)
    with pytest.raises(ValueError):
        cu.l2_distance_fast_cython(g["ka_x"][0].astype(np.float64), g["ka_y"][0])   # Buffer dtype mismatch


def test_batched_distances_vs_oracle(orc):
    from diskrag_b200 import ops
    rs = np.random.RandomState(5)
    for D in (1, 7, 64, 1536, 1538):
        A = rs.randn(33, D).astype(np.float32); B = rs.randn(33, D).astype(np.float32)
        l2 = ops.l2sq_batch(A, B); cs = ops.cosine_batch(A, B); dt = ops.dot_batch(A, B)
        for i in range(33):
            assert l2[i] == np.float32(orc.l2sq(A[i], B[i], orc.FLAVOR_WARP))         # bit-exact vs the restated order
            np.testing.assert_allclose(cs[i], orc.cosine_dist(A[i], B[i]), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(dt[i], float(A[i].astype(np.float64) @ B[i].astype(np.float64)), rtol=1e-4, atol=1e-4)
        l2b = ops.l2sq_batch(A, B[:1])
        assert l2b[3] == np.float32(orc.l2sq(A[3], B[0], orc.FLAVOR_WARP))
    assert ops.l2sq_batch(np.zeros((0, 8), np.float32), np.zeros((0, 8), np.float32)).shape == (0,)


def test_pq_object_against_reference_golden(golden, orc):
    from diskrag_b200.pq.fast_pq import DiskANNPQ, _wrap_kmeans
    g = golden
    pq = DiskANNPQ(g["M"])
    pq.sub_dim = g["D"] // g["M"]
    pq.kmeans_list = [_wrap_kmeans(g["codebook"][i], 42 + i) for i in range(g["M"])]
    pq.is_fitted = True
    # LUT: bit-identical to the reference's numpy expression
    for i in range(8):
        assert np.array_equal(pq.compute_distance_table(g["Q"][i]), g["exp_lut"][i])
    # ADC: bit-identical sequential fp32 sums
    T = g["exp_lut"][0]
    assert np.array_equal(pq.asymmetric_distance_sq(g["codes"], T), orc.adc(g["codes"], T))
    np.testing.assert_array_equal(pq.asymmetric_distance(g["codes"][:5], T), np.sqrt(orc.adc(g["codes"][:5], T)))
    # encode: identical to the reference's codes except provable near-ties
    codes = pq.encode(g["vec"])
    assert codes.dtype == np.uint8 and codes.shape == g["codes"].shape
    assert (codes == g["codes"]).mean() >= 0.9995
    ds = g["D"] // g["M"]
    for i, m in np.argwhere(codes != g["codes"]):
        x = g["vec"][i, m * ds:(m + 1) * ds].astype(np.float64)
        d = ((g["codebook"][m].astype(np.float64) - x) ** 2).sum(1)
        assert abs(d[codes[i, m]] - d[g["codes"][i, m]]) <= 1e-5 * max(d.min(), 1e-12) + 1e-9
    # decode == centroid gather
    dec = pq.decode(g["codes"][:100])
    exp = np.concatenate([g["codebook"][m][g["codes"][:100, m]] for m in range(g["M"])], axis=1)
    assert np.array_equal(dec, exp)
    # sdc
    from diskrag_b200 import cython_utils as cu
    for i in range(16):
        np.testing.assert_allclose(cu.pq_distance_fast_cython(pq, g["codes"][i], g["codes"][i + 1]), g["exp_sdc"][i], rtol=1e-5)
    # error behaviour (fast_pq.py:207-213, 255-256, 304-305)
    with pytest.raises(ValueError):
        DiskANNPQ(7).fit(g["vec"])                    # D % M != 0
    with pytest.raises(ValueError):
        DiskANNPQ(8).fit(g["vec"][:100])              # N < 256
    with pytest.raises(ValueError):
        DiskANNPQ(8).encode(g["vec"])                 # unfitted
    with pytest.raises(ValueError):
        DiskANNPQ(8).compute_distance_table(g["Q"][0])


@pytest.mark.parametrize("D,M", [(64, 8), (96, 4), (1536, 192), (60, 5)])
def test_lut_shapes_vs_oracle(orc, D, M):
    from diskrag_b200.pq.fast_pq import DiskANNPQ, _wrap_kmeans
    rs = np.random.RandomState(D + M)
    cb = rs.randn(M, 256, D // M).astype(np.float32)
    pq = DiskANNPQ(M); pq.sub_dim = D // M; pq.is_fitted = True
    pq.kmeans_list = [_wrap_kmeans(cb[i], i) for i in range(M)]
    Q = rs.randn(70, D).astype(np.float32)          # > one query tile, ragged tail
    T = pq.compute_distance_tables(Q)
    for i in (0, 1, 63, 64, 69):
        assert np.array_equal(T[i], orc.lut(cb, Q[i]))
        diff = cb - Q[i].reshape(M, 1, D // M)
        assert np.array_equal(T[i], np.sum(diff * diff, axis=2).astype(np.float32))   # numpy itself


def test_kmeans_quality_vs_sklearn(golden):
    """GPU Lloyd is not bit-comparable with sklearn's k-means++: judged by quantisation MSE."""
    from diskrag_b200.pq.fast_pq import DiskANNPQ
    g = golden
    pq = DiskANNPQ(g["M"])
    pq.fit(g["vec"])
    assert pq.is_fitted and len(pq.kmeans_list) == g["M"] and pq.kmeans_list[0].cluster_centers_.shape == (256, g["D"] // g["M"])
    assert pq.kmeans_list[0].cluster_centers_.dtype == np.float32 and pq.kmeans_list[0].n_clusters == 256
    codes = pq.encode(g["vec"])
    mse_gpu = float(((pq.decode(codes) - g["vec"]) ** 2).mean())
    ref_dec = np.concatenate([g["codebook"][m][g["codes"][:, m]] for m in range(g["M"])], axis=1)
    mse_ref = float(((ref_dec - g["vec"]) ** 2).mean())
    assert mse_gpu <= 1.05 * mse_ref, (mse_gpu, mse_ref)      # sklearn: k-means++, n_init=10; here k-means++ on the device, one run (the configs[0] A/B in test_golden_config0.py gates at 1.03)
    assert abs(pq.train_mse_ - mse_gpu) <= 0.05 * mse_gpu
    assert 0.5 < pq.estimate_selectivity(g["vec"]) <= 1.0


def test_medoid_vs_reference(golden):
    from diskrag_b200 import cython_utils as cu
    assert cu.compute_approximate_medoid_cython(golden["vec"][:500], sample_size=1000) == golden["exp_medoid_500"]


def test_topk_merge(orc):
    import torch
    from diskrag_b200 import ops
    rs = np.random.RandomState(9)
    G, B, k = 8, 257, 10
    d = np.sort(rs.rand(G, B, k).astype(np.float32), axis=2)
    ids = rs.permutation(G * B * k).astype(np.int32).reshape(G, B, k)
    ids[3, :, 7:] = -1; d[3, :, 7:] = np.inf                     # a shard that found fewer than k
    d[1, 5, 0] = d[0, 5, 0]                                       # a tie: lower id wins
    oi, od = ops.topk_merge(torch.from_numpy(ids).cuda(), torch.from_numpy(d).cuda())
    oi, od = oi.cpu().numpy(), od.cpu().numpy()
    for b in range(B):
        flat_i = ids[:, b, :].ravel(); flat_d = d[:, b, :].ravel()
        keep = flat_i >= 0
        o = np.lexsort((flat_i[keep], flat_d[keep]))[:k]
        assert np.array_equal(oi[b], flat_i[keep][o]) and np.array_equal(od[b], flat_d[keep][o])


def test_packed_exchange_kernels_equal_their_numpy_statement():
    """dr_topk_pack_dev / dr_topk_merge_keys_dev (the single-exchange form of the index-sharded merge) against
    dist.pack_topk_numpy / merge_keys_numpy, ragged slices, short lists, negative-zero and tied distances included."""
    import torch
    from diskrag_b200 import dist as D
    rs = np.random.RandomState(4)
    for G, B, k in ((1, 9, 4), (2, 37, 5), (8, 1001, 10), (3, 2, 7)):
        d = np.sort(rs.rand(B, k).astype(np.float32), axis=1)
        ids = rs.randint(0, 1 << 20, size=(B, k)).astype(np.int32)
        ids[::3, k - 2:] = -1; d[::3, k - 2:] = np.inf
        d[0, 0] = -0.0
        off = 123456
        send = D._pack_gpu(torch.from_numpy(ids).cuda(), torch.from_numpy(d).cuda(), off, G)
        assert np.array_equal(send.cpu().numpy(), D.pack_topk_numpy(ids, d, off, G))
        # pretend G ranks sent their blocks: merge G lists per query
        keys = np.stack([D.pack_topk_numpy(np.where(ids >= 0, (ids + 7 * g) % (1 << 20), -1).astype(np.int32), np.sort(d + np.float32(0.01 * g), axis=1), off, 1)[0]
                         for g in range(G)])
        keys[1 % G, 0, 0] = keys[0, 0, 0]                                 # an exact duplicate key: both copies survive the merge
        oi, od = D._merge_keys_gpu(torch.from_numpy(keys).cuda())
        ei, ed = D.merge_keys_numpy(keys)
        assert np.array_equal(oi.cpu().numpy(), ei) and np.array_equal(od.cpu().numpy(), ed)


def test_peer_routed_epilogue_on_one_device(orc):
    """The index-sharded exchange fused into the search kernel (dr_index_set_peer_route), exercised on ONE device: G receive buffers
    (one per would-be rank) all live here; the same index plays every rank in turn (its own id offset), the kernel's epilogue stores
    each query's packed top-k into the buffer of the rank that reduces the query; merging every buffer must give what the explicit
    pack -> exchange -> merge path gives (numpy statements of both kernels) — ragged slices, chunked launches, fewer than k hits."""
    import ctypes as C
    import torch
    from conftest import make_case
    from diskrag_b200 import dist as D
    from diskrag_b200._lib import check, lib
    from diskrag_b200.engine import GpuIndex
    c = make_case(orc, 1500, 64, 16, 16, 32, 13, nq=37)
    B, k, G = 37, 10, 3
    bq = D.padded_slice_len(B, G)
    with GpuIndex.from_arrays(c["X"], c["adj"], c["codes"], c["codebook"], c["medoid"]) as idx:
        plain = idx.search(c["Q"], k=k, L=24, W=4, dist="pq", rerank=True, lut_fmt="u8")
        bufs = [torch.full((G, bq, k), -1, dtype=torch.int64, device="cuda") for _ in range(G)]
        table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
        for r in range(G):                                   # "rank" r: same shard content, ids offset by r * N
            check(lib().dr_index_set_peer_route(idx._h, C.c_void_p(table.data_ptr()), G, r, B, r * c["N"]), "dr_index_set_peer_route")
            routed = idx.search(c["Q"], k=k, L=24, W=4, dist="pq", rerank=True, lut_fmt="u8", chunk=(0 if r == 0 else 11))
            assert np.array_equal(routed.ids, plain.ids) and np.array_equal(routed.dists, plain.dists)      # local output unchanged
        check(lib().dr_index_set_peer_route(idx._h, None, 0, 0, 0, 0), "dr_index_set_peer_route")
        torch.cuda.synchronize()
    for g in range(G):
        lo, hi = D.query_slice(B, g, G)
        got_i, got_d = D.merge_keys_numpy(bufs[g].cpu().numpy())
        # the explicit path: every rank packs its lists, block g of every send buffer goes to rank g
        send = [D.pack_topk_numpy(plain.ids, plain.dists, r * c["N"], G) for r in range(G)]
        exp_i, exp_d = D.merge_keys_numpy(np.stack([send[r][g] for r in range(G)]))
        assert np.array_equal(got_i[:hi - lo], exp_i[:hi - lo]) and np.array_equal(got_d[:hi - lo], exp_d[:hi - lo]), g
        assert (got_i[hi - lo:] == -1).all()                 # rows past the slice stay empty
