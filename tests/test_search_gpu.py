"""GPU parity tests of K1 (batched beam search) through the C ABI.

Bar: bit-exact ids / ADC distances / hop counts / visited order against (a) the golden vectors the real
reference produced and (b) the oracle on seeded cases; exact fp32 distances within 1e-4 relative of the
reference's numpy values (bit-equal to the oracle's restatement of the GPU summation order)."""
import numpy as np
import pytest

from conftest import canon, make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gidx(golden):
    from diskrag_b200.engine import GpuIndex
    g = golden
    idx = GpuIndex.from_records(g["records"], g["N"], g["D"], g["R"], g["codes"], g["codebook"], g["medoid"])
    yield idx
    idx.close()


def test_lut_bit_exact_vs_reference(golden, gidx):
    T = gidx.lut(golden["Q"][:8])
    assert np.array_equal(T, golden["exp_lut"])


@pytest.mark.parametrize("L", [10, 40])
@pytest.mark.parametrize("own_lut", [True, False])
def test_variant_A_golden(golden, gidx, orc, L, own_lut):
    g = golden
    lut = None if own_lut else np.stack([orc.lut(g["codebook"], q) for q in g["Q"]])
    r = gidx.search(g["Q"], k=min(10, L), L=L, W=1, dist="pq", adc_order="seq", rerank=False, lut=lut, want_list=True,
                    trace=2048)
    for qi in range(g["Q"].shape[0]):
        exp_ids = g[f"exp_A_ids_L{L}"][qi]; exp_d = g[f"exp_A_dist_L{L}"][qi]
        a = canon(exp_ids, exp_d)
        n = r.list_len[qi]
        b = canon(r.list_ids[qi, :n], r.list_dists[qi, :n])
        assert np.array_equal(a[0], b[0]), qi
        assert np.array_equal(a[1], b[1]), qi            # ADC distances bit-for-bit
        h = orc.search_heap(g["adj"], g["medoid"], L, codes=g["codes"], lut_=orc.lut(g["codebook"], g["Q"][qi]),
                            dist_mode=orc.DIST_ADC_SEQ, trace=2048)
        assert r.hops[qi] == h["hops"] and r.visited[qi] == h["visited"]
        assert np.array_equal(r.trace[qi, :h["visited"]], h["trace"])   # visited / neighbour order
        # top-k without rerank = head of the list
        assert np.array_equal(r.ids[qi], b[0][:r.ids.shape[1]])


def test_rerank_golden(golden, gidx):
    g = golden
    r = gidx.search(g["Q"], k=10, L=40, W=1, dist="pq", rerank=True)
    for qi in range(g["Q"].shape[0]):
        np.testing.assert_allclose(r.dists[qi], g["exp_rerank_d2"][qi], rtol=1e-4)   # north-star tolerance
        a = canon(r.ids[qi], np.round(r.dists[qi], 5)); b = canon(g["exp_rerank_ids"][qi], np.round(g["exp_rerank_d2"][qi], 5))
        assert set(a[0]) == set(b[0])


def test_variant_D_golden(golden, gidx):
    g = golden
    r = gidx.search(g["Q"], k=10, L=40, W=1, dist="exact", rerank=False, sqrt_out=True)
    for qi in range(g["Q"].shape[0]):
        np.testing.assert_allclose(r.dists[qi], g["exp_D_dist"][qi], rtol=1e-4)
        assert set(r.ids[qi].tolist()) == set(g["exp_D_ids"][qi].tolist())


def test_variant_B_golden(golden, gidx):
    g = golden
    r = gidx.search(g["Q"], k=10, L=40, W=1, dist="exact", rerank=False, want_list=True)
    for qi in range(g["Q"].shape[0]):
        exp = g["exp_B_ids"][qi]; exp = exp[exp >= 0]
        n = r.list_len[qi]
        assert set(r.list_ids[qi, :n].tolist()) == set(exp.tolist())


CASES = [
    # N, D, M, R, Lbuild, seed, dup
    (1500, 32, 4, 8, 16, 1, 40),      # tiny codes -> massive ADC ties (ghost path)
    (3000, 96, 12, 20, 32, 2, 0),     # R not a power of two, M % 16 != 0 (byte-load path)
    (4000, 128, 32, 32, 48, 3, 16),   # M % 16 == 0 (128-bit code loads)
    (2500, 64, 16, 64, 64, 4, 0),     # R = 64: two adjacency chunks per row
    (1200, 30, 5, 12, 24, 5, 0),      # D % 4 != 0 (scalar distance path), odd M
]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"N{c[0]}D{c[1]}M{c[2]}R{c[3]}")
def case(request, orc):
    N, D, M, R, Lb, seed, dup = request.param
    from diskrag_b200.engine import GpuIndex
    c = make_case(orc, N, D, M, R, Lb, seed, nq=24, dup=dup)
    c["idx"] = GpuIndex.from_arrays(c["X"], c["adj"], c["codes"], c["codebook"], c["medoid"])
    yield c
    c["idx"].close()


@pytest.mark.parametrize("L", [1, 7, 50, 128])
def test_pq_traversal_vs_oracle(case, orc, L):
    c = case
    r = c["idx"].search(c["Q"], k=min(5, L), L=L, W=1, dist="pq", adc_order="seq", rerank=False, want_list=True, trace=8192)
    ties = 0
    for qi in range(c["Q"].shape[0]):
        T = orc.lut(c["codebook"], c["Q"][qi])
        h = orc.search_heap(c["adj"], c["medoid"], L, codes=c["codes"], lut_=T, dist_mode=orc.DIST_ADC_SEQ, trace=8192)
        n = r.list_len[qi]
        a = canon(h["ids"], h["dists"]); b = canon(r.list_ids[qi, :n], r.list_dists[qi, :n])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (qi, L)
        assert (r.hops[qi], r.visited[qi]) == (h["hops"], h["visited"]), (qi, L)
        assert np.array_equal(r.trace[qi, :min(8192, h["visited"])], h["trace"]), (qi, L)
        ties += len(set(h["dists"].tolist())) != len(h["dists"])
    if c["M"] == 4 and L >= 50:
        assert ties > 0


@pytest.mark.parametrize("L", [5, 64])
def test_exact_traversal_vs_oracle(case, orc, L):
    c = case
    r = c["idx"].search(c["Q"], k=min(5, L), L=L, W=1, dist="exact", rerank=False, want_list=True, trace=8192)
    for qi in range(c["Q"].shape[0]):
        l = orc.search_list(c["adj"], c["medoid"], L, vec=c["X"], q=c["Q"][qi], dist_mode=orc.DIST_L2_SQ,
                            flavor=orc.FLAVOR_WARP, W=1, strict_ties=True, trace=8192)
        n = r.list_len[qi]
        assert np.array_equal(l["ids"], r.list_ids[qi, :n]) and np.array_equal(l["dists"], r.list_dists[qi, :n]), (qi, L)
        assert (r.hops[qi], r.visited[qi]) == (l["hops"], l["visited"])
        assert np.array_equal(r.trace[qi, :l["visited"]], l["trace"])


@pytest.mark.parametrize("W", [2, 4, 8])
@pytest.mark.parametrize("adc", ["seq", "tree"])
def test_beam_W_vs_oracle(case, orc, W, adc):
    """Throughput mode (W > 1 expansions per step, tree ADC) is checked bit-for-bit against its restatement."""
    c = case
    L = 48
    r = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", adc_order=adc, rerank=True, want_list=True)
    for qi in range(c["Q"].shape[0]):
        T = orc.lut(c["codebook"], c["Q"][qi])
        l = orc.search_list(c["adj"], c["medoid"], L, codes=c["codes"], lut_=T,
                            dist_mode=orc.DIST_ADC_TREE if adc == "tree" else orc.DIST_ADC_SEQ, W=W, strict_ties=False)
        n = r.list_len[qi]
        assert np.array_equal(l["ids"], r.list_ids[qi, :n]) and np.array_equal(l["dists"], r.list_dists[qi, :n]), (qi, W)
        assert (r.hops[qi], r.visited[qi]) == (l["hops"], l["visited"])
        oi, od = orc.rerank(c["X"], c["Q"][qi], l["ids"], 10, flavor=orc.FLAVOR_WARP)
        assert np.array_equal(oi, r.ids[qi, :len(oi)]) and np.array_equal(od, r.dists[qi, :len(od)])


@pytest.mark.parametrize("W,L", [(12, 48), (4, 300), (16, 20)])
def test_u8_table_mode_wide_and_long_lists(case, orc, W, L):
    """more expansions per step than warps (W = 12, 16), lists longer than 255 entries and shorter than W * R:
    the rank bookkeeping of the merge (u16 ranks, selection of the next step) against the restatement"""
    c = case
    r = c["idx"].search(c["Q"][:12], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", prefetch=5)
    for qi in range(12):
        t8, sc, off = orc.lut_u8(c["codebook"], c["Q"][qi])
        l = orc.search_list(c["adj"], c["medoid"], L, codes=c["codes"], lut_=t8, dist_mode=orc.DIST_ADC_U8, W=W, strict_ties=False)
        n = r.list_len[qi]
        assert np.array_equal(l["ids"], r.list_ids[qi, :n]), (qi, W, L)
        assert (r.hops[qi], r.visited[qi]) == (l["hops"], l["visited"])
        oi, od = orc.rerank(c["X"], c["Q"][qi], l["ids"], 10, flavor=orc.FLAVOR_WARP)
        assert np.array_equal(oi, r.ids[qi, :len(oi)]) and np.array_equal(od, r.dists[qi, :len(od)])


@pytest.mark.parametrize("W", [1, 4])
def test_u8_table_mode_vs_oracle(case, orc, W):
    """Throughput mode with the 8-bit ADC table (3 CTAs per SM): bit-for-bit against its restatement."""
    c = case
    L = 48
    r = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", prefetch=(W == 4))
    r2 = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=False, lut_fmt="u8", hash_cap=256)
    # visited set kept in the CTA's global (L2-resident) table, four CTAs per SM: same answers, with and without rerank
    r3 = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", hash_cap=-1, prefetch=5)
    r4 = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=False, lut_fmt="u8", hash_cap=-1)
    assert np.array_equal(r3.ids, r.ids) and np.array_equal(r3.dists, r.dists) and np.array_equal(r3.list_ids, r.list_ids)
    assert np.array_equal(r3.hops, r.hops) and np.array_equal(r3.visited, r.visited) and np.array_equal(r4.ids, r2.ids)
    for qi in range(c["Q"].shape[0]):
        t8, sc, off = orc.lut_u8(c["codebook"], c["Q"][qi])
        l = orc.search_list(c["adj"], c["medoid"], L, codes=c["codes"], lut_=t8, dist_mode=orc.DIST_ADC_U8, W=W, strict_ties=False)
        n = r.list_len[qi]
        assert np.array_equal(l["ids"], r.list_ids[qi, :n]), (qi, W)
        exp_d = (np.float64(np.float32(sc)) * l["dists"].astype(np.float64) + np.float64(np.float32(off))).astype(np.float32)
        np.testing.assert_allclose(r.list_dists[qi, :n], exp_d, rtol=1e-6)
        assert (r.hops[qi], r.visited[qi]) == (l["hops"], l["visited"])
        oi, od = orc.rerank(c["X"], c["Q"][qi], l["ids"], 10, flavor=orc.FLAVOR_WARP)
        assert np.array_equal(oi, r.ids[qi, :len(oi)]) and np.array_equal(od, r.dists[qi, :len(od)])
        assert np.array_equal(r2.ids[qi, :min(10, n)], l["ids"][:10])      # no rerank, tiny visited table -> overflow path


@pytest.mark.parametrize("W,L", [(8, 100), (4, 48), (12, 48), (16, 20), (1, 32)])
def test_u8_prefetch_bits_never_change_results(case, W, L):
    """prefetch bits 8 (next-in-line code rows) and 16 (adjacency rows bulk-copied to shared memory at selection time) move data
    earlier, nothing else: lists, hops, visited counts and reranked results are those of the oracle-checked mask 5 — also with a
    tiny visited table (overflow path, where bit 8 switches itself off) and with the visited set in the global table."""
    c = case
    ref = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", prefetch=5)
    for pf, hc in ((13, 0), (21, 0), (29, 0), (31, 0), (29, 256), (29, -1), (16, 0)):
        r = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", prefetch=pf, hash_cap=hc)
        assert np.array_equal(r.list_ids, ref.list_ids) and np.array_equal(r.ids, ref.ids) and np.array_equal(r.dists, ref.dists), (pf, hc)
        assert np.array_equal(r.hops, ref.hops) and np.array_equal(r.visited, ref.visited), (pf, hc)


@pytest.mark.parametrize("W,W2,L", [(8, 16, 100), (8, 20, 100), (4, 8, 48), (2, 32, 64), (8, 16, 20), (1, 4, 32)])
def test_u8_empty_step_doubling_vs_oracle(case, orc, W, W2, L):
    """w_after_empty: a step that follows a step without survivors expands up to W2 entries.  Bit-for-bit against the oracle's
    restatement of the rule (lists, hops, visited, reranked results), also through the overflow path of a tiny visited table."""
    c = case
    r = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", prefetch=5, w2=W2)
    r2 = c["idx"].search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", prefetch=5, w2=W2, hash_cap=256)
    assert np.array_equal(r.list_ids, r2.list_ids) and np.array_equal(r.ids, r2.ids) and np.array_equal(r.hops, r2.hops)
    fewer = 0
    for qi in range(c["Q"].shape[0]):
        t8, sc, off = orc.lut_u8(c["codebook"], c["Q"][qi])
        l = orc.search_list(c["adj"], c["medoid"], L, codes=c["codes"], lut_=t8, dist_mode=orc.DIST_ADC_U8, W=W, strict_ties=False,
                            w_after_empty=W2)
        n = r.list_len[qi]
        assert np.array_equal(l["ids"], r.list_ids[qi, :n]), (qi, W, W2)
        assert (r.hops[qi], r.visited[qi]) == (l["hops"], l["visited"])
        oi, od = orc.rerank(c["X"], c["Q"][qi], l["ids"], 10, flavor=orc.FLAVOR_WARP)
        assert np.array_equal(oi, r.ids[qi, :len(oi)]) and np.array_equal(od, r.dists[qi, :len(od)])


def test_visited_overflow_table(case, orc):
    """Force a tiny shared-memory visited table so the global overflow table is exercised; results must not change."""
    c = case
    L = 100
    r0 = c["idx"].search(c["Q"], k=10, L=L, W=1, dist="pq", rerank=False, want_list=True)
    r1 = c["idx"].search(c["Q"], k=10, L=L, W=1, dist="pq", rerank=False, want_list=True, hash_cap=256)
    assert r0.visited.max() > 192   # 3/4 of 256: the overflow path really ran
    assert np.array_equal(r0.list_ids, r1.list_ids) and np.array_equal(r0.list_dists, r1.list_dists)
    assert np.array_equal(r0.hops, r1.hops) and np.array_equal(r0.visited, r1.visited)
    r2 = c["idx"].search(c["Q"], k=10, L=L, W=1, dist="pq", rerank=False, want_list=True)   # table was cleaned
    assert np.array_equal(r0.list_ids, r2.list_ids)


def test_edge_cases(case):
    c = case
    idx = c["idx"]
    r = idx.search(np.zeros((0, c["D"]), np.float32), k=3, L=8)
    assert r.ids.shape == (0, 3)
    r = idx.search(c["Q"][:1], k=10, L=10, dist="exact", rerank=False)     # single query
    assert (r.ids[0] >= 0).all()
    with pytest.raises(ValueError):
        idx.search(np.zeros((1, c["D"] + 1), np.float32))                  # dimension mismatch
    with pytest.raises(ValueError):
        idx.search(c["Q"][:1], k=20, L=10)                                 # k > L
    # chunked launches give the same answers as one launch
    a = idx.search(c["Q"], k=5, L=32, chunk=5)
    b = idx.search(c["Q"], k=5, L=32)
    assert np.array_equal(a.ids, b.ids) and np.array_equal(a.dists, b.dists)
    # k larger than what the graph can reach is padded with -1 / inf
    from diskrag_b200.engine import GpuIndex
    tiny = GpuIndex.from_arrays(c["X"][:3], np.array([[1, 0], [0, 0], [2, 2]], np.uint32), medoid=0)
    r = tiny.search(c["Q"][:2], k=3, L=3, dist="exact", rerank=False)
    assert (r.ids[:, 2] == -1).all() and np.isinf(r.dists[:, 2]).all() and (r.ids[:, :2] >= 0).all()
    tiny.close()


def test_export_records_roundtrip(golden, gidx):
    assert np.array_equal(gidx.export_records(), golden["records"])


def test_bench_shape_specialisation_vs_oracle(orc):
    """The compile-time-specialised instantiation the bench runs (D = 1536, M = 192, R = 32, W = 8, 20 after an empty step,
    L = 100, 4096-slot visited table, prefetch mask 5, rerank): bit-for-bit against the restatement with the exact 8-bit table, and the same ids with
    the tensor-core table on all but near-tie queries."""
    from diskrag_b200.engine import GpuIndex
    c = make_case(orc, 2500, 1536, 192, 32, 48, 31, nq=16)
    L, W = 100, 8
    with GpuIndex.from_arrays(c["X"], c["adj"], c["codes"], c["codebook"], c["medoid"]) as idx:
        r = idx.search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=False, lut_fmt="u8", prefetch=5, w2=20)
        rl = idx.search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, want_list=True, lut_fmt="u8", prefetch=5, w2=20)   # generic flags path
        rt = idx.search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, lut_fmt="u8tc", prefetch=5, w2=20)
        rp = [idx.search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, lut_fmt="u8", prefetch=pf) for pf in (13, 21, 29)]
        r0 = idx.search(c["Q"], k=10, L=L, W=W, dist="pq", rerank=True, lut_fmt="u8", prefetch=5)
    assert np.array_equal(r.ids, rl.ids) and np.array_equal(r.dists, rl.dists) and np.array_equal(r.hops, rl.hops)
    for x in rp:    # the prefetch masks change nothing (fixed W: the generic instantiation)
        assert np.array_equal(r0.ids, x.ids) and np.array_equal(r0.dists, x.dists) and np.array_equal(r0.hops, x.hops) and np.array_equal(r0.visited, x.visited)
    for qi in range(c["Q"].shape[0]):
        t8, sc, off = orc.lut_u8(c["codebook"], c["Q"][qi])
        l0 = orc.search_list(c["adj"], c["medoid"], L, codes=c["codes"], lut_=t8, dist_mode=orc.DIST_ADC_U8, W=W, strict_ties=False)
        assert (r0.hops[qi], r0.visited[qi]) == (l0["hops"], l0["visited"])
        oi0, od0 = orc.rerank(c["X"], c["Q"][qi], l0["ids"], 10, flavor=orc.FLAVOR_WARP)
        assert np.array_equal(oi0, r0.ids[qi, :len(oi0)]) and np.array_equal(od0, r0.dists[qi, :len(od0)])
        l = orc.search_list(c["adj"], c["medoid"], L, codes=c["codes"], lut_=t8, dist_mode=orc.DIST_ADC_U8, W=W, strict_ties=False,
                            w_after_empty=20)
        assert (r.hops[qi], r.visited[qi]) == (l["hops"], l["visited"])
        assert np.array_equal(l["ids"], rl.list_ids[qi, :rl.list_len[qi]])
        oi, od = orc.rerank(c["X"], c["Q"][qi], l["ids"], 10, flavor=orc.FLAVOR_WARP)
        assert np.array_equal(oi, r.ids[qi, :len(oi)]) and np.array_equal(od, r.dists[qi, :len(od)])
    assert np.mean(np.all(rt.ids == r.ids, axis=1)) >= 0.8


def test_cosine_traversal_vs_oracle_and_reference(case, orc):
    """distance_metric='cosine' (compute_query_distance, vamana_graph.py:301-329 -> cosine_similarity_cython,
    cython_utils.pyx:53-70: the distance is 1 - cos): the GPU traversal is bit-for-bit the restatement in the GPU's summation
    order, and agrees with the REAL reference's greedy_search_cython on the same graph up to near-ties (the reference sums
    sequentially in fp32 under -ffast-math, so only a tolerance comparison is meaningful: distances within 1e-5)."""
    import sys
    from pathlib import Path
    c = case
    L = 40
    r = c["idx"].search(c["Q"], k=10, L=L, W=1, dist="cosine", rerank=False, want_list=True)
    for qi in range(c["Q"].shape[0]):
        h = orc.search_heap(c["adj"], c["medoid"], L, vec=c["X"], q=c["Q"][qi], dist_mode=orc.DIST_COSINE, flavor=orc.FLAVOR_WARP)
        n = r.list_len[qi]
        ei, ed = canon(h["ids"], h["dists"])
        gi, gd = canon(r.list_ids[qi, :n], r.list_dists[qi, :n])
        assert np.array_equal(ei, gi) and np.array_equal(ed, gd), qi
        assert (r.hops[qi], r.visited[qi]) == (h["hops"], h["visited"])
    with pytest.raises(ValueError):
        c["idx"].search(c["Q"][:1], k=5, L=L, dist="cosine", rerank=True)        # the fused rerank is a squared-L2 rerank
    # the shim: a graph object with distance_metric='cosine' searches through the same kernel
    from diskrag_b200.vamana_graph import VamanaGraphWithPQ, greedy_search
    g = VamanaGraphWithPQ.from_arrays(c["X"], c["adj"], medoid_idx=c["medoid"], distance_metric='cosine')
    ids = greedy_search(g, c["medoid"], c["Q"][0], L)
    assert ids == [int(x) for x in r.list_ids[0, :r.list_len[0]]]
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    ref = ref_loader.load()
    vg, cu = ref["vamana_graph"], ref["cython_utils"]
    rg = vg.VamanaGraphWithPQ(c["R"], None, distance_metric='cosine')
    for i in range(c["N"]):
        node = vg.Node(i, c["X"][i], None)
        node.neighbors = [int(x) for x in c["adj"][i]]
        rg.nodes[i] = node
    rg.medoid_idx = c["medoid"]
    same, overlap = 0, []
    for qi in range(8):
        rid = cu.greedy_search_cython(rg, c["medoid"], c["Q"][qi], L, vg.compute_query_distance)
        mine = [int(x) for x in r.list_ids[qi, :r.list_len[qi]]]
        same += (set(rid) == set(mine))            # order inside an exact tie group is heap-layout dependent in the reference
        overlap.append(len(set(rid) & set(mine)) / max(1, len(rid)))
        dref = np.array([orc.cosine_dist(c["X"][i], c["Q"][qi]) for i in rid], np.float64)
        dgpu = np.sort(r.list_dists[qi, :r.list_len[qi]].astype(np.float64))
        np.testing.assert_allclose(np.sort(dref), dgpu, atol=1e-5)
    assert same >= 5 and np.mean(overlap) >= 0.95, (same, overlap)


def test_concurrent_callers_on_one_handle(golden, gidx):
    """ctypes releases the GIL, so FastAPI worker threads can enter dr_search_batch on the same handle at the same time; the
    handle's scratch (tables, staging, counters) is shared, so the library serialises them (per-handle mutex): every thread
    must get exactly the single-threaded answer, in every mode."""
    import threading
    Q = golden["Q"]
    modes = [dict(W=1, dist="pq", adc_order="seq", rerank=False), dict(W=4, dist="pq", rerank=True, lut_fmt="u8"),
             dict(W=1, dist="exact", rerank=False), dict(W=2, dist="pq", adc_order="tree", rerank=True)]
    want = [gidx.search(Q, k=10, L=40, **m) for m in modes]
    errs = []

    def worker(t):
        try:
            for it in range(6):
                j = (t + it) % len(modes)
                r = gidx.search(Q, k=10, L=40, **modes[j])
                if not (np.array_equal(r.ids, want[j].ids) and np.array_equal(r.dists, want[j].dists)
                        and np.array_equal(r.hops, want[j].hops)):
                    errs.append((t, it, j))
        except Exception as e:          # noqa: BLE001
            errs.append((t, repr(e)))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs[:4]


@pytest.mark.gpu
def test_lazy_delete_mask_equals_the_oracle(orc):
    """§8 a14: with a delete mask (dr_index_set_deleted) the device search equals the oracle's restatement of the reference's
    is_deleted handling (cython_utils.pyx:100-109, pinned live in test_oracle_vs_reference.py::test_variant_A_with_lazily_deleted_nodes):
    reference order (PQ and exact, W = 1) with lists, hops and visited counts bit-for-bit; throughput mode (u8 table, W = 4, rerank)."""
    from diskrag_b200._lib import check, lib, ptr
    from diskrag_b200.engine import GpuIndex
    c = make_case(orc, 2000, 64, 8, 16, 32, seed=12, nq=24, dup=30)
    rng = np.random.default_rng(4)
    dead = np.zeros(c["N"], np.uint8)
    dead[rng.choice(c["N"], 150, replace=False)] = 1
    dead[c["medoid"]] = 0
    L = 40
    with GpuIndex.from_arrays(c["X"], c["adj"], c["codes"], c["codebook"], c["medoid"]) as idx:
        check(lib().dr_index_set_deleted(idx._h, ptr(dead)))
        rp = idx.search(c["Q"], k=10, L=L, W=1, dist="pq", adc_order="seq", rerank=False, want_list=True)
        rx = idx.search(c["Q"], k=10, L=L, W=1, dist="exact", rerank=False, want_list=True)
        r8 = idx.search(c["Q"], k=10, L=L, W=4, dist="pq", rerank=True, lut_fmt="u8")
    for qi, q in enumerate(c["Q"]):
        for r, kw in ((rp, dict(codes=c["codes"], lut_=orc.lut(c["codebook"], q), dist_mode=orc.DIST_ADC_SEQ)),
                      (rx, dict(vec=c["X"], q=q, dist_mode=orc.DIST_L2_SQ, flavor=orc.FLAVOR_WARP))):
            l = orc.search_list(c["adj"], c["medoid"], L, W=1, strict_ties=True, deleted=dead, **kw)
            n = int(r.list_len[qi])
            assert np.array_equal(l["ids"], r.list_ids[qi, :n]) and np.array_equal(l["dists"], r.list_dists[qi, :n])
            assert (int(r.hops[qi]), int(r.visited[qi])) == (l["hops"], l["visited"]) and not dead[l["ids"]].any()
        t8, _, _ = orc.lut_u8(c["codebook"], q)
        l = orc.search_list(c["adj"], c["medoid"], L, codes=c["codes"], lut_=t8, dist_mode=orc.DIST_ADC_U8, W=4, strict_ties=False,
                            deleted=dead)
        oi, od = orc.rerank(c["X"], q, l["ids"], 10, flavor=orc.FLAVOR_WARP)
        assert np.array_equal(oi, r8.ids[qi]) and np.array_equal(od, r8.dists[qi])
