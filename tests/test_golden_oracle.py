"""The oracle (oracle/oracle.c) against the golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only.  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import ROOT, canon


def test_lut_bit_exact(golden, orc):
    for i in range(golden["exp_lut"].shape[0]):
        T = orc.lut(golden["codebook"], golden["Q"][i])
        assert np.array_equal(T, golden["exp_lut"][i])  # bit-for-bit (fast_pq.py:294-318)


@pytest.mark.parametrize("L", [10, 40])
def test_variant_A_pq_traversal(golden, orc, L):
    """greedy_search_cython + ADC callback: heap form reproduces ids in the reference's output order and
    the ADC distances bit-for-bit; the list form (what the GPU runs) equals it after canonical ordering."""
    g = golden
    n_ties = 0
    for qi in range(g["Q"].shape[0]):
        T = orc.lut(g["codebook"], g["Q"][qi])
        exp_ids = g[f"exp_A_ids_L{L}"][qi]; exp_d = g[f"exp_A_dist_L{L}"][qi]
        n = int((exp_ids >= 0).sum())
        h = orc.search_heap(g["adj"], g["medoid"], L, codes=g["codes"], lut_=T, dist_mode=orc.DIST_ADC_SEQ, trace=4096)
        assert list(h["ids"]) == list(exp_ids[:n])
        assert np.array_equal(h["dists"], exp_d[:n])
        l = orc.search_list(g["adj"], g["medoid"], L, codes=g["codes"], lut_=T, dist_mode=orc.DIST_ADC_SEQ, W=1,
                            strict_ties=True, trace=4096)
        a, b = canon(h["ids"], h["dists"]), canon(l["ids"], l["dists"])
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert h["hops"] == l["hops"] and h["visited"] == l["visited"]
        assert np.array_equal(h["trace"], l["trace"])          # bit-exact visited / neighbour order
        n_ties += len(set(h["dists"].tolist())) != len(h["dists"])
    assert n_ties > 0  # the fixture contains duplicate rows, so tie handling is exercised


def test_rerank_composition(golden, orc):
    g = golden
    for qi in range(g["Q"].shape[0]):
        ids = g["exp_A_ids_L40"][qi]; ids = ids[ids >= 0]
        oi, od = orc.rerank(g["vec"], g["Q"][qi], ids, 10, flavor=orc.FLAVOR_NUMPY)
        assert np.array_equal(od, g["exp_rerank_d2"][qi])       # numpy pairwise order reproduced bit-for-bit
        assert list(oi) == list(g["exp_rerank_ids"][qi])
        oi2, od2 = orc.rerank(g["vec"], g["Q"][qi], ids, 10, flavor=orc.FLAVOR_WARP)  # the GPU's order
        np.testing.assert_allclose(od2, g["exp_rerank_d2"][qi], rtol=1e-5)
        a, b = canon(oi2, np.round(od2, 5)), canon(g["exp_rerank_ids"][qi], np.round(g["exp_rerank_d2"][qi], 5))
        assert set(a[0]) == set(b[0])


def test_variant_D_disk_beam_search(golden, orc):
    g = golden
    for qi in range(g["Q"].shape[0]):
        r = orc.search_heap(g["adj"], g["medoid"], 40, vec=g["vec"], q=g["Q"][qi], dist_mode=orc.DIST_L2_SQRT,
                            flavor=orc.FLAVOR_DOUBLE, truncate_frontier=True)
        np.testing.assert_allclose(r["dists"][:10], g["exp_D_dist"][qi], rtol=1e-5)  # BLAS order unknowable
        a = canon(r["ids"][:10], np.round(r["dists"][:10], 5)); b = canon(g["exp_D_ids"][qi], np.round(g["exp_D_dist"][qi], 5))
        assert np.array_equal(a[0], b[0])


def test_variant_B_exact_greedy(golden, orc):
    g = golden
    for qi in range(g["Q"].shape[0]):
        exp = g["exp_B_ids"][qi]; exp = exp[exp >= 0]
        r = orc.search_heap(g["adj"], g["medoid"], 40, vec=g["vec"], q=g["Q"][qi], dist_mode=orc.DIST_L2_SQRT,
                            flavor=orc.FLAVOR_DOUBLE)
        assert set(r["ids"].tolist()) == set(exp.tolist())
        # the GPU's formulation: squared distances in warp order, list form
        l = orc.search_list(g["adj"], g["medoid"], 40, vec=g["vec"], q=g["Q"][qi], dist_mode=orc.DIST_L2_SQ,
                            flavor=orc.FLAVOR_WARP, W=1)
        assert set(l["ids"].tolist()) == set(exp.tolist())


def test_known_answer_distances(golden, orc):
    g = golden
    for i in range(16):
        # tolerances of the reference's own test (scripts/test_pydiskann_cython.sh:36-56): rtol 1e-5, atol 1e-6
        np.testing.assert_allclose(orc.l2sq(g["ka_x"][i], g["ka_y"][i], orc.FLAVOR_SEQ), g["exp_l2"][i], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(orc.l2sq(g["ka_x"][i], g["ka_y"][i], orc.FLAVOR_WARP), g["exp_l2"][i], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(orc.cosine_dist(g["ka_x"][i], g["ka_y"][i]), g["exp_cos"][i], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(orc.pq_sdc(g["codebook"], g["codes"][i], g["codes"][i + 1]), g["exp_sdc"][i], rtol=1e-5)


def test_encode_matches_reference(golden, orc):
    g = golden
    codes = orc.pq_encode(g["codebook"], g["vec"])
    agree = (codes == g["codes"]).mean()
    assert agree >= 0.9995  # sklearn predict vs exact argmin: only provable near-ties may differ
    bad = np.argwhere(codes != g["codes"])
    ds = g["D"] // g["M"]
    for i, m in bad:
        x = g["vec"][i, m * ds:(m + 1) * ds].astype(np.float64)
        d = ((g["codebook"][m].astype(np.float64) - x) ** 2).sum(1)
        assert abs(d[codes[i, m]] - d[g["codes"][i, m]]) <= 1e-5 * max(d[codes[i, m]], 1e-12)


def test_medoid(golden, orc):
    g = golden
    assert orc.medoid(g["vec"][:500], np.arange(500, dtype=np.int32), skip_self=True) == g["exp_medoid_500"]


def test_sequential_build_bit_equal(golden, orc):
    g = golden
    NB, RB = g["exp_build_adj"].shape
    rows = orc.vamana_build(g["vec"][:NB], RB, 16, 1.2, 5, g["build_sigma0"], g["build_sigma1"])
    same = 0
    for i, row in enumerate(rows):
        exp = g["exp_build_adj"][i]; exp = exp[exp >= 0]
        same += list(exp) == list(row)
    # the builder restatement follows the summation order the reference's compiler produced (oracle.c:l2sq_refbuild): every row
    assert same == NB


def test_variants_C_and_E_golden(golden, orc):
    """tests/golden/ref_variants_ce.npz (tests/golden/make_golden_variants.py): what the REAL reference returned on the ref_small
    index for variant C (beam_search_with_pq, vamana_graph.py:535-605, live and with lazily deleted nodes, PQ and exact) and for the
    stochastic variant E (SearchEngineCorrect._pq_accelerated_graph_search, search_engine.py:398-506, under np.random.seed).  The
    oracle's literal restatements reproduce every list: ids in the reference's output order, distances bit-equal, E's visited /
    exact / PQ / step counts, and the generator position after each run."""
    g = golden
    z = np.load(ROOT / "tests" / "golden" / "ref_variants_ce.npz")
    nq = int(z["nq"])
    Q = g["Q"][:nq]
    luts = [orc.lut(g["codebook"], q) for q in Q]
    for tag, dead in (("live", np.zeros(g["N"], np.uint8)), ("del", z["deleted"])):
        for bw, k in z["shapes_c"]:
            for qi, q in enumerate(Q):
                e_ids = z[f"exp_C_{tag}_pq_bw{bw}_k{k}_ids"][qi]; e_d = z[f"exp_C_{tag}_pq_bw{bw}_k{k}_d"][qi]
                n = int((e_ids >= 0).sum())
                o = orc.beam_c(g["adj"], g["medoid"], int(bw), int(k), codes=g["codes"], lut_=luts[qi], dist_mode=orc.DIST_ADC_SEQ,
                               deleted=dead)
                assert list(o["ids"]) == list(e_ids[:n]) and np.array_equal(o["dists"], e_d[:n].astype(np.float32))
                e_ids = z[f"exp_C_{tag}_l2_bw{bw}_k{k}_ids"][qi]; e_d = z[f"exp_C_{tag}_l2_bw{bw}_k{k}_d"][qi]
                n = int((e_ids >= 0).sum())
                o = orc.beam_c(g["adj"], g["medoid"], int(bw), int(k), vec=g["vec"], q=q, dist_mode=orc.DIST_L2_SQ,
                               flavor=orc.FLAVOR_REFCC, deleted=dead, sqrt_out=False)
                assert list(o["ids"]) == list(e_ids[:n]) and np.array_equal(np.sqrt(o["dists"].astype(np.float64)), e_d[:n])
    gated = 0
    for seed, (L, k, bw) in enumerate(z["shapes_e"]):
        rng = orc.NumpyLegacyRandom(seed)
        for qi, q in enumerate(Q):
            ids, d2, st = orc.search_e(g["adj"], g["vec"], g["codes"], luts[qi], q, g["medoid"], int(L), int(k), int(bw), rng)
            e_ids = z[f"exp_E_seed{seed}_ids"][qi]
            n = int((e_ids >= 0).sum())
            assert list(ids) == list(e_ids[:n]) and np.array_equal(d2, z[f"exp_E_seed{seed}_d2"][qi][:n])
            assert [st["nodes_visited"], st["exact_distance_computations"], st["pq_distance_computations"], st["search_steps"]] \
                == list(z[f"exp_E_seed{seed}_stats"][qi])
            gated += st["pq_distance_computations"] - (st["exact_distance_computations"] - 1)
        assert rng.random() == float(z[f"exp_E_seed{seed}_next_draw"])
    assert gated > 0
