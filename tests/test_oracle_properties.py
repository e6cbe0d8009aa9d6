"""Oracle self-checks on adversarial small graphs (CPU, seconds): the C restatement (oracle/oracle.c) against a pure-Python
model of the reference's best-first search that uses CPython's own heapq — so tie behaviour comes from the real heap, not
from our reading of it — and the sorted-list formulation the GPU kernel runs against the two-heap form.

Model: greedy_search_cython, cython_utils.pyx:72-122 (frontier min-heap of (d, id), result max-heap of (-d, id) capped at
L, stop when the popped frontier distance exceeds the worst result, accept a first-seen neighbour when the result heap is
short or it beats the worst; output = result heap array stably sorted by distance), with compute_query_distance's ADC
(vamana_graph.py:301-329 -> fast_pq.py:320-328: sequential fp32 sum over the M subspaces)."""
import heapq

import numpy as np
import pytest


def adc_seq(lut, code):
    s = np.float32(0.0)
    for m in range(len(code)):
        s = np.float32(s + lut[m, code[m]])
    return float(s)


def model_best_first(adj, start, L, dist):
    seen = {start}
    d0 = dist(start)
    frontier, best = [(d0, start)], [(-d0, start)]
    hops = 0
    while frontier:
        d, cur = heapq.heappop(frontier)
        if d > -best[0][0]:
            break
        hops += 1
        for nb in adj[cur]:
            nb = int(nb)
            if nb in seen:
                continue
            seen.add(nb)
            dn = dist(nb)
            if len(best) < L or dn < -best[0][0]:
                heapq.heappush(frontier, (dn, nb))
                heapq.heappush(best, (-dn, nb))
                if len(best) > L:
                    heapq.heappop(best)
    out = sorted(best, key=lambda t: -t[0])                      # stable: ties keep the heap-array order
    return [i for _, i in out], [-nd for nd, _ in out], hops, len(seen)


def random_case(rng, N, R, M, levels, pad_frac):
    """Random digraph with 0-padded short rows, duplicate ids inside a row and self loops; ADC tables with few distinct
    values so that exact distance ties are the rule, not the exception."""
    adj = rng.integers(0, N, size=(N, R), dtype=np.uint32)
    short = rng.random(N) < pad_frac
    for i in np.flatnonzero(short):
        adj[i, rng.integers(1, R):] = 0                          # what DiskANNPersist.save_index pads with
    adj[rng.integers(0, N, N // 8), 0] = rng.integers(0, N, N // 8)   # a few arbitrary rewires
    adj[np.arange(0, N, 7), 1] = np.arange(0, N, 7)              # self loops
    codes = rng.integers(0, 256, size=(N, M), dtype=np.uint8)
    lut = rng.integers(0, levels, size=(M, 256)).astype(np.float32) * np.float32(0.25)
    return adj, codes, lut


CASES = [(60, 4, 2, 2, 0.5), (200, 8, 4, 3, 0.3), (500, 6, 3, 2, 0.2), (300, 16, 8, 4, 0.0), (120, 5, 1, 2, 0.6)]


@pytest.mark.parametrize("N,R,M,levels,pad", CASES)
def test_heap_form_is_the_reference_algorithm_under_ties(orc, N, R, M, levels, pad):
    rng = np.random.default_rng(N * 31 + R)
    adj, codes, lut = random_case(rng, N, R, M, levels, pad)
    for L in (1, 2, 7, 33, N + 5):
        for start in (0, int(rng.integers(0, N))):
            ids, ds, hops, vis = model_best_first(adj, start, L, lambda i: adc_seq(lut, codes[i]))
            h = orc.search_heap(adj, start, L, codes=codes, lut_=lut, dist_mode=orc.DIST_ADC_SEQ)
            assert [int(x) for x in h["ids"]] == ids, (L, start)              # the reference's own output order
            assert np.array_equal(h["dists"], np.array(ds, np.float32))
            assert (h["hops"], h["visited"]) == (hops, vis)


@pytest.mark.parametrize("N,R,M,levels,pad", CASES)
def test_list_form_equals_heap_form_under_ties(orc, N, R, M, levels, pad):
    """W = 1 with the strict-tie (ghost) rule is the same search: same list as a set of (d, id), same expansions, same
    number of distance evaluations — the property the GPU's reference-order mode is tested against."""
    rng = np.random.default_rng(N * 17 + M)
    adj, codes, lut = random_case(rng, N, R, M, levels, pad)
    for L in (1, 3, 10, 40):
        for start in (0, int(rng.integers(0, N))):
            h = orc.search_heap(adj, start, L, codes=codes, lut_=lut, dist_mode=orc.DIST_ADC_SEQ)
            l = orc.search_list(adj, start, L, codes=codes, lut_=lut, dist_mode=orc.DIST_ADC_SEQ, W=1, strict_ties=True)
            o = np.lexsort((h["ids"], h["dists"]))
            assert np.array_equal(h["ids"][o], l["ids"]) and np.array_equal(h["dists"][o], l["dists"]), (L, start)
            assert (h["hops"], h["visited"]) == (l["hops"], l["visited"]), (L, start)


def test_exact_mode_matches_the_model_on_duplicated_points(orc):
    """Variant B/D arithmetic (vamana_graph.py:607-640, 719-760) with exact duplicates in the data: squared-L2 ties."""
    rng = np.random.default_rng(5)
    N, D, R = 150, 12, 6
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[100:] = X[:50]
    adj, _, _ = random_case(rng, N, R, 1, 2, 0.3)
    q = rng.standard_normal(D).astype(np.float32)
    for L in (5, 25):
        h = orc.search_heap(adj, 3, L, vec=X, q=q, dist_mode=orc.DIST_L2_SQ, flavor=orc.FLAVOR_SEQ)
        ids, ds, hops, vis = model_best_first(adj, 3, L, lambda i: orc.l2sq(X[i], q, orc.FLAVOR_SEQ))
        assert [int(x) for x in h["ids"]] == ids and (h["hops"], h["visited"]) == (hops, vis)


def test_batch_entry_is_the_per_query_composition(orc):
    """orc_search_batch (the bench's CPU-baseline leg) = search_list + rerank per query, whatever the thread count."""
    rng = np.random.default_rng(9)
    N, D, R, M = 400, 16, 8, 4
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((9, D)).astype(np.float32)
    adj = rng.integers(0, N, size=(N, R), dtype=np.uint32)
    cb = rng.standard_normal((M, 256, D // M)).astype(np.float32)
    codes = orc.pq_encode(cb, X)
    for W in (1, 4):
        ids, d, hops, vis = orc.search_batch(adj, X, Q, 2, 30, 10, codes=codes, codebook=cb, dist_mode=orc.DIST_ADC_SEQ,
                                             flavor=orc.FLAVOR_WARP, W=W, rerank_=True, nthreads=3)
        for qi in range(Q.shape[0]):
            l = orc.search_list(adj, 2, 30, codes=codes, lut_=orc.lut(cb, Q[qi]), dist_mode=orc.DIST_ADC_SEQ, W=W, strict_ties=W == 1)
            oi, od = orc.rerank(X, Q[qi], l["ids"], 10, flavor=orc.FLAVOR_WARP)
            assert np.array_equal(ids[qi, :len(oi)], oi) and np.array_equal(d[qi, :len(od)], od)
            assert (hops[qi], vis[qi]) == (l["hops"], l["visited"])
