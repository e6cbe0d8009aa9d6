"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY — the checker, never the product.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; nothing under
diskrag_b200/ does.  Parity status: pinned against the real reference (oracle/_ref) and the golden
vectors under tests/golden/ by tests/test_oracle_vs_reference.py (live) and tests/test_golden_oracle.py and tests/test_golden_config0.py (committed vectors; the latter at BASELINE configs[0] scale).
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

DIST_ADC_SEQ, DIST_L2_SQRT, DIST_L2_SQ, DIST_ADC_TREE, DIST_ADC_U8, DIST_COSINE = 0, 1, 2, 3, 4, 5
FLAVOR_DOUBLE, FLAVOR_NUMPY, FLAVOR_WARP, FLAVOR_SEQ, FLAVOR_REFCC = 0, 1, 2, 3, 4


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
                               str(src), "-o", str(so), "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_l2sq_seq.restype = C.c_float
        _LIB.orc_l2sq.restype = C.c_float
        _LIB.orc_l2sq_numpy.restype = C.c_float
        _LIB.orc_l2sq_warp.restype = C.c_float
        _LIB.orc_np_sum_f32.restype = C.c_float
        _LIB.orc_cosine_dist.restype = C.c_double
        _LIB.orc_pq_sdc.restype = C.c_float
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def l2sq(x, y, flavor=FLAVOR_SEQ):
    x, y = _f32(x), _f32(y)
    return float(lib().orc_l2sq(_p(x), _p(y), C.c_int(x.size), C.c_int(flavor)))


def cosine_dist(x, y):
    x, y = _f32(x), _f32(y)
    return float(lib().orc_cosine_dist(_p(x), _p(y), C.c_int(x.size)))


def np_sum_f32(a):
    a = _f32(a)
    return float(lib().orc_np_sum_f32(_p(a), C.c_long(a.size)))


def lut(codebook, q):
    """codebook f32[M,256,ds], q f32[D] -> f32[M,256] (fast_pq.py:294-318)."""
    codebook, q = _f32(codebook), _f32(q)
    M, _, ds = codebook.shape
    out = np.empty((M, 256), np.float32)
    lib().orc_lut(_p(codebook), _p(q), C.c_int(M), C.c_int(ds), _p(out))
    return out


def lut_u8(codebook, q):
    """Throughput-mode table: -> (u8[M,256], scale, offset); d ~ offset + scale * sum."""
    codebook, q = _f32(codebook), _f32(q)
    M, _, ds = codebook.shape
    out = np.empty((M, 256), np.uint8)
    sc = C.c_float(0); off = C.c_float(0)
    lib().orc_lut_u8(_p(codebook), _p(q), C.c_int(M), C.c_int(ds), _p(out), C.byref(sc), C.byref(off))
    return out, float(sc.value), float(off.value)


def adc(codes, lut_, tree=False):
    codes = np.ascontiguousarray(codes, np.uint8)
    lut_ = _f32(lut_)
    n, M = codes.shape
    out = np.empty(n, np.float32)
    lib().orc_adc(_p(codes), _p(lut_), C.c_long(n), C.c_int(M), C.c_int(int(tree)), _p(out))
    return out


def pq_encode(codebook, X):
    codebook, X = _f32(codebook), _f32(X)
    M, _, ds = codebook.shape
    N, D = X.shape
    out = np.empty((N, M), np.uint8)
    lib().orc_pq_encode(_p(codebook), _p(X), C.c_long(N), C.c_int(D), C.c_int(M), _p(out))
    return out


def pq_sdc(codebook, c1, c2):
    codebook = _f32(codebook)
    c1 = np.ascontiguousarray(c1, np.uint8); c2 = np.ascontiguousarray(c2, np.uint8)
    M, _, ds = codebook.shape
    return float(lib().orc_pq_sdc(_p(codebook), _p(c1), _p(c2), C.c_int(M), C.c_int(ds)))


def _search(fn_name, adj, codes, lut_, vec, q, flavor, dist_mode, extra, start, L, trace, deleted=None):
    adj = np.ascontiguousarray(adj, np.uint32)
    N, R = adj.shape
    M = 0
    if codes is not None:
        codes = np.ascontiguousarray(codes, np.uint8); M = codes.shape[1]
        lut_ = np.ascontiguousarray(lut_, np.uint8) if dist_mode == DIST_ADC_U8 else _f32(lut_)
    D = 0
    if vec is not None:
        vec = _f32(vec); D = vec.shape[1]
    if q is not None:
        q = _f32(q)
    ids = np.full(L + 1, -1, np.int32); d = np.full(L + 1, np.inf, np.float32)
    hops = C.c_int32(0); nvis = C.c_int32(0)
    tr = np.full(trace, -1, np.int32) if trace else None
    tail = ()
    if deleted is not None:                     # is_deleted flags (cython_utils.pyx:100-109, 120); the caller resolves a deleted start
        deleted = np.ascontiguousarray(deleted, np.uint8)
        fn_name += "_del"; tail = (_p(deleted),)
    n = getattr(lib(), fn_name)(
        _p(adj), C.c_int(R), C.c_long(N), _p(codes), C.c_int(M), _p(lut_),
        _p(vec), C.c_int(D), _p(q), C.c_int(flavor), C.c_int(dist_mode), *extra,
        C.c_int(int(start)), C.c_int(L), _p(ids), _p(d), C.byref(hops), C.byref(nvis),
        _p(tr), C.c_int(trace), *tail)
    res = {"ids": ids[:n].copy(), "dists": d[:n].copy(), "hops": hops.value, "visited": nvis.value}
    if trace:
        res["trace"] = tr[:min(trace, nvis.value)].copy()
    return res


def search_heap(adj, start, L, *, codes=None, lut_=None, vec=None, q=None, dist_mode=DIST_ADC_SEQ,
                flavor=FLAVOR_DOUBLE, truncate_frontier=False, trace=0, deleted=None):
    """Literal two-heap form of variants A/B/D (cython_utils.pyx:72-122, vamana_graph.py:607-640,719-760)."""
    return _search("orc_search_heap", adj, codes, lut_, vec, q, flavor, dist_mode,
                   (C.c_int(int(truncate_frontier)),), start, L, trace, deleted)


def search_list(adj, start, L, *, codes=None, lut_=None, vec=None, q=None, dist_mode=DIST_ADC_SEQ,
                flavor=FLAVOR_WARP, W=1, strict_ties=True, trace=0, deleted=None, w_after_empty=0):
    """Sorted-L-list form with W expansions per step: the exact statement of the GPU kernel.  w_after_empty > W (throughput mode,
    strict_ties = False): a step that follows a step without survivors expands up to that many entries (dr_search_params)."""
    lib().orc_set_w_after_empty(C.c_int(int(w_after_empty)))
    try:
        return _search("orc_search_list", adj, codes, lut_, vec, q, flavor, dist_mode,
                       (C.c_int(W), C.c_int(int(strict_ties))), start, L, trace, deleted)
    finally:
        lib().orc_set_w_after_empty(C.c_int(0))


def beam_c(adj, start, beam_width, k, *, codes=None, lut_=None, vec=None, q=None, dist_mode=DIST_ADC_SEQ,
           flavor=FLAVOR_REFCC, deleted=None, sqrt_out=True, trace=0):
    """Variant C, literally: beam_search_with_pq / beam_search (vamana_graph.py:535-605, 690-717) with the reference's
    inverted frontier truncation; rows scanned in stored order."""
    adj = np.ascontiguousarray(adj, np.uint32)
    N, R = adj.shape
    M = D = 0
    if codes is not None:
        codes = np.ascontiguousarray(codes, np.uint8); M = codes.shape[1]; lut_ = _f32(lut_)
    if vec is not None:
        vec = _f32(vec); D = vec.shape[1]
    if q is not None:
        q = _f32(q)
    if deleted is not None:
        deleted = np.ascontiguousarray(deleted, np.uint8)
    ids = np.full(k + 1, -1, np.int32); d = np.full(k + 1, np.inf, np.float32)
    hops = C.c_int32(0); nvis = C.c_int32(0)
    tr = np.full(trace, -1, np.int32) if trace else None
    n = lib().orc_beam_c(_p(adj), C.c_int(R), C.c_long(N), _p(codes), C.c_int(M), _p(lut_), _p(vec), C.c_int(D), _p(q),
                         C.c_int(flavor), C.c_int(dist_mode), _p(deleted), C.c_int(int(sqrt_out)), C.c_int(int(start)),
                         C.c_int(int(beam_width)), C.c_int(int(k)), _p(ids), _p(d), C.byref(hops), C.byref(nvis),
                         _p(tr), C.c_int(trace))
    res = {"ids": ids[:n].copy(), "dists": d[:n].copy(), "hops": hops.value, "visited": nvis.value}
    if trace:
        res["trace"] = tr[:min(trace, nvis.value)].copy()
    return res


class NumpyLegacyRandom:
    """np.random.seed(int) + np.random.random(), restated (MT19937 init_genrand / genrand_res53) for variant E."""

    def __init__(self, seed):
        self.state = np.zeros(625, np.uint32)
        lib().orc_mt_seed(_p(self.state), C.c_uint32(int(seed) & 0xFFFFFFFF))

    def random(self):
        lib().orc_mt_random.restype = C.c_double
        return float(lib().orc_mt_random(_p(self.state)))


def search_e(adj, vec, codes, lut_, q, start, L, k, beam_width=None, rng=None):
    """Variant E, literally: SearchEngineCorrect._pq_accelerated_graph_search (search_engine.py:398-506), stochastic gate
    included; rng = NumpyLegacyRandom(seed) plays np.random.seed(seed).  -> ids, exact squared distances, stats dict."""
    adj = np.ascontiguousarray(adj, np.uint32); vec = _f32(vec); q = _f32(q); lut_ = _f32(lut_)
    codes = np.ascontiguousarray(codes, np.uint8)
    N, R = adj.shape
    rng = rng or NumpyLegacyRandom(0)
    ids = np.full(k, -1, np.int32); d = np.full(k, np.inf, np.float32); st = np.zeros(4, np.int32)
    n = lib().orc_search_e(_p(adj), C.c_int(R), C.c_long(N), _p(codes), C.c_int(codes.shape[1]), _p(lut_), _p(vec),
                           C.c_int(vec.shape[1]), _p(q), C.c_int(int(start)), C.c_int(int(L)), C.c_int(int(k)),
                           C.c_int(int(beam_width or 0)), _p(rng.state), _p(ids), _p(d), _p(st))
    return ids[:n].copy(), d[:n].copy(), {"nodes_visited": int(st[0]), "exact_distance_computations": int(st[1]),
                                          "pq_distance_computations": int(st[2]), "search_steps": int(st[3])}


def rerank(vec, q, ids, k, flavor=FLAVOR_WARP):
    vec, q = _f32(vec), _f32(q)
    ids = np.ascontiguousarray(ids, np.int32)
    oi = np.full(k, -1, np.int32); od = np.full(k, np.inf, np.float32)
    m = lib().orc_rerank(_p(vec), C.c_int(vec.shape[1]), _p(q), C.c_int(flavor), _p(ids), C.c_int(ids.size),
                         C.c_int(k), _p(oi), _p(od))
    return oi[:m], od[:m]


def search_batch(adj, vec, Q, start, L, k, *, codes=None, codebook=None, dist_mode=DIST_ADC_SEQ,
                 flavor=FLAVOR_WARP, W=1, rerank_=True, nthreads=0, w_after_empty=0):
    adj = np.ascontiguousarray(adj, np.uint32); vec = _f32(vec); Q = _f32(Q)
    N, R = adj.shape; B, D = Q.shape
    M = 0
    if codes is not None:
        codes = np.ascontiguousarray(codes, np.uint8); M = codes.shape[1]; codebook = _f32(codebook)
    ids = np.empty((B, k), np.int32); d = np.empty((B, k), np.float32)
    hops = np.empty(B, np.int32); nvis = np.empty(B, np.int32)
    lib().orc_set_w_after_empty(C.c_int(int(w_after_empty)))      # read-only inside the (OpenMP) batch
    lib().orc_search_batch(_p(adj), C.c_int(R), C.c_long(N), _p(codes), C.c_int(M), _p(codebook),
                           _p(vec), C.c_int(D), _p(Q), C.c_long(B), C.c_int(dist_mode), C.c_int(flavor),
                           C.c_int(W), C.c_int(int(start)), C.c_int(L), C.c_int(k), C.c_int(int(rerank_)),
                           _p(ids), _p(d), _p(hops), _p(nvis), C.c_int(nthreads))
    lib().orc_set_w_after_empty(C.c_int(0))
    return ids, d, hops, nvis


def vamana_build(P, R, L, alpha, medoid, sigma0, sigma1):
    """Sequential 2-pass Vamana (cython_utils.pyx:269-492) -> list of rows (python lists)."""
    P = _f32(P); N, D = P.shape
    s0 = np.ascontiguousarray(sigma0, np.int32); s1 = np.ascontiguousarray(sigma1, np.int32)
    cap = R + 1
    adj = np.empty((N, cap), np.int32); deg = np.empty(N, np.int32)
    ov = lib().orc_vamana_build(_p(P), C.c_long(N), C.c_int(D), C.c_int(R), C.c_int(L), C.c_float(alpha),
                                C.c_int(int(medoid)), _p(s0), _p(s1), _p(adj), C.c_int(cap), _p(deg))
    assert ov == 0
    return [adj[i, :deg[i]].tolist() for i in range(N)]


def medoid(P, samples, skip_self=False):
    P = _f32(P); s = np.ascontiguousarray(samples, np.int32)
    return int(lib().orc_medoid(_p(P), C.c_long(P.shape[0]), C.c_int(P.shape[1]), _p(s), C.c_int(s.size),
                                C.c_int(int(skip_self))))


def ground_truth(X, Q, k):
    X, Q = _f32(X), _f32(Q)
    out = np.empty((Q.shape[0], k), np.int32)
    lib().orc_ground_truth(_p(X), C.c_long(X.shape[0]), C.c_int(X.shape[1]), _p(Q), C.c_long(Q.shape[0]),
                           C.c_int(k), _p(out))
    return out


def num_threads():
    return int(lib().orc_num_threads())


class InMemGraph:
    """Restatement of the reference's LIVE in-memory graph and its dynamic updates (VamanaGraphWithPQ.insert_node / delete_node,
    pydiskann/vamana_graph.py:58-125), PQ search off.  Neighbour sets are real Python sets, so their iteration order — the order
    greedy_search_cython scans a row in (cython_utils.pyx:108) — is CPython's own; the searches see true degrees (rows padded with
    0xFFFFFFFF, which every search here and on the GPU skips as "no neighbour").

      insert_node(idx, v):  cands = greedy_search_cython(graph, medoid, v, L = 2R, exact)          (:96-99, cython_utils.pyx:72-122)
                            neighbours = the R nearest live candidates by (l2_distance_fast, id): robust_prune_cython's removal loop
                            rebinds the list it iterates over, so nothing is ever pruned (cython_utils.pyx:147-165), and the
                            metric string sits in compute_distance's query_vector slot, leaving its metric at 'l2' (:273)
                            reverse edges added without a re-prune (:106-114): rows outgrow R
      delete_node(idx):     lazy flag (:116-125)
    Distances use l2_distance_fast_cython's compiled summation order (FLAVOR_REFCC).  Pinned against the real reference by
    tests/test_golden_inmem.py (committed outputs of a 1000-insert script) and tests/test_oracle_vs_reference.py (live)."""

    def __init__(self, vec, rows, deg, medoid, R):
        self.vec = [np.ascontiguousarray(v, np.float32) for v in vec]
        self.nbrs = [set() for _ in range(len(self.vec))]
        for i in range(len(self.vec)):
            for x in rows[i, :deg[i]]:
                self.nbrs[i].add(int(x))
        self.deleted = [False] * len(self.vec)
        self.medoid, self.R = int(medoid), int(R)
        self._rows = None

    def rows(self):
        n = len(self.vec)
        w = max(1, max(len(s) for s in self.nbrs))
        if self._rows is None or self._rows.shape[0] < n or self._rows.shape[1] < w:
            self._rows = np.full((n + 256, max(w, 16) * 2), 0xFFFFFFFF, np.uint32)
            self._stale = set(range(n))
        for i in self._stale:
            nb = list(self.nbrs[i])
            self._rows[i, :len(nb)] = nb
            self._rows[i, len(nb):] = 0xFFFFFFFF
        self._stale = set()
        return self._rows[:n]

    def _touch(self, i):
        if self._rows is not None:
            self._stale.add(i)

    def search(self, q, L, start=None, flavor=FLAVOR_REFCC):
        start = self.medoid if start is None else start
        dead = np.array(self.deleted, np.uint8)
        if dead[start]:                                            # cython_utils.pyx:84-90
            live = np.flatnonzero(dead == 0)
            if live.size == 0:
                return []
            start = int(live[0])
        r = search_heap(self.rows(), start, L, vec=np.stack(self.vec), q=q, dist_mode=DIST_L2_SQ, flavor=flavor,
                        deleted=dead if dead.any() else None)
        return [int(x) for x in r["ids"]]

    def insert_node(self, idx, v, L_insert=None):
        assert idx == len(self.vec), "new ids are dense"
        v = np.ascontiguousarray(v, np.float32)
        self.vec.append(v); self.nbrs.append(set()); self.deleted.append(False)
        if self._rows is not None:
            self._stale.add(idx)
        if len(self.vec) == 1:
            self.medoid = idx
            return
        cands = self.search(v, L_insert if L_insert is not None else 2 * self.R)
        scored = sorted((l2sq(self.vec[idx], self.vec[c], FLAVOR_REFCC), c) for c in set(cands) if not self.deleted[c])
        new = set()
        for _, c in scored:
            if len(new) >= self.R:
                break
            new.add(c)
        self.nbrs[idx] = new
        self._touch(idx)
        for nb in list(new):
            if not self.deleted[nb] and nb != idx:
                self.nbrs[nb].add(idx)
                self._touch(nb)

    def delete_node(self, idx):
        self.deleted[idx] = True
