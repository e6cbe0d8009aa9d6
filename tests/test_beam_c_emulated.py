"""CPU: the TEXT of csrc/beam_c.cu's kernel, executed without a GPU by a lockstep warp emulator (tests/tools/warp_emul.h: one OS thread
per lane, every warp collective an exchange between two barriers), against the oracle's literal restatement of the reference's
beam_search_with_pq (oracle.c:orc_beam_c, pinned to the real reference).  The kernel body and the device helpers it uses are cut out
of the .cu / .cuh files at test time, so what runs here is what nvcc compiles; what this cannot show is anything that depends on
the GPU's memory system or scheduler; the device run is tests/test_beam_c_gpu.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, canon, make_case

TOOLS = ROOT / "tests" / "tools"
BUILD = TOOLS / "build"


@pytest.fixture(scope="module")
def emul():
    cm = (ROOT / "diskrag_b200" / "csrc" / "common.cuh").read_text()
    bc = (ROOT / "diskrag_b200" / "csrc" / "beam_c.cu").read_text()
    helpers = cm[cm.index("#define DR_FULL"):cm.index("// Canonical cosine DISTANCE")]       # f2ord ... warp_l2sq
    kernel = bc[bc.index("struct BeamCArgs"):bc.index("// number of resident CTAs")]          # args, ADC helper, the kernel
    BUILD.mkdir(exist_ok=True)
    (BUILD / "beam_c_kernel.inc").write_text(helpers + "\n" + kernel)
    so = BUILD / "beam_c_emul.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-I", str(BUILD), "-I", str(TOOLS),
                           str(TOOLS / "beam_c_emul_main.cpp"), "-o", str(so)])
    return C.CDLL(str(so))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def run_emulated(emul, c, Q, luts, k, bw, dist, dead, sqrt_out):
    B = Q.shape[0]
    ids = np.empty((B, k), np.int32); dd = np.empty((B, k), np.float32); hops = np.empty(B, np.int32); vis = np.empty(B, np.int32)
    rc = emul.emul_beam_c(_p(c["X"]), _p(c["adj"]), _p(c["codes"]), _p(dead), _p(Q), _p(luts), C.c_longlong(c["N"]), C.c_int(c["D"]),
                          C.c_int(c["R"]), C.c_int(c["M"]), C.c_longlong(B), C.c_int(k), C.c_int(bw), C.c_int(0 if dist == "pq" else 1),
                          C.c_int(int(sqrt_out)), C.c_uint32(c["medoid"]), _p(ids), _p(dd), _p(hops), _p(vis))
    assert rc == 0
    return ids, dd, hops, vis


@pytest.mark.parametrize("shape", [(1500, 32, 4, 8, 16, 1, 40), (1200, 96, 12, 40, 48, 2, 0), (900, 30, 5, 12, 24, 5, 0)],
                         ids=["ties", "R40_two_passes", "D_not_multiple_of_4"])
def test_kernel_text_equals_the_oracle(emul, orc, shape):
    N, D, M, R, Lb, seed, dup = shape
    c = make_case(orc, N, D, M, R, Lb, seed, nq=6, dup=dup)
    c["X"] = np.ascontiguousarray(c["X"], np.float32); c["adj"] = np.ascontiguousarray(c["adj"], np.uint32)
    c["codes"] = np.ascontiguousarray(c["codes"], np.uint8)
    Q = np.ascontiguousarray(c["Q"], np.float32)
    luts = np.ascontiguousarray(np.stack([orc.lut(c["codebook"], q) for q in Q]), np.float32)
    rng = np.random.default_rng(seed)
    dead = np.zeros(N, np.uint8); dead[rng.choice(N, N // 20, replace=False)] = 1; dead[c["medoid"]] = 0
    for deleted in (None, dead):
        for bw, k in [(5, 3), (2, 10), (16, 10), (1, 1), (0, 4)]:
            for dist in ("pq", "exact"):
                ids, dd, hops, vis = run_emulated(emul, c, Q, luts, k, bw, dist, deleted, sqrt_out=(dist == "pq"))
                for qi, q in enumerate(Q):
                    if dist == "pq":
                        o = orc.beam_c(c["adj"], c["medoid"], bw, k, codes=c["codes"], lut_=luts[qi], dist_mode=orc.DIST_ADC_SEQ,
                                       deleted=deleted, sqrt_out=True)
                    else:
                        o = orc.beam_c(c["adj"], c["medoid"], bw, k, vec=c["X"], q=q, dist_mode=orc.DIST_L2_SQ, flavor=orc.FLAVOR_WARP,
                                       deleted=deleted, sqrt_out=False)
                    a = canon(o["ids"], o["dists"]); b = canon(ids[qi], dd[qi])
                    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (bw, k, dist, qi)     # ids, distances: bits
                    assert (int(hops[qi]), int(vis[qi])) == (o["hops"], o["visited"]), (bw, k, dist, qi)
                    assert (ids[qi, len(o["ids"]):] == -1).all()


def test_kernel_text_equals_the_real_reference_outputs(emul, golden, orc):
    """The same kernel text against what the REAL reference's beam_search_with_pq returned (tests/golden/ref_variants_ce.npz, made by
    tests/golden/make_golden_variants.py): PQ lists bit-equal (ids, sqrt'ed ADC distances), live and with lazily deleted nodes; exact
    lists: same ids, distances within 1e-4 relative (the device sums in warp order, the reference in its compiled order)."""
    g = golden
    v = np.load(ROOT / "tests" / "golden" / "ref_variants_ce.npz")
    Q = np.ascontiguousarray(g["Q"][:int(v["nq"])], np.float32)
    c = dict(X=np.ascontiguousarray(g["vec"], np.float32), adj=np.ascontiguousarray(g["adj"], np.uint32),
             codes=np.ascontiguousarray(g["codes"], np.uint8), N=g["N"], D=g["D"], R=g["R"], M=g["M"], medoid=g["medoid"])
    luts = np.ascontiguousarray(np.stack([orc.lut(g["codebook"], q) for q in Q]), np.float32)
    for tag, dead in (("live", None), ("del", np.ascontiguousarray(v["deleted"]))):
        for bw, k in v["shapes_c"]:
            bw, k = int(bw), int(k)
            ids, dd, _, _ = run_emulated(emul, c, Q, luts, k, bw, "pq", dead, sqrt_out=True)
            xi, xd, _, _ = run_emulated(emul, c, Q, luts, k, bw, "exact", dead, sqrt_out=True)
            for qi in range(len(Q)):
                a = canon(v[f"exp_C_{tag}_pq_bw{bw}_k{k}_ids"][qi], v[f"exp_C_{tag}_pq_bw{bw}_k{k}_d"][qi].astype(np.float32))
                b = canon(ids[qi], dd[qi])
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (tag, bw, k, qi)
                e_ids = v[f"exp_C_{tag}_l2_bw{bw}_k{k}_ids"][qi]; e_d = v[f"exp_C_{tag}_l2_bw{bw}_k{k}_d"][qi]
                m = int((e_ids >= 0).sum())
                assert set(xi[qi][xi[qi] >= 0].tolist()) == set(e_ids[:m].tolist()), (tag, bw, k, qi)
                np.testing.assert_allclose(np.sort(xd[qi][:m]), np.sort(e_d[:m]), rtol=1e-4)


def test_kernel_text_on_adversarial_rows(emul, orc):
    """Rows full of repeated ids (inside one 32-wide pass and across the two passes of an R = 40 row), self loops, 0-padding, two-byte
    codes (massive exact ADC ties), duplicated points: first-occurrence claiming (match_any + atomicOr) and the tie rules of the
    eviction must follow the reference's sequential scan exactly."""
    rng = np.random.default_rng(77)
    N, D, M, R = 96, 8, 2, 40
    X = rng.standard_normal((N, D)).astype(np.float32)
    X[N // 2:] = X[:N // 2]                                             # every point has an exact duplicate
    cb = rng.standard_normal((M, 256, D // M)).astype(np.float32)
    codes = rng.integers(0, 3, (N, M)).astype(np.uint8)                 # 9 distinct codes: ADC ties everywhere
    adj = rng.integers(0, N, (N, R)).astype(np.uint32)
    adj[:, 5:9] = adj[:, 4:5]                                           # repeats inside the first pass
    adj[:, 33:36] = adj[:, 2:3]                                         # repeats across passes
    adj[np.arange(N), 10] = np.arange(N)                                # self loops
    adj[::3, 20:] = 0                                                   # 0-padded tails
    Q = rng.standard_normal((8, D)).astype(np.float32)
    c = dict(X=X, adj=adj, codes=codes, N=N, D=D, R=R, M=M, medoid=7)
    luts = np.ascontiguousarray(np.stack([orc.lut(cb, q) for q in Q]), np.float32)
    dead = np.zeros(N, np.uint8); dead[[1, 2, 30, 31, 90]] = 1
    for deleted in (None, dead):
        for bw, k in [(1, 1), (3, 2), (5, 3), (8, 12), (40, 20), (0, 2)]:
            for dist in ("pq", "exact"):
                ids, dd, hops, vis = run_emulated(emul, c, Q, luts, k, bw, dist, deleted, sqrt_out=False)
                for qi, q in enumerate(Q):
                    kw = (dict(codes=codes, lut_=luts[qi], dist_mode=orc.DIST_ADC_SEQ) if dist == "pq" else
                          dict(vec=X, q=q, dist_mode=orc.DIST_L2_SQ, flavor=orc.FLAVOR_WARP))
                    o = orc.beam_c(adj, 7, bw, k, deleted=deleted, sqrt_out=False, **kw)
                    a = canon(o["ids"], o["dists"]); b = canon(ids[qi], dd[qi])
                    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (bw, k, dist, qi, a, b)
                    assert (int(hops[qi]), int(vis[qi])) == (o["hops"], o["visited"]), (bw, k, dist, qi)
