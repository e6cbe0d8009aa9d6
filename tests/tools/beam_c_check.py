"""Variant C with the reference's own semantics on the device (csrc/beam_c.cu through dr_beam_search_c) against the oracle's
literal restatement (oracle.c:orc_beam_c, itself pinned live against the real reference in tests/test_oracle_vs_reference.py).
Run by tests/test_beam_c_gpu.py in a child process (a first device run must not be able to take the shared suite's CUDA context
down with it); prints one JSON line and exits non-zero on the first mismatch.

  python tests/tools/beam_c_check.py            # all cases
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))

import oracle as orc                                   # noqa: E402  (checker)
from conftest import canon, make_case                  # noqa: E402
from diskrag_b200.engine import GpuIndex               # noqa: E402


def check_case(c, shapes, deleted_ids=()):
    n_checked = 0
    dead = np.zeros(c["N"], np.uint8)
    deleted_ids = [i for i in deleted_ids if i != c["medoid"]]       # the shims resolve a deleted start before the call
    dead[deleted_ids] = 1
    with GpuIndex.from_arrays(c["X"], c["adj"], c["codes"], c["codebook"], c["medoid"]) as idx:
        if deleted_ids:
            from diskrag_b200._lib import check, lib, ptr
            check(lib().dr_index_set_deleted(idx._h, ptr(dead)))
        for bw, k in shapes:
            for dist in ("pq", "exact"):
                r = idx.beam_search_c(c["Q"], k=k, beam_width=bw, dist=dist, sqrt_out=(dist == "pq"))
                for qi, q in enumerate(c["Q"]):
                    if dist == "pq":
                        o = orc.beam_c(c["adj"], c["medoid"], bw, k, codes=c["codes"], lut_=orc.lut(c["codebook"], q),
                                       dist_mode=orc.DIST_ADC_SEQ, deleted=dead, sqrt_out=True)
                    else:
                        o = orc.beam_c(c["adj"], c["medoid"], bw, k, vec=c["X"], q=q, dist_mode=orc.DIST_L2_SQ,
                                       flavor=orc.FLAVOR_WARP, deleted=dead, sqrt_out=False)
                    a = canon(o["ids"], o["dists"]); b = canon(r.ids[qi], r.dists[qi])
                    ok = (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])            # ids and distances bit-for-bit
                          and int(r.hops[qi]) == o["hops"] and int(r.visited[qi]) == o["visited"])
                    if not ok:
                        print(json.dumps({"ok": False, "bw": bw, "k": k, "dist": dist, "query": qi,
                                          "gpu": [r.ids[qi].tolist(), r.dists[qi].tolist(), int(r.hops[qi]), int(r.visited[qi])],
                                          "oracle": [o["ids"].tolist(), o["dists"].tolist(), o["hops"], o["visited"]]}))
                        sys.exit(1)
                    n_checked += 1
    return n_checked


def check_golden():
    """the device against what the REAL reference returned (tests/golden/ref_variants_ce.npz, ref_small index): PQ lists bit-equal
    (ids and sqrt'ed ADC distances), exact lists: same ids, distances within 1e-4 relative (BASELINE's tolerance)"""
    z = np.load(ROOT / "tests" / "golden" / "ref_small.npz"); v = np.load(ROOT / "tests" / "golden" / "ref_variants_ce.npz")
    N, D, R, med = int(z["N"]), int(z["D"]), int(z["R"]), int(z["medoid"])
    Q = z["Q"][:int(v["nq"])]
    n = 0
    from diskrag_b200._lib import check, lib, ptr
    with GpuIndex.from_records(z["records"], N, D, R, z["codes"], z["codebook"], med) as idx:
        for tag in ("live", "del"):
            check(lib().dr_index_set_deleted(idx._h, ptr(np.ascontiguousarray(v["deleted"])) if tag == "del" else None))
            for bw, k in v["shapes_c"]:
                r = idx.beam_search_c(Q, k=int(k), beam_width=int(bw), dist="pq", sqrt_out=True)
                x = idx.beam_search_c(Q, k=int(k), beam_width=int(bw), dist="exact", sqrt_out=True)
                for qi in range(len(Q)):
                    a = canon(v[f"exp_C_{tag}_pq_bw{bw}_k{k}_ids"][qi], v[f"exp_C_{tag}_pq_bw{bw}_k{k}_d"][qi].astype(np.float32))
                    b = canon(r.ids[qi], r.dists[qi])
                    e_ids = v[f"exp_C_{tag}_l2_bw{bw}_k{k}_ids"][qi]; e_d = v[f"exp_C_{tag}_l2_bw{bw}_k{k}_d"][qi]
                    m = int((e_ids >= 0).sum())
                    ok = (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                          and set(x.ids[qi][x.ids[qi] >= 0].tolist()) == set(e_ids[:m].tolist())
                          and np.allclose(np.sort(x.dists[qi][:m]), np.sort(e_d[:m]), rtol=1e-4))
                    if not ok:
                        print(json.dumps({"ok": False, "golden": True, "tag": tag, "bw": int(bw), "k": int(k), "query": qi,
                                          "gpu_pq": [r.ids[qi].tolist(), r.dists[qi].tolist()], "gpu_l2": [x.ids[qi].tolist(), x.dists[qi].tolist()]}))
                        sys.exit(1)
                    n += 1
    return n


def main():
    orc.build()
    shapes = [(5, 3), (8, 5), (2, 10), (16, 10), (1, 1), (64, 10), (0, 4)]
    n = 0
    n += check_case(make_case(orc, 2000, 64, 8, 16, 32, seed=5, nq=24, dup=40), shapes)                   # exact ties (duplicates)
    n += check_case(make_case(orc, 3000, 96, 24, 40, 48, seed=6, nq=16), shapes)                          # R > 32: two passes per row
    n += check_case(make_case(orc, 1500, 50, 10, 12, 24, seed=7, nq=16), shapes[:3],                      # D % 4 != 0, lazy deletes
                    deleted_ids=(1, 2, 3, 50, 51, 700, 1499))
    ng = check_golden()
    print(json.dumps({"ok": True, "queries_checked": n, "golden_lists_checked": ng}))


if __name__ == "__main__":
    main()
