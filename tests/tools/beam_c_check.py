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


def main():
    orc.build()
    shapes = [(5, 3), (8, 5), (2, 10), (16, 10), (1, 1), (64, 10), (0, 4)]
    n = 0
    n += check_case(make_case(orc, 2000, 64, 8, 16, 32, seed=5, nq=24, dup=40), shapes)                   # exact ties (duplicates)
    n += check_case(make_case(orc, 3000, 96, 24, 40, 48, seed=6, nq=16), shapes)                          # R > 32: two passes per row
    n += check_case(make_case(orc, 1500, 50, 10, 12, 24, seed=7, nq=16), shapes[:3],                      # D % 4 != 0, lazy deletes
                    deleted_ids=(1, 2, 3, 50, 51, 700, 1499))
    print(json.dumps({"ok": True, "queries_checked": n}))


if __name__ == "__main__":
    main()
