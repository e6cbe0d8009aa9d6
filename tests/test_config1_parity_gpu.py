"""GPU: BASELINE configs[1] — 100k x 1536, the reference-equivalent graph in its COMPILED summation order
(tests/golden/config1_adj_100000.npz, adj_sha256 in profiles/r01m_config2_graph.json), written in the pydiskann/io
layout, loaded through the index directory and searched as ONE 10 000-query batch on the device.

Checked against the oracle's literal two-heap restatement of the reference (cython_utils.pyx:72-122,
vamana_graph.py:719-760) on the first queries of the batch:
  variant D (exact)   lists, hops, visited bit-equal in the GPU summation order; top-k id sets equal to the
                      reference-like (double-accumulated) order, distances within 1e-4 relative
  variant A (PQ)      ids, ADC distances, hops, visited bit-equal
  throughput (u8, W=8, rerank)   ids and distances bit-equal to its restatement
and reports how many top-10 id sets of the bench mode (u8 table from tensor cores, W=8) equal the reference's own
answer (variant A + exact rerank, search_engine.py:374-379).  The JSON summary is kept in profiles/."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests" / "tools"))
GRAPH = ROOT / "tests" / "golden" / "config1_adj_100000.npz"
ADJ_SHA256 = "6d78fd4877fad185fcfcd02113b119c68483a7bb5089c0c58654ee34efe952b6"


def test_committed_graph_is_the_compiled_order_build():
    z = np.load(GRAPH)
    assert (int(z["N"]), int(z["D"]), int(z["R"]), int(z["L"])) == (100_000, 1536, 32, 64)
    assert hashlib.sha256(np.ascontiguousarray(z["adj"]).tobytes()).hexdigest() == ADJ_SHA256


@pytest.mark.gpu
def test_config1_100k_batch_equals_the_reference_restatement():
    import parity_config2
    nq_oracle = 300
    out = parity_config2.run(10_000, nq_oracle)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "config1_parity_test.json").write_text(json.dumps(out, indent=1))
    assert out["N"] == 100_000
    full = f"{nq_oracle}/{nq_oracle}"
    D = out["variant_D_exact"]
    assert D["lists_bit_equal_to_oracle_gpu_order"] == full and D["hops_and_visited_equal"] == full
    # near-ties between the two summation orders may move an id across the top-k boundary: allow 1 %, distances 1e-4 relative
    assert int(D["top_k_id_sets_equal_to_reference_like_order"].split("/")[0]) >= 0.99 * nq_oracle
    assert D["max_rel_distance_diff_on_those"] <= 1e-4
    assert out["variant_A_pq"]["ids_adc_distances_hops_visited_bit_equal"] == full
    T = out["throughput_u8_W8_rerank"]
    assert T["top_k_ids_and_distances_bit_equal"] == full and T["hops_visited_equal"] == full
    Bm = out["bench_mode_vs_reference_answer"]
    assert int(Bm["gpu_variantA_rerank_sets_equal_to_oracle_composition"].split("/")[0]) >= 0.99 * nq_oracle
    # the bench mode visits nodes in a different order with an 8-bit table: its answer is the reference's on most queries and
    # nearly the same set on the rest (the fraction itself is the reported number, profiles/r02*_config1_parity.json)
    assert Bm["u8tc_W8_fraction"] >= 0.80 and Bm["mean_top10_overlap_u8tc_W8"] >= 0.97
    r = out["recall_at_10"]
    assert r["bench_mode_u8tc"] >= r["variant_A_rerank"] - 0.005
