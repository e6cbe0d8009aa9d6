"""GPU: variant C with the reference's own semantics (dr_beam_search_c, csrc/beam_c.cu) == the oracle's literal restatement
(ids, distances bit-for-bit, hops, visited counts).  Runs in a child process with a timeout: a fault there fails this test
without poisoning the CUDA context of the rest of the suite."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
def test_beam_c_equals_the_oracle_restatement():
    p = subprocess.run([sys.executable, str(ROOT / "tests" / "tools" / "beam_c_check.py")], capture_output=True, text=True,
                       timeout=300, cwd=str(ROOT))
    last = (p.stdout.strip().splitlines() or ["{}"])[-1]
    assert p.returncode == 0, (last, p.stderr[-2000:])
    assert json.loads(last)["ok"] and json.loads(last)["queries_checked"] > 500
