"""Import the real reference (compiled into oracle/_ref by oracle/build_ref.py).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline
legs may import this module.
"""
import importlib
import os
import sys
from pathlib import Path

_REF = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (_REF / ".built").exists()


def load():
    """Return the reference modules as a namespace dict; raises RuntimeError if oracle/_ref is absent."""
    if not available():
        raise RuntimeError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
    os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
    p = str(_REF)
    if p not in sys.path:
        sys.path.insert(0, p)
    mods = {}
    for name in ("pydiskann.cython_utils", "pydiskann.vamana_graph", "pydiskann.pq.fast_pq",
                 "pydiskann.io.diskann_persist"):
        mods[name.split(".")[-1]] = importlib.import_module(name)
    return mods
