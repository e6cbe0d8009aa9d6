"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/diskrag_b200.h declares,
the ctypes table covers them all, and compute calls fail loudly (no CPU fallback) without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    src = (ROOT / "include" / "diskrag_b200.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dr_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from diskrag_b200 import build_ext, _lib
    build_ext.build()
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/diskrag_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert L.dr_abi_version() == 2


def test_search_params_layout_matches_header():
    from diskrag_b200 import _lib
    src = (ROOT / "include" / "diskrag_b200.h").read_text()
    body = re.search(r"typedef struct dr_search_params \{(.*?)\} dr_search_params;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"int32_t\s+(\w+);", body)
    assert fields == [f[0] for f in _lib.SearchParams._fields_]
    assert C.sizeof(_lib.SearchParams) == 4 * len(fields)


def test_no_cpu_fallback_without_gpu():
    from diskrag_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    from diskrag_b200.engine import GpuIndex
    from diskrag_b200.pq.fast_pq import DiskANNPQ
    with pytest.raises(_lib.DiskragError):
        GpuIndex.from_arrays(np.zeros((4, 8), np.float32), np.zeros((4, 2), np.uint32))
    with pytest.raises(_lib.DiskragError):
        DiskANNPQ(2).fit(np.zeros((300, 8), np.float32))
    out = np.zeros(1, np.float32)
    rc = _lib.lib().dr_l2sq_batch(_lib.ptr(np.zeros((1, 8), np.float32)), _lib.ptr(np.zeros((1, 8), np.float32)), 1, 1, 8,
                                  _lib.ptr(out), 0)
    assert rc != 0 and b"no CPU fallback" in _lib.lib().dr_last_error()


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    for p in (ROOT / "diskrag_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "ref_loader" not in txt, p
    for p in (ROOT / "diskrag_b200" / "csrc").glob("*"):
        txt = p.read_text()
        assert "liboracle" not in txt and "dlopen" not in txt, p
        assert not [l for l in txt.splitlines() if l.strip().startswith("#include") and "oracle" in l], p


def test_only_the_checkers_touch_the_oracle():
    """oracle/ is test infrastructure: only tests/ (tests/tools/ included), __graft_entry__.py and bench.py's CPU-baseline /
    reference legs may import it; the tooling under scripts/ may compile it (building is not using) but never imports it."""
    for p in (ROOT / "scripts").glob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "ref_loader" not in txt and '"oracle"' not in txt, p
    bench = (ROOT / "bench.py").read_text()
    uses = [i for i, l in enumerate(bench.splitlines()) if "import oracle" in l or "import ref_loader" in l]
    legs = [i for i, l in enumerate(bench.splitlines()) if l.startswith("def cpu_baseline_port") or l.startswith("def run_reference")]
    assert uses and legs and min(uses) > min(legs), "bench.py imports the oracle outside its cpu_baseline / reference legs"
