"""Import the real reference (compiled into oracle/_ref by oracle/build_ref.py).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline
legs may import this module.

oracle/_ref/pydiskann holds cython_utils.<abi>.so plus sourceless bytecode (*.pycbin) of the pure-python
modules; a meta-path finder maps `pydiskann[.sub].mod` onto those files.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
from pathlib import Path

_REF = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (_REF / ".built").exists() and (_REF / "pydiskann" / "vamana_graph.pycbin").exists()


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "pydiskann" and not fullname.startswith("pydiskann."):
            return None
        rel = Path(*fullname.split("."))
        pkg_init = _REF / rel / "__init__.pycbin"
        if pkg_init.exists():
            loader = importlib.machinery.SourcelessFileLoader(fullname, str(pkg_init))
            return importlib.util.spec_from_file_location(fullname, str(pkg_init), loader=loader,
                                                          submodule_search_locations=[str(_REF / rel)])
        mod = (_REF / rel).with_suffix(".pycbin")
        if mod.exists():
            loader = importlib.machinery.SourcelessFileLoader(fullname, str(mod))
            return importlib.util.spec_from_file_location(fullname, str(mod), loader=loader)
        return None  # extension modules (cython_utils.so) are found by the normal path finder via __path__


def load():
    """Return the reference modules as a namespace dict; raises RuntimeError if oracle/_ref is absent."""
    if not available():
        raise RuntimeError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
    os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    mods = {}
    for name in ("pydiskann.cython_utils", "pydiskann.vamana_graph", "pydiskann.pq.fast_pq",
                 "pydiskann.io.diskann_persist"):
        mods[name.split(".")[-1]] = importlib.import_module(name)
    return mods


def load_dataset_benchmark():
    """The reference's dataset_benchmark.py (its run_benchmark(args) prints the recall / latency / QPS table), or None."""
    f = _REF / "dataset_benchmark.pycbin"
    if not f.exists():
        return None
    load()                                   # pydiskann must resolve to oracle/_ref first
    loader = importlib.machinery.SourcelessFileLoader("_reference_dataset_benchmark", str(f))
    spec = importlib.util.spec_from_file_location("_reference_dataset_benchmark", str(f), loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod
