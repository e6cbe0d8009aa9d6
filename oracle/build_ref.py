#!/usr/bin/env python3
"""Build the REAL reference (pydiskann) into oracle/_ref/ as compiled artefacts only.

TEST INFRASTRUCTURE.  Nothing under diskrag_b200/ may import this or its outputs.

Recipe (no reference source is copied into the repo; oracle/_ref/ is git-ignored and holds
binaries only, so it travels to the GPU box with the gpurun snapshot):

  1. pydiskann/cython_utils.pyx  --cython-->  C++  --g++ -O3 -ffast-math-->  cython_utils.<abi>.so
     (same flags as /root/reference/pydiskann/setup.py:10; we call cython and g++ directly and do
     not run the reference's setup.py)
  2. every pydiskann/**/*.py     --py_compile-->  sourceless bytecode next to it, stored as *.pycbin
     (vamana_graph, pq/fast_pq, pq/adaptive_pq, io/diskann_persist and the package __init__s; the
     extension is not .pyc because the gpurun snapshot drops *.pyc files — ref_loader.py installs an
     import hook that loads *.pycbin with importlib's SourcelessFileLoader)

Usage:  python oracle/build_ref.py [--ref /root/reference] [--force]
Import: oracle/ref_loader.py puts oracle/_ref first on sys.path (and sets NUMBA_DISABLE_JIT=1,
        because the dead first class in pq/fast_pq.py decorates with numba cache=True which needs
        the .py source next to the bytecode).
"""
import argparse
import os
import py_compile
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"


def build(ref: Path, force: bool = False) -> bool:
    src_pkg = ref / "pydiskann"
    if not src_pkg.is_dir():
        return False
    dst_pkg = OUT / "pydiskann"
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    so = dst_pkg / f"cython_utils{ext_suffix}"
    stamp = OUT / ".built"
    if stamp.exists() and so.exists() and (OUT / "dataset_benchmark.pycbin").exists() and not force:
        return True
    (OUT / "build").mkdir(parents=True, exist_ok=True)
    dst_pkg.mkdir(parents=True, exist_ok=True)

    # 1. the Cython translation unit -> .so
    import numpy
    cpp = OUT / "build" / "cython_utils.cpp"
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3",
                           str(src_pkg / "cython_utils.pyx"), "-o", str(cpp)])
    inc = [f"-I{sysconfig.get_paths()['include']}", f"-I{numpy.get_include()}"]
    subprocess.check_call(["g++", "-shared", "-fPIC", "-O3", "-ffast-math", "-w",
                           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", *inc,
                           str(cpp), "-o", str(so)])

    # 2. the pure-python modules -> sourceless bytecode
    for py in sorted(src_pkg.rglob("*.py")):
        rel = py.relative_to(src_pkg)
        if rel.name == "setup.py":
            continue
        dst = (dst_pkg / rel).with_suffix(".pycbin")
        dst.parent.mkdir(parents=True, exist_ok=True)
        py_compile.compile(str(py), cfile=str(dst), dfile=f"<reference>/pydiskann/{rel}", doraise=True)
    # 3. the reference's own benchmark driver (dataset_benchmark.py:75-176: the recall / latency / QPS table), as bytecode only
    bench_py = ref / "dataset_benchmark.py"
    if bench_py.exists():
        py_compile.compile(str(bench_py), cfile=str(OUT / "dataset_benchmark.pycbin"), dfile="<reference>/dataset_benchmark.py", doraise=True)
    stamp.write_text("ok\n")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(Path(a.ref), a.force)
    print("oracle/_ref built" if ok else "reference tree not present; nothing built")
