"""Compile libdiskrag_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libdiskrag_b200.so"
SOURCES = ["api.cu", "search.cu", "search_fast.cu", "lut_tc.cu", "kmeans_tc.cu", "pq.cu", "distance.cu", "build.cu", "beam_c.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # FMAs only where written explicitly: the summation orders are part of the contract
    "-Xcompiler", "-fPIC", "-shared",
] + (["-DDR_PHASE_TIMING"] if os.environ.get("DR_PHASE_TIMING") else []) + (
    ["-DDR_FAST_NT=" + os.environ["DR_FAST_NT"]] if os.environ.get("DR_FAST_NT") else [])   # debug: per-phase cycle accounting in the search kernel


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / "common.cuh", CSRC / "tc_common.cuh", HERE.parent / "include" / "diskrag_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    (HERE / "build").mkdir(exist_ok=True)
    for s in SOURCES:
        o = HERE / "build" / (s + ".o")
        cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f != "-shared"], "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    subprocess.check_call([_nvcc(), "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
