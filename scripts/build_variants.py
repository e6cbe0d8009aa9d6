"""Build kernel variants of libdiskrag_b200.so for A/B runs on the GPU box (one gpurun call, several libraries):
  python scripts/build_variants.py name1="-DFOO=1 -DBAR=2" name2="-DFOO=3" name3@build.cu="-DDR_BUILD_W=8"
-> diskrag_b200/variants/lib_<name>.so; run with DISKRAG_B200_LIB=<that path> python bench.py ...
Only one source is recompiled per variant (search_fast.cu unless name@source.cu says otherwise; the other objects come from the
regular build)."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diskrag_b200 import build_ext as B

B.build()
out = B.HERE / "variants"
out.mkdir(exist_ok=True)
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    src = "search_fast.cu"
    if "@" in name:
        name, src = name.split("@", 1)
    repl = {}
    for one in src.split(","):                                  # name@a.cu,b.cu: the same flags on several sources
        o = out / f"{one[:-3]}_{name}.o"
        cmd = [B._nvcc(), *[f for f in B.NVCC_FLAGS if f != "-shared"], *flags.split(), "-c", str(B.CSRC / one), "-o", str(o)]
        subprocess.check_call(cmd)
        repl[one] = str(o)
    objs = [repl.get(s, str(B.HERE / "build" / (s + ".o"))) for s in B.SOURCES]
    lib = out / f"lib_{name}.so"
    subprocess.check_call([B._nvcc(), "-shared", "-o", str(lib), *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    print(lib)
