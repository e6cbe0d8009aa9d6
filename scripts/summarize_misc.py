"""gpurun_out/r02_misc_*.csv (scripts/gpu_profile_misc.sh) -> profiles/r02_misc_kernels.txt: one block per profiled launch."""
import csv, sys
from pathlib import Path
root = Path(__file__).resolve().parent.parent
out = []
for f in sorted((root / "gpurun_out").glob("r02_misc_*.csv")):
    rows = [r for r in csv.reader(f.read_text().splitlines()) if len(r) > 10]
    if not rows:
        continue
    h = rows[0]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    idi = h.index("ID")
    cur = None
    log = (root / "gpurun_out" / (f.stem + ".log")).read_text().strip().splitlines()
    out.append(f"== {f.stem}  ({log[-1] if log else ''})")
    for r in rows[1:]:
        if r[idi] != cur:
            cur = r[idi]
            out.append(f"  launch {cur}: {r[ki][:120]}")
        out.append(f"      {r[mi]:75s} {r[vi]:>18s} {r[ui]}")
(root / "profiles" / "r02_misc_kernels.txt").write_text("\n".join(out) + "\n")
print("\n".join(out[:60]))
