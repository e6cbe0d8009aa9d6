"""Turn the ncu captures of scripts/profile.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
usage: python scripts/summarize_profile.py <tag> <kernel-name-substring> <cu-file> <queries-in-profiled-launch>"""
import csv, json, shutil, subprocess, sys
from pathlib import Path
tag, kname, cufile, nq = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
root = Path(__file__).resolve().parent.parent
rep = root / "gpurun_out" / f"{tag}_search.ncu-rep"
out = root / "profiles"
det = subprocess.run(["ncu", "-i", str(rep), "--page", "details"], capture_output=True, text=True).stdout
(out / f"{tag}_search_kernel_ncu.txt").write_text(det)
lines = subprocess.run([sys.executable, str(root / "scripts" / "ncu_lines.py"), str(rep), "40", str(nq)], capture_output=True, text=True, cwd=root).stdout
(out / f"{tag}_search_kernel_lines.txt").write_text(lines)
shutil.copy(root / "gpurun_out" / f"{tag}_launches.csv", out / f"{tag}_step_launches.csv")
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
d = {k: (val, unit) for k, unit, val in zip(h, u, v)}
def to_bytes(key):
    val, unit = d[key]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(val) * mult
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
info = {"profile": f"profiles/{tag}_search_kernel_ncu.txt", "kernel": d["Kernel Name"][0] if "Kernel Name" in d else kname,
        "queries_in_profiled_launch": nq, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_query": (rd + wr) / nq, "duration_ms": float(d["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[d["gpu__time_duration.sum"][1]],
        "source": f"ncu --set full capture profiles/{tag}_search_kernel_ncu.txt (commit " + subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=root).stdout.strip() + "), scaled per query",
        "note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture; bench.py scales it to its own launch size"}
(out / "search_kernel_traffic.json").write_text(json.dumps(info, indent=1))
print(json.dumps(info, indent=1))
