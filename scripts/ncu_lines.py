"""Join ncu's per-SASS-instruction stall samples with source lines (nvdisasm -g line info).
usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel-substring> <cu-file-substring> [top]"""
import csv, re, subprocess, sys, tempfile, os, glob
rep, kname, cufile = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si = h.index("Warp Stall Sampling (All Samples)"); ei = h.index("Instructions Executed")
inst = [(r[1].strip(), int(r[si] or 0), int(r[ei] or 0)) for r in rows[hi + 1:] if len(r) > ei and r[0].startswith("0x")]
# line table from the cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("diskrag_b200/libdiskrag_b200.so")], cwd=tmp, capture_output=True)
sass = None
for f in glob.glob(tmp + "/*.cubin"):
    t = subprocess.run(["nvdisasm", "-c", "-g", f], capture_output=True, text=True).stdout
    if kname in t:
        sass = t
        break
lines = []
cur = None
infn = False
for l in sass.splitlines():
    if l.startswith("\t.section") or l.startswith(".section"):
        infn = kname in l and ".text." in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
n = min(len(lines), len(inst))
print(f"{len(inst)} instructions in report, {len(lines)} in cubin")
agg = {}
tot = sum(s for _, s, _ in inst)
for (txt, s, e), loc in zip(inst[:n], lines[:n]):
    key = loc if loc and cufile in loc[0] else (loc[0].split("/")[-1] if loc else "?", loc[1] if loc else 0)
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += e
src = {}
for key in agg:
    f = key[0]
    if os.path.exists(f) and f not in src:
        src[f] = open(f).read().splitlines()
print(f"total samples {tot}")
for key, (s, e) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    f, ln = key
    text = src[f][ln - 1].strip()[:100] if f in src and 0 < ln <= len(src[f]) else ""
    print(f"{s:8d} {100.0 * s / max(tot, 1):5.1f}%  exec {e:12d}  {os.path.basename(f)}:{ln}  {text}")
