"""Per-source-line stall samples and executed instructions of one profiled kernel, from ncu's own SASS <-> CUDA correlation
(the report must have been captured with --import-source on and the library built with -lineinfo).
usage: python scripts/ncu_lines.py <report.ncu-rep> [top] [queries-in-profiled-launch]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, kernel = None, None, None
agg = collections.defaultdict(lambda: [0, 0])
src = {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        kernel = r[1]
    elif r[0] == "Line No":
        hdr = r
        si, ei = r.index("Warp Stall Sampling (All Samples)"), r.index("Instructions Executed")
    elif hdr is not None and r[0].isdigit() and not r[2].startswith("0x"):      # a source-line row (its SASS rows follow)
        key = (cur, int(r[0]))
        src[key] = r[1].strip()
        agg[key][0] += int(r[si]) if r[si].isdigit() else 0
        agg[key][1] += int(r[ei]) if r[ei].isdigit() else 0
ts = sum(v[0] for v in agg.values()) or 1
te = sum(v[1] for v in agg.values()) or 1
print(f"kernel: {kernel}")
print(f"total stall samples {ts}, warp-level instructions executed {te}" + (f" ({te / nq:.0f} per query)" if nq else ""))
print("  samples  share   instr-share" + ("  instr/query" if nq else "") + "  source line")
for (f, ln), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    per = f" {e / nq:11.0f}" if nq else ""
    print(f"{s:9d} {100 * s / ts:5.1f}% {100 * e / te:12.1f}%{per}  {f}:{ln}  {src[(f, ln)][:110]}")
