"""SASS-level view of one ncu capture (--import-source on, -lineinfo): executed warp-instructions per opcode and per
(source line, opcode), to see where a kernel's instruction budget goes.
usage: python scripts/ncu_sass.py <report.ncu-rep> [queries-in-profiled-launch] [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, line = None, None, None
byop = collections.Counter(); byline = collections.Counter(); bylineop = collections.Counter(); stall = collections.Counter()
src = {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        si, ei = r.index("Warp Stall Sampling (All Samples)"), r.index("Instructions Executed")
    elif hdr is not None and r[0].isdigit() and not r[2].startswith("0x"):
        line = (cur, int(r[0])); src[line] = r[1].strip()
    elif hdr is not None and len(r) > 2 and r[2].startswith("0x"):
        sass = r[3] if len(r) > 3 else ""
        # columns of a SASS row: '', '', address, sass text, then the metric columns shifted like the header
        toks = sass.replace("@!", "@").split()
        op = next((t for t in toks if not t.startswith("@")), "?").split(".")[0]
        e = int(r[ei]) if r[ei].isdigit() else 0
        s = int(r[si]) if r[si].isdigit() else 0
        byop[op] += e; byline[line] += e; bylineop[(line, op)] += e; stall[line] += s
te = sum(byop.values()) or 1
print(f"executed warp-instructions: {te} ({te / nq:.0f} per query)")
print("by opcode:")
for op, e in byop.most_common(30):
    print(f"  {op:12s} {e / nq:9.0f}  {100 * e / te:5.1f}%")
print("by source line (instr/query, share, stall-sample share):")
ts = sum(stall.values()) or 1
for ln, e in byline.most_common(top):
    ops = ", ".join(f"{o}:{c / nq:.0f}" for (l, o), c in sorted(bylineop.items(), key=lambda kv: -kv[1]) if l == ln)[:150]
    print(f"  {e / nq:8.0f} {100 * e / te:5.1f}% st {100 * stall[ln] / ts:4.1f}%  {ln[0]}:{ln[1]}  {src.get(ln, '')[:70]} | {ops}")
