"""Phase-level view of one ncu capture of search_fast_kernel: stall samples, executed instructions and the top stall reasons per
phase of a query (table staging, P1 expand + claim, P2 ADC, merge, rerank ...), from ncu's own SASS-to-CUDA correlation and the
source imported into the report (--import-source on).  Phase boundaries are the marker comments / statements of search_fast.cu.
usage: python scripts/ncu_phases.py <report.ncu-rep>"""
import collections, csv, subprocess, sys
rep = sys.argv[1]


def page(view):
    return subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout


# 1. the profiled source: where the phases start
text, cur = {}, None
for r in csv.reader(page("cuda").splitlines()):
    if not r:
        continue
    if r[0] == "File Name":
        cur = r[1].split("/")[-1]
    elif cur == "search_fast.cu" and r[0].isdigit():
        text[int(r[0])] = r[1]
MARKS = [("prologue / next-query fetch (block barrier)", "__global__ void __launch_bounds__(DR_FAST_NT"), ("table staging", "---- stage the query's table"),
         ("hash clear", "s_hash[i] = DR_EMPTY"), ("start node", "if (wid == 0) {"), ("P1 expand + claim", "(1+2) The selection"),
         ("P2 ADC + survivor append", "(3) quantised ADC"), ("merge", "(4) merge"), ("list output", "const u64 *lst = cur ? s_list1 : s_list0;"),
         ("rerank (staging, waits)", "if (do_rerank) {"), ("rerank rank + output", "DR_PT(5)"), ("host code", "typedef void (*fast_kernel_t)")]
HELPERS = [("P2 helper: code-row loads + table lookups", "adc_u8_warp("), ("P1 helper: visited insert", "fib_slot(uint32_t id"),
           ("rerank helper: L2 distance piece", "l2sq_piece_smem("), ("merge helper: binary search", "lower_bound_u64(")]
starts = []
for name, pat in MARKS + HELPERS:
    ln = next((l for l in sorted(text) if pat in text[l]), None)
    if ln is not None:
        starts.append((ln, name))
starts.sort()


def phase(f, ln):
    if f != "search_fast.cu":
        return "inlined headers (atomics, shuffles, redux, mbarrier)"
    p = "file prologue"
    for l, name in starts:
        if ln >= l:
            p = name
    return p


# 2. per-line samples -> phases
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
hdr = None
for r in csv.reader(page("cuda,sass").splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        kernel = r[1]
    elif r[0] == "Line No":
        hdr = r
        si, ei = r.index("Warp Stall Sampling (All Samples)"), r.index("Instructions Executed")
        cols = [(i, c) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]
    elif hdr is not None and r[0].isdigit() and not r[2].startswith("0x"):
        a = agg[phase(cur, int(r[0]))]
        a[0] += int(r[si]) if r[si].isdigit() else 0
        a[1] += int(r[ei]) if r[ei].isdigit() else 0
        for i, c in cols:
            if r[i].isdigit():
                a[2][c] += int(r[i])
ts = sum(v[0] for v in agg.values()) or 1
te = sum(v[1] for v in agg.values()) or 1
print(f"kernel: {kernel}")
print(f"{'phase':55s} stall%  instr%  top stall reasons (share of the phase's samples)")
for p, (s, e, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = ", ".join(f"{k[6:]} {100 * v / max(1, sum(c.values())):.0f}%" for k, v in c.most_common(3))
    print(f"{p:55s} {100 * s / ts:5.1f}  {100 * e / te:5.1f}   {top}")
