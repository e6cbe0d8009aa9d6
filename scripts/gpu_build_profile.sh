# launch list (durations only) of one graph build: which kernels the build's seconds go to.  usage: gpu_build_profile.sh <tag> "N D R L"
TAG=$1; shape="$2"
mkdir -p gpurun_out
BUILD_AB_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_build_launches.csv \
  python tests/tools/build_ab.py $shape > gpurun_out/${TAG}_build_profile.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/${TAG}_build_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum": continue
    name = r[ki].split("(")[0]; tot[name] += float(r[vi].replace(",", "")); cnt[name] += 1
s = sum(tot.values())
print("BUILD_PROFILE total %.1f ms over %d launches (serialised, cold-cache per launch)" % (s / 1e6, sum(cnt.values())))
for k, v in tot.most_common(12): print("  %-60s %6d launches %9.1f ms %5.1f %%" % (k[:60], cnt[k], v / 1e6, 100 * v / s))
PY
