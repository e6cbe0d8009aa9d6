# First GPU call of the next round: the device checks that were written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash scripts/gpu_pending.sh'
# Writes gpurun_out/pending_*.log; each check runs in its own process under its own timeout.  Once they pass, drop the
# `first device run pending` xfail markers in tests/test_golden_config0.py and tests/test_beam_c_gpu.py.
set -x
mkdir -p gpurun_out
(cd oracle && gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm)
# 1. variant C with the reference's own semantics (csrc/beam_c.cu) == oracle.c:orc_beam_c, under compute-sanitizer first
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/tools/beam_c_check.py > gpurun_out/pending_beam_c_memcheck.log 2>&1; echo "beam_c memcheck rc=$?"
timeout 300 python tests/tools/beam_c_check.py > gpurun_out/pending_beam_c.log 2>&1; echo "beam_c rc=$?"; tail -1 gpurun_out/pending_beam_c.log
# 2. every device check still carrying the pending_device marker (configs[0] fixture against the REAL reference outputs, beam_c, delete-mask parity), with xfail off
timeout 900 python -m pytest tests -q -m "gpu and pending_device" --runxfail 2>&1 | tee gpurun_out/pending_golden_config0.log | tail -5
# 3. build.cu single-point prune after the per-call scratch change (commit 0534d7e) — part of the regular suite
timeout 900 python -m pytest tests/test_build_gpu.py -q -m gpu -x 2>&1 | tail -3
# 4. BASELINE configs[1] parity on the 100k graph rebuilt with the compiled summation order (.cache/config2_adj_100000.npz; if absent:
#    python tests/tools/build_config2_graph.py first, ~12 min of one CPU core, in the build container)
ls .cache/config2_adj_100000.npz && timeout 1500 python tests/tools/parity_config2.py 10000 2000 > gpurun_out/pending_config2_parity.json 2> gpurun_out/pending_config2_parity.log; tail -c 1500 gpurun_out/pending_config2_parity.json
