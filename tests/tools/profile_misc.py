"""Workloads for the ncu captures of the kernels other than the throughput search: run under
  ncu --set full -k regex:<kernel> ...  (scripts/gpu_profile_misc.sh); each section prints its algorithmic bytes per launch.
usage: python tests/tools/profile_misc.py build|refsearch|rowdist|lut|kmeans"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from diskrag_b200 import engine
from diskrag_b200._lib import check, lib
from diskrag_b200.synth import synth_torch

what = sys.argv[1]
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
st = torch.cuda.current_stream(dev).cuda_stream
if what == "build":                                             # K5: prune_kernel (both modes) + the build's exact search (search_kernel, W = 4, row map)
    N, D, R, L = 200_000, 768, 64, 100
    X = synth_torch(N, D, seed=20243, device=dev)
    adj = torch.empty((N, R), dtype=torch.int32, device=dev); deg = torch.empty(N, dtype=torch.int32, device=dev)
    check(lib().dr_vamana_build_dev(X.data_ptr(), N, D, R, L, 1.2, 0, 7, adj.data_ptr(), deg.data_ptr(), 0, st))
    torch.cuda.synchronize()
    print(f"build {N}x{D} R={R} L={L}: mean degree {deg.float().mean().item():.1f}, truncated runs {lib().dr_vamana_build_last_truncated()}")
elif what in ("refsearch", "lut", "kmeans"):
    N, D, R, M = 1_000_000, 1536, 32, 192
    X = synth_torch(N, D, seed=20242, device=dev)
    cb = torch.empty((M, 256, D // M), dtype=torch.float32, device=dev); codes = torch.empty((N, M), dtype=torch.uint8, device=dev)
    mse = C.c_double(0)
    check(lib().dr_pq_train_dev(X.data_ptr(), N, D, M, 25, 42, cb.data_ptr(), C.byref(mse), 0, st))          # K3tc + update kernels
    check(lib().dr_pq_encode_dev(cb.data_ptr(), X.data_ptr(), N, D, M, codes.data_ptr(), 0, st))
    if what == "kmeans":
        torch.cuda.synchronize(); print("pq mse", mse.value); sys.exit(0)
    adj = torch.empty((N, R), dtype=torch.int32, device=dev); deg = torch.empty(N, dtype=torch.int32, device=dev)
    check(lib().dr_vamana_build_dev(X.data_ptr(), N, D, R, 64, 1.2, 0, 1234, adj.data_ptr(), deg.data_ptr(), 0, st))
    idx = engine.GpuIndex.from_device_ptrs(X.data_ptr(), adj.data_ptr(), codes.data_ptr(), cb.data_ptr(), N, D, R, M, 0, 0, keepalive=(X, adj, codes, cb))
    B = 20000
    Q = synth_torch(B, D, seed=20242, sample_seed=1000, device=dev)
    ids = torch.empty((B, 10), dtype=torch.int32, device=dev); dd = torch.empty((B, 10), dtype=torch.float32, device=dev)
    hops = torch.empty(B, dtype=torch.int32, device=dev); vis = torch.empty(B, dtype=torch.int32, device=dev); ll = torch.empty(B, dtype=torch.int32, device=dev)
    if what == "refsearch":                                       # the reference-order kernel: f32 table, sequential ADC, W = 1, rerank
        p = engine.make_params(k=10, L=100, W=1, dist="pq", adc_order="seq", rerank=True, lut="f32")
    else:                                                         # the two tcgen05 table kernels + the throughput search
        p = engine.make_params(k=10, L=100, W=8, dist="pq", adc_order="tree", rerank=True, lut="u8tc", prefetch=5, w2=20)
    torch.cuda.profiler.start()
    idx.search_dev(Q.data_ptr(), B, p, ids.data_ptr(), dd.data_ptr(), hops.data_ptr(), vis.data_ptr(), d_list_len=ll.data_ptr(), stream=st)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    h, v, l = hops.float().mean().item(), vis.float().mean().item(), ll.float().mean().item()
    print(f"{what}: {B} queries, mean hops {h:.1f} visited {v:.1f}; algorithmic bytes per query {4 * D + h * 4 * R + v * M + l * 4 * D + 80:.0f}")
elif what == "rowdist":                                          # K4: row-wise L2^2 of 1M gathered rows against one query
    from diskrag_b200 import ops
    n, D = 262144, 1536
    A = np.random.default_rng(0).standard_normal((n, D)).astype(np.float32)
    q = A[:1].copy()
    ops.l2sq_batch(A, q)
    print(f"rowdist: {n} rows x {D}: algorithmic bytes {n * D * 4 + D * 4 + n * 4}")
