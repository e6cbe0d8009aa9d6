"""Time the GPU Vamana build + PQ train and report recall (run on the GPU box)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, torch
import ctypes as C
from diskrag_b200 import _lib
from diskrag_b200._lib import lib, check
from diskrag_b200.synth import synth_torch
from diskrag_b200.engine import GpuIndex, make_params

N, D, R, L, M = [int(x) for x in sys.argv[1:6]]
NQ = 2000
dev = torch.device("cuda:0")
X = synth_torch(N, D, seed=20242, device=dev)
Q = synth_torch(NQ, D, seed=20242, sample_seed=1, device=dev)
torch.cuda.synchronize()
# ground truth by brute force (torch, validation only)
t = time.time()
gt = torch.empty((NQ, 10), dtype=torch.int64, device=dev)
xn = (X * X).sum(1)
for s in range(0, NQ, 500):
    d = xn[None, :] - 2.0 * (Q[s:s + 500] @ X.T)
    gt[s:s + 500] = d.topk(10, largest=False).indices
gt = gt.cpu().numpy(); torch.cuda.synchronize(); print("gt", time.time() - t)
st = torch.cuda.current_stream().cuda_stream
# medoid: sample 1000
samples = torch.randperm(N, device=dev)[:1000]
cent = X.mean(0)
med = int(((X - cent) ** 2).sum(1).argmin().item())
adj = torch.empty((N, R), dtype=torch.int32, device=dev); deg = torch.empty(N, dtype=torch.int32, device=dev)
torch.cuda.synchronize(); t = time.time()
check(lib().dr_vamana_build_dev(X.data_ptr(), N, D, R, L, 1.2, med, 1234, adj.data_ptr(), deg.data_ptr(), 0, st))
torch.cuda.synchronize(); tb = time.time() - t
print(f"build N={N} D={D} R={R} L={L}: {tb:.2f}s  mean deg {deg.float().mean().item():.1f} min {deg.min().item()}")
cb = torch.empty((M, 256, D // M), dtype=torch.float32, device=dev); codes = torch.empty((N, M), dtype=torch.uint8, device=dev)
mse = C.c_double(0)
torch.cuda.synchronize(); t = time.time()
check(lib().dr_pq_train_dev(X.data_ptr(), N, D, M, 25, 42, cb.data_ptr(), C.byref(mse), 0, st))
torch.cuda.synchronize(); tt = time.time() - t; t = time.time()
check(lib().dr_pq_encode_dev(cb.data_ptr(), X.data_ptr(), N, D, M, codes.data_ptr(), 0, st))
torch.cuda.synchronize(); te = time.time() - t
print(f"pq train {tt:.2f}s mse {mse.value:.3e} encode {te:.2f}s")
idx = GpuIndex.from_device_ptrs(X.data_ptr(), adj.data_ptr(), codes.data_ptr(), cb.data_ptr(), N, D, R, M, med, 0, keepalive=(X, adj, codes, cb))
ids = torch.empty((NQ, 10), dtype=torch.int32, device=dev); dd = torch.empty((NQ, 10), dtype=torch.float32, device=dev)
hops = torch.empty(NQ, dtype=torch.int32, device=dev); vis = torch.empty(NQ, dtype=torch.int32, device=dev)
def rec(ids):
    a = ids.cpu().numpy()
    return float(np.mean([len(set(a[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(NQ)]))
for (dist, W, adc, Ls, rr) in (("exact", 1, "seq", 100, False), ("pq", 1, "seq", 100, True), ("pq", 4, "tree", 100, True), ("pq", 8, "tree", 100, True), ("pq", 4, "tree", 64, True)):
    p = make_params(k=10, L=Ls, W=W, dist=dist, adc_order=adc, rerank=rr)
    for it in range(2):
        torch.cuda.synchronize(); t = time.time()
        idx.search_dev(Q.data_ptr(), NQ, p, ids.data_ptr(), dd.data_ptr(), hops.data_ptr(), vis.data_ptr(), stream=st)
        torch.cuda.synchronize(); ts = time.time() - t
    print(f"search {dist} W={W} {adc} L={Ls} rerank={rr}: {NQ / ts:.0f} QPS  recall@10 {rec(ids):.4f}  hops {hops.float().mean().item():.1f} visited {vis.float().mean().item():.0f}")
