#!/usr/bin/env python3
"""bench.py — the hot-path benchmark (contract in the task statement; metric from BASELINE.json).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 1M x 1536 synthetic unit-norm
embeddings, Vamana R=32 (GPU-built, L_build=64, alpha=1.2), PQ M=192, search list L=100, top-10 with exact
rerank, 100k-query batch per GPU, index replicated, queries sharded ("weak": per-GPU work fixed).
One step = one pass of the hot path over one 100k-query batch: ADC-table build + beam search + rerank.

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                      the reference's own CPU code (oracle/_ref) on host cores
  python bench.py --scaling strong [...]                      configs[2] literally: ONE 100k-query batch split over the N GPUs
  python bench.py --mode index-sharded [--exchange p2p]       configs[4] shape: every GPU owns a 1.25M x 3072 shard (own graph, PQ),
                                                              searches all queries on it, partial top-k exchanged and merged
Under torchrun each rank drives one GPU; rank 0 prints one JSON line (same schema in every mode).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "QPS@recall10>=0.95"
UNIT = "queries/s"
# the arithmetic the path computes in, per table mode (the final rerank and every reported distance are fp32 in all of them)
DTYPES = {"f32": "f32 (fp32 ADC table in the reference's operation order, fp32 sums, fp32 rerank)",
          "u8": "f32 rerank; u8 ADC table (exact fp32 build), integer sums",
          "u8tc": "f32 rerank; u8 ADC table (tf32 tensor-core build), integer sums"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="query-sharded", choices=["query-sharded", "index-sharded"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="query-sharded: weak = --queries per GPU; strong = --queries in total, split over the GPUs")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"],
                    help="index-sharded: one packed NCCL all-to-all, or the search kernel's epilogue storing into peer memory")
    ap.add_argument("--n", type=int, default=None, help="corpus rows per GPU (default 1M; index-sharded: 1.25M per shard)")
    ap.add_argument("--dim", type=int, default=None, help="default 1536; index-sharded: 3072")
    ap.add_argument("--R", type=int, default=32)
    ap.add_argument("--Lbuild", type=int, default=64)
    ap.add_argument("--M", type=int, default=192)
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--W", type=int, default=8)
    ap.add_argument("--W2", type=int, default=20, help="expansions of a step that follows a step without survivors (0 = off; results restated by the oracle)")
    ap.add_argument("--adc", default="tree", choices=["seq", "tree"])
    ap.add_argument("--lut", default="u8tc", choices=["f32", "u8", "u8tc"], help="ADC table: f32 reference arithmetic, u8 exact 8-bit, u8tc 8-bit built on tensor cores")
    ap.add_argument("--prefetch", type=int, default=5, help="L2 prefetch bit mask of the throughput kernel (include/diskrag_b200.h); results are unchanged")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=100_000, help="queries per GPU per step")
    ap.add_argument("--gt-queries", type=int, default=1000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--hash-cap", type=int, default=0, help="force the visited-table size (occupancy experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-points", action="store_true", help="skip the other operating points / parity-mode legs (kernel A/B runs)")
    ap.add_argument("--cuda-profile", action="store_true", help="wrap one extra step in cudaProfilerStart/Stop (ncu --profile-from-start off)")
    a = ap.parse_args()
    if a.n is None:
        a.n = 1_250_000 if a.mode == "index-sharded" else 1_000_000
    if a.dim is None:
        a.dim = 3072 if a.mode == "index-sharded" else 1536
    return a


def _gpu_numa_node(local):
    """NUMA node of GPU `local` from sysfs, via its PCI bus id (NVML, else nvidia-smi); None when the box does not say."""
    bus = None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
    except Exception:
        try:
            bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
        except Exception:
            bus = None
    if not bus:
        return None
    bus = bus.lower()
    if len(bus.split(":")[0]) == 8:                      # NVML prints an 8-digit domain, sysfs a 4-digit one
        bus = bus[4:]
    p = Path(f"/sys/bus/pci/devices/{bus}/numa_node")
    try:
        node = int(p.read_text())
        return node if node >= 0 else None
    except Exception:
        return None


def pin_rank_to_local_cpus(local, world):
    """Each rank on its own cores, on the NUMA node of its GPU when sysfs says which: the pinned host buffers of the end-to-end leg
    are then allocated node-local (first touch) and 8 ranks do not pull 8 x 614 MB per step through one socket's memory controllers
    and the inter-socket link.  Best effort, silent."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        node = _gpu_numa_node(local)
        node_cpus = None
        if node is not None:
            lst = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
            node_cpus = []
            for part in lst.split(","):
                lo, _, hi = part.partition("-")
                node_cpus += list(range(int(lo), int(hi or lo) + 1))
            node_cpus = [c for c in node_cpus if c in cpus] or None
        if node_cpus:
            # the ranks that share this node split its cores among themselves (GPUs are numbered node by node on HGX boards)
            peers = [r for r in range(world) if _gpu_numa_node(r) == node] or [local]
            per = max(1, len(node_cpus) // len(peers))
            k = peers.index(local) if local in peers else 0
            mine = node_cpus[k * per:(k + 1) * per] or node_cpus
        else:
            per = max(1, len(cpus) // max(1, min(world, len(cpus))))
            mine = cpus[(local * per) % len(cpus):(local * per) % len(cpus) + per] or cpus
        os.sched_setaffinity(0, mine)
        return {"cpus": len(mine), "numa_node": node, "numa_local": bool(node_cpus)}
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
# setup shared by both arms: synthetic corpus, PQ, graph (all on the GPU, untimed)
# ------------------------------------------------------------------------------------------------------
def build_index(a, dev, data_seed=20242, sample_seed=None, build_seed=1234):
    import torch
    from diskrag_b200._lib import check, lib
    from diskrag_b200.synth import synth_torch
    st = torch.cuda.current_stream(dev).cuda_stream
    N, D, R, M = a.n, a.dim, a.R, a.M
    t0 = time.time()
    X = synth_torch(N, D, seed=data_seed, device=dev) if sample_seed is None else synth_torch(N, D, seed=data_seed, sample_seed=sample_seed, device=dev)
    # medoid: sampled, as compute_approximate_medoid_cython does (1000 samples), through our kernel
    g = torch.Generator(device=dev); g.manual_seed(77)
    smp = torch.randperm(N, generator=g, device=dev)[:1000].to(torch.int32).cpu().numpy()
    med = medoid_dev(X, smp, dev)
    cb = torch.empty((M, 256, D // M), dtype=torch.float32, device=dev)
    codes = torch.empty((N, M), dtype=torch.uint8, device=dev)
    mse = C.c_double(0)
    t1 = time.time()
    check(lib().dr_pq_train_dev(X.data_ptr(), N, D, M, 25, 42, cb.data_ptr(), C.byref(mse), dev.index, st), "dr_pq_train_dev")
    check(lib().dr_pq_encode_dev(cb.data_ptr(), X.data_ptr(), N, D, M, codes.data_ptr(), dev.index, st), "dr_pq_encode_dev")
    torch.cuda.synchronize(dev)
    t2 = time.time()
    adj = torch.empty((N, R), dtype=torch.int32, device=dev)
    deg = torch.empty(N, dtype=torch.int32, device=dev)
    check(lib().dr_vamana_build_dev(X.data_ptr(), N, D, R, a.Lbuild, 1.2, med, build_seed, adj.data_ptr(), deg.data_ptr(), dev.index, st),
          "dr_vamana_build_dev")
    torch.cuda.synchronize(dev)
    t3 = time.time()
    info = {"gen_s": round(t1 - t0, 2), "pq_train_encode_s": round(t2 - t1, 2), "graph_build_s": round(t3 - t2, 2),
            "pq_mse": mse.value, "mean_degree": round(float(deg.float().mean().item()), 2)}
    return X, adj, deg, codes, cb, med, info


def medoid_dev(X, samples, dev):
    """argmin over the samples of sum_j ||x_s - x_j|| (compute_approximate_medoid_cython, cython_utils.pyx:210-263) on the
    device-resident corpus, through the library (dr_medoid_dev): no torch arithmetic in the setup either."""
    import torch
    from diskrag_b200._lib import check, lib
    smp = torch.from_numpy(np.ascontiguousarray(samples, np.int32)).to(dev)
    out = C.c_int64(0)
    check(lib().dr_medoid_dev(X.data_ptr(), X.shape[0], X.shape[1], smp.data_ptr(), int(smp.numel()), C.byref(out), dev.index,
                              torch.cuda.current_stream(dev).cuda_stream), "dr_medoid_dev")
    return int(out.value)


def ground_truth(X, Q, k):
    import torch
    xn = (X * X).sum(1)
    out = torch.empty((Q.shape[0], k), dtype=torch.int64, device=X.device)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    for s in range(0, Q.shape[0], 250):
        d = xn[None, :] - 2.0 * (Q[s:s + 250] @ X.T)
        out[s:s + 250] = d.topk(k, largest=False).indices
    torch.backends.cuda.matmul.allow_tf32 = prev
    return out.cpu().numpy()


def recall(ids, gt, k):
    return float(np.mean([len(set(ids[i, :k].tolist()) & set(gt[i, :k].tolist())) / k for i in range(gt.shape[0])]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(a, hops, visited, list_len):
    """bytes(q) = 4D + h*4R + v*M + L_r*4D + 8k, summed over the batch (int64)."""
    h = hops.astype(np.int64); v = visited.astype(np.int64); lr = list_len.astype(np.int64)
    return int((4 * a.dim + h * 4 * a.R + v * a.M + lr * 4 * a.dim + 8 * a.k).sum())


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from diskrag_b200 import engine
    from diskrag_b200.synth import synth_torch
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    affinity = pin_rank_to_local_cpus(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.lut == "u8tc" and ((a.dim // a.M) % 8 != 0 or a.M % 4 != 0):
        a.lut = "u8"                                   # the tcgen05 table needs a sub-dimension that is a multiple of 8
    X, adj, deg, codes, cb, med, info = build_index(a, dev)
    idx = engine.GpuIndex.from_device_ptrs(X.data_ptr(), adj.data_ptr(), codes.data_ptr(), cb.data_ptr(), a.n, a.dim, a.R, a.M,
                                           med, local, keepalive=(X, adj, codes, cb))
    k = a.k
    if a.scaling == "strong":
        # BASELINE configs[2] literally: ONE --queries batch, query-sharded over the ranks (contiguous slices); the index is
        # replicated, there is no exchange on the data path, the slices' results are concatenated by the caller
        from diskrag_b200.dist import query_slice
        qlo, qhi = query_slice(a.queries, rank, world)
        B = qhi - qlo
        Q = synth_torch(a.queries, a.dim, seed=20242, sample_seed=1000, device=dev)[qlo:qhi].contiguous()
        total_queries = a.queries
    else:
        B = a.queries
        Q = synth_torch(B, a.dim, seed=20242, sample_seed=1000 + rank, device=dev)     # each rank: its own query shard
        total_queries = world * B
    p = engine.make_params(k=k, L=a.L, W=a.W, dist="pq", adc_order=a.adc, rerank=True, threads=a.threads, lut=a.lut, hash_cap=a.hash_cap,
                           prefetch=int(a.prefetch), w2=a.W2 if a.lut != "f32" else 0)
    ids = torch.empty((B, k), dtype=torch.int32, device=dev); dd = torch.empty((B, k), dtype=torch.float32, device=dev)
    hops = torch.empty(B, dtype=torch.int32, device=dev); vis = torch.empty(B, dtype=torch.int32, device=dev)
    llen = torch.empty(B, dtype=torch.int32, device=dev); stat = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def step():
        idx.search_dev(Q.data_ptr(), B, p, ids.data_ptr(), dd.data_ptr(), hops.data_ptr(), vis.data_ptr(),
                       d_list_len=llen.data_ptr(), d_status=stat.data_ptr(), stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # recall on a ground-truth subset (rank 0's shard)
    step(); torch.cuda.synchronize(dev)
    assert int(stat.abs().sum().item()) == 0, "search reported a non-zero status"
    ngt = min(a.gt_queries, B)
    gt = ground_truth(X, Q[:ngt], k)
    rec = recall(ids[:ngt].cpu().numpy(), gt, k)

    # ---- value: device-resident inputs, CUDA events on the launching stream ---------------------------
    for _ in range(a.warmup):
        step()
    if a.cuda_profile:
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        step()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    l0 = engine.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = engine.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = total_queries * a.steps / (ms_total / 1e3)

    # ---- roofline of the search kernel: algorithmic bytes / its own launch durations (CUDA events in the library)
    h_np, v_np, l_np = hops.cpu().numpy(), vis.cpu().numpy(), llen.cpu().numpy()
    abytes = algorithmic_bytes(a, h_np, v_np, l_np)
    idx.kernel_timing(True)
    for _ in range(2):
        step()
    torch.cuda.synchronize(dev)
    k_ms, k_launches = idx.kernel_timing(False)
    peak, peak_src = measured_peak()
    per_launch_bytes = abytes * 2 / max(1, k_launches)
    per_launch_ms = k_ms / max(1, k_launches)
    achieved = per_launch_bytes / (per_launch_ms / 1e3) / 1e9
    traffic = traffic_src = None
    tf = ROOT / "profiles" / "search_kernel_traffic.json"
    if tf.exists() and a.lut != "f32":
        try:
            # NOT measured in this run: one ncu --set full capture of this kernel (scripts/profile.sh + scripts/summarize_profile.py),
            # dram__bytes_read.sum + dram__bytes_write.sum per query, scaled to this launch size; the file says which capture
            tj = json.loads(tf.read_text())
            traffic = round(tj["dram_bytes_per_query"] * (B * 2 / max(1, k_launches)))
            traffic_src = tj.get("source", "profiles/search_kernel_traffic.json")
        except Exception:
            traffic = traffic_src = None
    tab_b = a.M * (256 if a.lut != "f32" else 1024)       # the ADC table of one query: written by the table kernels, read back by the search
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "search_fast_kernel" if a.lut != "f32" else "search_kernel",
                "kernel_ms_per_launch": round(per_launch_ms, 3), "launches_per_step": k_launches // 2,
                "algorithmic_bytes_per_query": round(abytes / B, 1), "kernel_share_of_step": round((k_ms / 2) / (ms_total / a.steps), 3),
                # NOT in the algorithmic bytes (SURVEY §8d excludes table traffic): the table's round trip through HBM
                "adc_table_bytes_per_query": {"written_by_table_kernels": tab_b, "read_by_search_kernel": tab_b},
                "frac_counting_table_read": round((abytes / B + tab_b) * (B * 2 / max(1, k_launches)) / (per_launch_ms / 1e3) / 1e9 / peak, 4),
                "step_level_frac": round(abytes / (ms_total / a.steps / 1e3) / 1e9 / peak, 4)}

    # ---- e2e: host buffers through the public API (H2D of the queries and D2H of the results inside) ------
    Qh = torch.empty((B, a.dim), dtype=torch.float32, pin_memory=True); Qh.copy_(Q)
    ids_h = torch.empty((B, k), dtype=torch.int32, pin_memory=True); dd_h = torch.empty((B, k), dtype=torch.float32, pin_memory=True)
    hops_h = torch.empty(B, dtype=torch.int32, pin_memory=True); vis_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
    st_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
    Qn, idn, ddn, hn, vn, sn = Qh.numpy(), ids_h.numpy(), dd_h.numpy(), hops_h.numpy(), vis_h.numpy(), st_h.numpy()

    def step_e2e():
        idx.search_host(Qn, p, idn, ddn, hn, vn, sn)

    for _ in range(max(1, a.warmup - 1)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = total_queries * a.steps / float(te.item())
    assert np.array_equal(idn, ids.cpu().numpy()), "host-API results differ from the device-API results"
    e2e = {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": int(B * a.dim * 4),
           "d2h_bytes_per_step": int(B * k * 8 + 3 * B * 4)}

    # ---- the recall / throughput trade-off around the named configuration (context for "QPS at recall >= 0.95"; 1 GPU only)
    points = parity = None
    if world == 1 and not a.no_points:
        # ---- the reference-identical mode on the same index: variant A (f32 table in the reference's operation order, sequential
        # ADC, W = 1: bit-exact with cython_utils.pyx:72-122 by tests/) + the exact rerank of search_engine.py:374-379 — its own
        # throughput and roofline fraction, and how many of the bench mode's top-k id SETS equal its answer
        npar = min(B, 20000)
        bench_ids = ids[:npar].cpu().numpy().copy()
        pr = engine.make_params(k=k, L=a.L, W=1, dist="pq", adc_order="seq", rerank=True, lut="f32")
        runp = lambda: idx.search_dev(Q.data_ptr(), npar, pr, ids.data_ptr(), dd.data_ptr(), hops.data_ptr(), vis.data_ptr(),
                                      d_list_len=llen.data_ptr(), d_status=stat.data_ptr(), stream=stream)
        runp(); torch.cuda.synchronize(dev)
        ref_ids = ids[:npar].cpu().numpy().copy()
        pbytes = algorithmic_bytes(a, hops[:npar].cpu().numpy(), vis[:npar].cpu().numpy(), llen[:npar].cpu().numpy())
        idx.kernel_timing(True)
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(); runp(); runp(); f1.record(); torch.cuda.synchronize(dev)
        pk_ms, pk_n = idx.kernel_timing(False)
        same = np.array([set(bench_ids[i].tolist()) == set(ref_ids[i].tolist()) for i in range(npar)])
        parity = {"reference_mode": "variant A: f32 table (reference operation order), sequential ADC, W=1, L=%d + exact rerank" % a.L,
                  "queries": npar, "qps": round(2 * npar / (f0.elapsed_time(f1) / 1e3), 1),
                  "recall_at_10": round(recall(ref_ids[:ngt], gt, k), 4),
                  "kernel": "search_kernel", "kernel_ms_per_launch": round(pk_ms / max(1, pk_n), 3),
                  "roofline_frac": round(pbytes * 2 / max(1, pk_n) / (pk_ms / max(1, pk_n) / 1e3) / 1e9 / peak, 4),
                  "bench_mode_top%d_sets_equal_to_reference_mode" % k: f"{int(same.sum())}/{npar}",
                  "fraction": round(float(same.mean()), 4)}
        points = []
        for Lp in (50, 64, 80):
            pp = engine.make_params(k=k, L=Lp, W=a.W, dist="pq", adc_order=a.adc, rerank=True, threads=a.threads, lut=a.lut,
                                    hash_cap=a.hash_cap, prefetch=int(a.prefetch), w2=a.W2 if a.lut != "f32" else 0)
            run = lambda: idx.search_dev(Q.data_ptr(), B, pp, ids.data_ptr(), dd.data_ptr(), hops.data_ptr(), vis.data_ptr(),
                                         d_list_len=llen.data_ptr(), d_status=stat.data_ptr(), stream=stream)
            run(); torch.cuda.synchronize(dev)
            f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
            f0.record(); run(); run(); f1.record(); torch.cuda.synchronize(dev)
            points.append({"L": Lp, "recall_at_10": round(recall(ids[:ngt].cpu().numpy(), gt, k), 4),
                           "qps": round(2 * B / (f0.elapsed_time(f1) / 1e3), 1)})
        step(); torch.cuda.synchronize(dev)            # leave the named configuration's results in the buffers

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base = cpu_baseline_port(a, X, adj, codes, cb, med, Q, idn)
    if rank == 0:
        out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": round(ms_total / a.steps, 3), "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
               "dtype": DTYPES[a.lut], "data": "synthetic",
               "config": {"workload": f"{a.n}x{a.dim} synthetic unit-norm, Vamana R={a.R} (GPU-built, Lbuild={a.Lbuild}, alpha=1.2), "
                                      f"PQ M={a.M}, L={a.L}, W={a.W}" + (f" ({a.W2} after a step without survivors)" if a.W2 > a.W and a.lut != "f32" else "")
                                      + f", adc={a.adc}, table={a.lut}, rerank, top-{a.k}"
                                      + (f"; ONE {a.queries}-query batch split over the GPUs" if a.scaling == "strong" else ""),
                          "queries_per_gpu_per_step": B, "queries_per_step_total": total_queries, "index": "replicated",
                          "queries": "sharded", "recall_at_10": round(rec, 4), "parity_mode": parity, "cpu_affinity": affinity,
                          "recall_queries": ngt, "l2_flush": f"inputs larger than L2 (index {(a.n * (a.dim * 4 + a.R * 4 + a.M)) / 1e9:.1f} GB, per-step ADC tables {B * a.M * (256 if a.lut != 'f32' else 1024) / 1e9:.1f} GB, queries {B * a.dim * 4 / 1e9:.2f} GB)",
                          "mean_hops": round(float(h_np.mean()), 1), "mean_visited": round(float(v_np.mean()), 1), "setup": info,
                          "other_operating_points": points},
               "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_base}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_index_sharded(a):
    """BASELINE configs[4] shape under the same contract: rank r owns shard r (a.n x a.dim rows, its OWN Vamana graph, medoid and PQ),
    every rank searches ALL --queries queries on its shard (throughput kernel, fused exact rerank), the per-shard top-k lists are
    exchanged as packed 64-bit keys and k-way merged on the rank that owns the query slice.  Exchange: one NCCL all-to-all of a
    device-packed buffer (--exchange nccl) or none at all — the search kernel's epilogue stores into the owner's buffer over
    NVLink peer memory (--exchange p2p).  value = queries answered over the WHOLE corpus per second; "weak": the corpus grows with N."""
    import torch
    import torch.distributed as dist
    from diskrag_b200 import dist as DD, engine
    from diskrag_b200.synth import synth_torch
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local); torch.cuda.set_device(dev)
    affinity = pin_rank_to_local_cpus(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.lut == "u8tc" and ((a.dim // a.M) % 8 != 0 or a.M % 4 != 0):
        a.lut = "u8"
    a_shard = argparse.Namespace(**vars(a))
    X, adj, deg, codes, cb, med, info = build_index(a_shard, dev, data_seed=20245, sample_seed=rank, build_seed=1234 + rank)
    idx = engine.GpuIndex.from_device_ptrs(X.data_ptr(), adj.data_ptr(), codes.data_ptr(), cb.data_ptr(), a.n, a.dim, a.R, a.M, med, local,
                                           keepalive=(X, adj, codes, cb))
    B, k = a.queries, a.k
    Q = synth_torch(B, a.dim, seed=20245, sample_seed=100_000, device=dev)          # the SAME batch on every rank
    p = engine.make_params(k=k, L=a.L, W=a.W, dist="pq", adc_order=a.adc, rerank=True, lut=a.lut, prefetch=int(a.prefetch),
                           w2=a.W2 if a.lut != "f32" else 0)
    ids = torch.empty((B, k), dtype=torch.int32, device=dev); dd = torch.empty((B, k), dtype=torch.float32, device=dev)
    hops = torch.empty(B, dtype=torch.int32, device=dev); vis = torch.empty(B, dtype=torch.int32, device=dev)
    llen = torch.empty(B, dtype=torch.int32, device=dev); stat = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    offset = rank * a.n
    qlo, qhi = DD.query_slice(B, rank, world)
    peer = DD.PeerExchange(idx, B, k, offset, local) if (world > 1 and a.exchange == "p2p") else None

    def search(qptr=None):
        idx.search_dev(qptr or Q.data_ptr(), B, p, ids.data_ptr(), dd.data_ptr(), hops.data_ptr(), vis.data_ptr(),
                       d_list_len=llen.data_ptr(), d_status=stat.data_ptr(), stream=stream)

    def step(qptr=None):
        search(qptr)
        if world == 1:
            return ids, dd
        if peer is not None:
            return peer.merge()
        return DD.index_sharded_topk(ids, dd, offset, gather=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    mi, md = step(); torch.cuda.synchronize(dev)
    assert int(stat.abs().sum().item()) == 0, "search reported a non-zero status"
    # global exact ground truth for the first queries: per-shard exact top-k, the same packed exchange (NCCL path)
    ng = min(a.gt_queries, B)
    xn = (X * X).sum(1)
    prev = torch.backends.cuda.matmul.allow_tf32; torch.backends.cuda.matmul.allow_tf32 = False
    ei = torch.empty((ng, k), dtype=torch.int32, device=dev); ed = torch.empty((ng, k), dtype=torch.float32, device=dev)
    for s0 in range(0, ng, 100):
        d = xn[None, :] - 2.0 * (Q[s0:s0 + 100] @ X.T) + 1.0
        t = d.topk(k, largest=False)
        ei[s0:s0 + 100] = t.indices.to(torch.int32); ed[s0:s0 + 100] = t.values
    torch.backends.cuda.matmul.allow_tf32 = prev
    if world > 1:
        ti, _ = DD.index_sharded_topk(ei, ed, offset, gather=True)
        ai, _ = DD.index_sharded_topk(ids[:ng].contiguous(), dd[:ng].contiguous(), offset, gather=True)
        ti, ai = ti.cpu().numpy(), ai.cpu().numpy()
        if peer is not None:                       # the peer-routed exchange must give exactly what the NCCL exchange gives
            ni, nd = DD.index_sharded_topk(ids, dd, offset, gather=False)
            assert torch.equal(ni, mi) and torch.equal(nd, md), "p2p exchange differs from the NCCL exchange"
    else:
        ti, ai = ei.cpu().numpy(), ids[:ng].cpu().numpy()
    rec = recall(ai, ti, k)
    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    l0 = engine.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = engine.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = B * a.steps / (ms_total / 1e3)
    h_np, v_np, l_np = hops.cpu().numpy(), vis.cpu().numpy(), llen.cpu().numpy()
    abytes = algorithmic_bytes(a, h_np, v_np, l_np)
    idx.kernel_timing(True)
    for _ in range(2):
        search()
    torch.cuda.synchronize(dev)
    k_ms, k_launches = idx.kernel_timing(False)
    peak, peak_src = measured_peak()
    per_launch_ms = k_ms / max(1, k_launches)
    achieved = abytes * 2 / max(1, k_launches) / (per_launch_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                "peak_source": peak_src, "kernel": "search_fast_kernel (this rank's shard)", "kernel_ms_per_launch": round(per_launch_ms, 3),
                "algorithmic_bytes_per_query": round(abytes / B, 1), "kernel_share_of_step": round((k_ms / 2) / (ms_total / a.steps), 3)}
    # e2e: the batch arrives from pinned host memory every step, this rank's merged slice goes back to the host
    Qh = torch.empty((B, a.dim), dtype=torch.float32, pin_memory=True); Qh.copy_(Q)
    Qd = torch.empty_like(Q)
    oi_h = torch.empty((qhi - qlo, k), dtype=torch.int32, pin_memory=True); od_h = torch.empty((qhi - qlo, k), dtype=torch.float32, pin_memory=True)

    # Every rank needs the WHOLE batch on its device.  Uploading it G times (once per rank: G x B x D x 4 bytes over PCIe) is what bounds the
    # end-to-end number; instead every rank uploads only its slice (the one it also reduces) and the slices are all-gathered over NVLink:
    # the one place on this path where a collective carries real data (B x D x 4 bytes per step, 1.2 GB at the named shape).
    bq = DD.padded_slice_len(B, world)
    Qpad = torch.empty((world * bq, a.dim), dtype=torch.float32, device=dev) if world > 1 else None
    Qslice_h = Qh[qlo:qhi]

    def step_e2e():
        if world == 1:
            Qd.copy_(Qh, non_blocking=True)
            qptr = Qd.data_ptr()
        else:
            mine = Qpad[rank * bq:rank * bq + (qhi - qlo)]
            mine.copy_(Qslice_h, non_blocking=True)                                   # H2D: this rank's slice only
            dist.all_gather_into_tensor(Qpad, Qpad[rank * bq:(rank + 1) * bq])         # NVLink: everybody's slices
            if B % world == 0:
                qptr = Qpad.data_ptr()
            else:                                                                      # ragged slices: close the gaps
                for g in range(world):
                    glo, ghi = DD.query_slice(B, g, world)
                    Qd[glo:ghi].copy_(Qpad[g * bq:g * bq + (ghi - glo)], non_blocking=True)
                qptr = Qd.data_ptr()
        ri, rd = step(qptr)
        oi_h.copy_(ri[:qhi - qlo], non_blocking=True); od_h.copy_(rd[:qhi - qlo], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if world > 1:                                           # the gathered batch is the batch (checked once, outside the timed region)
        step_e2e()
        torch.cuda.synchronize(dev)
        got = Qpad if B % world == 0 else Qd
        assert torch.equal(got[:B], Q), "all-gathered query batch differs from the batch"
    e2e = {"value": round(B * a.steps / float(te.item()), 1), "unit": UNIT,
           "h2d_bytes_per_step": int((qhi - qlo) * a.dim * 4 * world),       # all ranks together: the batch crosses PCIe once
           "d2h_bytes_per_step": int(B * k * 8),
           "nvlink_allgather_bytes_per_step": 0 if world == 1 else int(B * a.dim * 4)}
    if rank == 0:
        emit({"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
              "ms_per_step": round(ms_total / a.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": DTYPES[a.lut], "data": "synthetic",
              "config": {"workload": f"index-sharded (BASELINE configs[4] shape): {world} shard(s) x {a.n} x {a.dim} synthetic unit-norm = corpus "
                                     f"{world * a.n}, per-shard Vamana R={a.R} (GPU-built) + PQ M={a.M}, L={a.L}, W={a.W}, table={a.lut}, rerank, "
                                     f"top-{k}; every GPU searches the whole {B}-query batch on its shard",
                         "exchange": ("none (1 shard)" if world == 1 else
                                      ("search-kernel epilogue stores packed (dist,id) keys into the owner rank's buffer over NVLink peer memory, "
                                       "barrier, k-way merge kernel" if peer is not None else
                                       "device-packed (dist,id) keys, ONE NCCL all_to_all_single, k-way merge kernel")),
                         "exchange_bytes_per_rank_per_step": 0 if world == 1 else B * k * 8,
                         "queries_per_step_total": B, "index": "sharded", "queries": "replicated", "recall_at_10": round(rec, 4),
                         "recall_is": "global: merged top-10 vs the exact top-10 over the whole sharded corpus", "recall_queries": ng,
                         "l2_flush": f"inputs larger than L2 (shard {(a.n * (a.dim * 4 + a.R * 4 + a.M)) / 1e9:.1f} GB per GPU)",
                         "mean_hops": round(float(h_np.mean()), 1), "mean_visited": round(float(v_np.mean()), 1), "setup": info,
                         "cpu_affinity": affinity},
              "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
              "cpu_baseline": None})
    if peer is not None:
        peer.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_port(a, X, adj, codes, cb, med, Q, ids_gpu):
    """The oracle (C restatement, OpenMP over queries) on the host cores, bounded sample of the same workload."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as O
    O.build()
    Xh = X.cpu().numpy(); adjh = adj.cpu().numpy().view(np.uint32); ch = codes.cpu().numpy(); cbh = cb.cpu().numpy()
    cores = O.num_threads()
    n = min(Q.shape[0], 4096 * cores)          # ~10-20 s of CPU work at ~6k QPS on 16 cores
    Qs = Q[:n].cpu().numpy()
    t = time.perf_counter()
    ids, d, hops, vis = O.search_batch(adjh, Xh, Qs, med, a.L, a.k, codes=ch, codebook=cbh,
                                       dist_mode=O.DIST_ADC_U8 if a.lut != "f32" else (O.DIST_ADC_TREE if a.adc == "tree" else O.DIST_ADC_SEQ),
                                       flavor=O.FLAVOR_WARP,
                                       W=a.W, rerank_=True, w_after_empty=a.W2 if a.lut != "f32" else 0)
    dt = time.perf_counter() - t
    same = float(np.mean(np.all(ids == ids_gpu[:n], axis=1)))
    return {"value": round(n / dt, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} queries of the same batch on the same index, oracle/oracle.c with OpenMP over queries; "
                      f"{same * 100:.2f}% of its top-{a.k} lists are identical to the GPU's"}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref = pydiskann compiled from /root/reference)
# ------------------------------------------------------------------------------------------------------
_W = {}


def _ref_worker_init():
    pass


def _ref_search(qi):
    g, vg, cu, Q, vec, L, k, med = _W["g"], _W["vg"], _W["cu"], _W["Q"], _W["vec"], _W["L"], _W["k"], _W["med"]
    q = Q[qi]
    g._distance_table_cache.clear()
    ids = cu.greedy_search_cython(g, med, q, L, vg.compute_query_distance)          # variant A, PQ ADC callback
    d2 = [float(np.sum((vec[i] - q) * (vec[i] - q))) for i in ids]                   # search_engine.py:374-379
    order = np.argsort(np.array(d2, np.float32), kind="stable")[:k]
    return [int(ids[i]) for i in order]


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    import torch
    sys.path.insert(0, str(ROOT / "oracle"))
    import ref_loader
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    X, adj, deg, codes, cb, med, info = build_index(a, dev)                          # setup only: same index as our arm
    from diskrag_b200.synth import synth_torch
    cores = os.cpu_count() or 1
    per_step = 2 * cores
    nq = per_step * (a.steps + a.warmup)
    Q = synth_torch(nq, a.dim, seed=20242, sample_seed=1000, device=dev).cpu().numpy()
    gt = ground_truth(X, torch.from_numpy(Q).to(dev), a.k)
    vec = X.cpu().numpy(); adjh = adj.cpu().numpy().view(np.uint32); ch = codes.cpu().numpy(); cbh = cb.cpu().numpy()
    del X, adj, codes
    torch.cuda.empty_cache()
    kind = "reference"
    if ref_loader.available():
        m = ref_loader.load()
        vg, cu, fp = m["vamana_graph"], m["cython_utils"], m["fast_pq"]
        from diskrag_b200.pq.fast_pq import _wrap_kmeans
        pq = fp.DiskANNPQ(a.M, 256)
        pq.sub_dim = a.dim // a.M; pq.is_fitted = True
        pq.kmeans_list = [_wrap_kmeans(cbh[i], 42 + i) for i in range(a.M)]
        g = vg.VamanaGraphWithPQ(a.R, pq)

        class LazyNodes(dict):                   # the reference's dict of Node objects, materialised on first touch
            def __missing__(self, i):
                n = vg.Node(i, vec[i], ch[i])
                n.neighbors = [int(x) for x in adjh[i]]
                self[i] = n
                return n
        g.nodes = LazyNodes()
        g.medoid_idx = med
        g.use_pq_for_search = True
        _W.update(g=g, vg=vg, cu=cu, Q=Q, vec=vec, L=a.L, k=a.k, med=med)
        pool = mp.get_context("fork").Pool(cores)
        run = lambda qs: pool.map(_ref_search, qs, chunksize=1)
        sample = (f"{per_step} queries per step ({2} per core) of the same workload on the same index; the reference's own "
                  f"greedy_search_cython + compute_query_distance (PQ ADC) + numpy exact rerank, {cores} processes")
    else:
        kind = "port"
        import oracle as O
        O.build()
        per_step = 64 * cores
        nq = per_step * (a.steps + a.warmup)
        Q = np.concatenate([Q] * (nq // Q.shape[0] + 1))[:nq]
        gt = np.concatenate([gt] * (nq // gt.shape[0] + 1))[:nq]
        run = lambda qs: O.search_batch(adjh, vec, Q[qs], med, a.L, a.k, codes=ch, codebook=cbh, dist_mode=O.DIST_ADC_SEQ,
                                        flavor=O.FLAVOR_NUMPY, W=1, rerank_=True)[0].tolist()
        sample = f"{per_step} queries per step, oracle/oracle.c (C restatement) with OpenMP, {cores} threads"
    res = []
    for s in range(a.warmup):
        res += run(list(range(s * per_step, (s + 1) * per_step)))
    t0 = time.perf_counter()
    for s in range(a.warmup, a.warmup + a.steps):
        res += run(list(range(s * per_step, (s + 1) * per_step)))
    dt = time.perf_counter() - t0
    value = per_step * a.steps / dt
    rec = recall(np.array(res, np.int64), gt[:len(res)], a.k)
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": round(dt / a.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{a.n}x{a.dim} synthetic unit-norm, Vamana R={a.R} (GPU-built, Lbuild={a.Lbuild}, alpha=1.2), "
                                  f"PQ M={a.M}, L={a.L}, W=1, adc=seq, rerank, top-{a.k}",
                      "queries_per_step": per_step, "recall_at_10": round(rec, 4), "setup": info,
                      "note": "index built by the GPU builder outside the timed region (the reference's builder needs hours at 1M)"},
           "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else a library prints there (NCCL's version banner at
    communicator creation, for one) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


if __name__ == "__main__":
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "index-sharded":
        run_index_sharded(args)
    else:
        run_ours(args)
