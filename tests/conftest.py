import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _device_count():
    from diskrag_b200._lib import device_count
    return int(device_count())


def _box_has_nvidia_gpu():
    import shutil
    import subprocess
    if not shutil.which("nvidia-smi"):
        return False
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout.count("GPU ") > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a box without CUDA skips the device tests instead of erroring in them (the product itself has no CPU fallback:
    # every compute entry point fails loudly there, which tests/test_abi.py checks)
    if any(it.get_closest_marker("gpu") for it in items) and _device_count() == 0:
        # never skip silently on a GPU box: if the driver sees a GPU and the library does not, that is a failure of the product
        if _box_has_nvidia_gpu():
            raise pytest.UsageError("nvidia-smi lists a GPU but libdiskrag_b200.so reports no CUDA device: refusing to skip the device tests")
        skip = pytest.mark.skip(reason="no CUDA device on this box (device tests run with -m gpu on the B200 box)")
        for it in items:
            if it.get_closest_marker("gpu"):
                it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    z = np.load(ROOT / "tests" / "golden" / "ref_small.npz")
    g = {k: z[k] for k in z.files}
    for k in ("N", "D", "M", "R", "medoid", "exp_medoid_500"):
        g[k] = int(g[k])
    rec32 = g["records"].view(np.uint32).reshape(g["N"], g["D"] + g["R"])
    g["vec"] = rec32[:, :g["D"]].copy().view(np.float32)
    g["adj"] = rec32[:, g["D"]:].copy()
    return g


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


def canon(ids, d):
    """Sort a result list by (distance, id): the canonical order for comparing lists that may hold
    exact distance ties (the reference's own order inside a tie group is heap-layout dependent)."""
    ids = np.asarray(ids); d = np.asarray(d)
    keep = ids >= 0
    ids, d = ids[keep], d[keep]
    o = np.lexsort((ids, d))
    return ids[o], d[o]


def make_case(orc, N, D, M, R, L, seed, nq=32, dup=0, K=64, r=16, train=None):
    """Seeded synthetic index built entirely by the oracle (sequential Vamana + a cheap PQ): no
    reference and no GPU needed, so the same case exists on the GPU box."""
    from diskrag_b200.synth import synth_numpy
    X = synth_numpy(N, D, seed=seed, K=K, r=min(r, D))
    Q = synth_numpy(nq, D, seed=seed, sample_seed=1, K=K, r=min(r, D))
    if dup:
        X[N - dup:] = X[:dup]
    rng = np.random.default_rng(seed)
    ds = D // M
    # codebook: 256 random training sub-vectors per subspace + 3 Lloyd steps in numpy (quality is irrelevant here)
    cb = np.empty((M, 256, ds), np.float32)
    for m in range(M):
        sub = X[:, m * ds:(m + 1) * ds]
        c = sub[rng.choice(N, 256, replace=N < 256)].copy()
        for _ in range(3):
            d2 = ((sub[:, None, :] - c[None, :, :]) ** 2).sum(-1) if N * 256 * ds < 4e7 else None
            if d2 is None:
                break
            a = d2.argmin(1)
            for j in range(256):
                if (a == j).any():
                    c[j] = sub[a == j].mean(0)
        cb[m] = c
    codes = orc.pq_encode(cb, X)
    s0 = rng.permutation(N).astype(np.int32); s1 = rng.permutation(N).astype(np.int32)
    med = orc.medoid(X, rng.choice(N, min(N, 64), replace=False).astype(np.int32))
    rows = orc.vamana_build(X, R, L, 1.2, med, s0, s1)
    adj = np.zeros((N, R), np.uint32)          # 0-padding exactly like DiskANNPersist.save_index
    for i, row in enumerate(rows):
        adj[i, :len(row)] = row[:R]
    return dict(X=X, Q=Q, codebook=cb, codes=codes, adj=adj, medoid=int(med), N=N, D=D, M=M, R=R)
