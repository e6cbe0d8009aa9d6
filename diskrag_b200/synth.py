"""Synthetic embeddings of the shape BASELINE.json names (no OpenAI key, no network).

SURVEY.md §8(d): iid Gaussian data cannot reach recall@10 >= 0.95 with any graph/PQ search, so the
generator is a low-rank Gaussian mixture lifted to D dims, plus a little isotropic noise, then
L2-normalised (OpenAI embeddings are unit-norm).  Frozen parameters: r = 64 latent dims,
K = 4096 cluster centres, within-cluster sigma 0.35, noise 0.05/sqrt(D).

  numpy path  (tests, small configs)   : synth_numpy(n, D, seed)
  torch path  (bench, 1M x 1536 on GPU): synth_torch(n, D, seed, device)

The two paths use different RNG streams; each is reproducible from its seed.  Queries are held-out
draws from the same mixture (same centres / lift, different sample seed).
"""
import numpy as np

R_LATENT = 64
K_CLUSTERS = 4096
SIGMA = 0.35
NOISE = 0.05


def _model_numpy(D, seed, r=R_LATENT, K=K_CLUSTERS):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((K, r)).astype(np.float32)
    A, _ = np.linalg.qr(rng.standard_normal((D, r)))  # D x r, orthonormal columns
    return centres, A.T.astype(np.float32)             # lift: r x D


def synth_numpy(n, D, seed=20240, sample_seed=0, r=R_LATENT, K=K_CLUSTERS, sigma=SIGMA, noise=NOISE):
    """-> float32[n, D], unit-norm rows.  (seed fixes the mixture, sample_seed the draws.)"""
    r = min(r, D)
    centres, lift = _model_numpy(D, seed, r, K)
    rng = np.random.default_rng([seed, 7919, sample_seed])
    out = np.empty((n, D), np.float32)
    step = 65536
    for s in range(0, n, step):
        m = min(step, n - s)
        c = rng.integers(0, centres.shape[0], m)
        z = centres[c] + sigma * rng.standard_normal((m, centres.shape[1])).astype(np.float32)
        x = z @ lift + (noise / np.sqrt(D)) * rng.standard_normal((m, D)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        out[s:s + m] = x
    return out


def synth_torch(n, D, seed=20240, sample_seed=0, device="cuda", r=R_LATENT, K=K_CLUSTERS, sigma=SIGMA,
                noise=NOISE, out=None):
    """Same mixture family generated on `device` with torch RNG (plumbing for the bench input only)."""
    import torch
    r = min(r, D)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    centres = torch.randn(K, r, generator=g, device=device)
    # the orthonormal lift is tiny (D x r): numpy QR on the host avoids loading cuSOLVER (~50 s cold)
    A, _ = np.linalg.qr(np.random.default_rng(seed).standard_normal((D, r)))
    lift = torch.from_numpy(np.ascontiguousarray(A.T.astype(np.float32))).to(device)
    g.manual_seed(seed * 1000003 + 7919 * (sample_seed + 1))
    if out is None:
        out = torch.empty(n, D, device=device, dtype=torch.float32)
    step = 131072
    for s in range(0, n, step):
        m = min(step, n - s)
        c = torch.randint(0, K, (m,), generator=g, device=device)
        z = centres[c] + sigma * torch.randn(m, r, generator=g, device=device)
        x = z @ lift + (noise / D ** 0.5) * torch.randn(m, D, generator=g, device=device)
        x /= x.norm(dim=1, keepdim=True)
        out[s:s + m] = x
    return out
