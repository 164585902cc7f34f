"""Synthetic feature matrices of the benchmark shapes.

Latent slow processes: S independent AR(1) chains ``z_k(t+1) = rho_k z_k(t) + sqrt(1-rho_k^2) eps``
with ``rho_k = exp(-1/T_k)``, ``T_k = T0 * 2^-k`` -> well-separated TICA eigenvalues.  The chains
are a function of (seed, global frame index) only -- generated on the host with a linear filter
for the whole series, so every sharding sees the same latent trajectory -- and are mixed into F
features on the device: ``X = (Z A + noise * eps_f) * s + m`` with per-feature scale
``s ~ U(0.05, 0.5)`` and offset ``m ~ U(0.5, 3.0)`` (like nm distances: |mean| >> std stresses the
standardisation).  float32, row-major.
"""
from __future__ import annotations

import numpy as np
import torch


_CHAINS = {}


def latent_chains(n_total: int, n_slow: int, seed: int = 0, t0: float = 4000.0) -> np.ndarray:
    """(n_total x n_slow) latent AR(1) chains; the last result is cached (callers generate a shard in
    row chunks, and the whole-series filter costs seconds at 10M frames)."""
    key = (n_total, n_slow, seed, t0)
    if key not in _CHAINS:
        _CHAINS.clear()
        _CHAINS[key] = _latent_chains(n_total, n_slow, seed, t0)
    return _CHAINS[key]


def _latent_chains(n_total: int, n_slow: int, seed: int = 0, t0: float = 4000.0) -> np.ndarray:
    from scipy.signal import lfilter
    rng = np.random.default_rng(seed)
    T = t0 * 2.0 ** (-np.arange(n_slow))
    rho = np.exp(-1.0 / T)
    z = np.empty((n_total, n_slow), dtype=np.float64)
    for k in range(n_slow):
        eps = rng.standard_normal(n_total)
        eps[0] /= np.sqrt(1.0 - rho[k] ** 2)          # start in the stationary distribution
        z[:, k] = lfilter([np.sqrt(1.0 - rho[k] ** 2)], [1.0, -rho[k]], eps)
    return z


NOISE_CHUNK = 1 << 16       # rows per noise block; blocks sit on a GLOBAL grid (multiples of this)


def feature_matrix(n_total: int, f: int, start: int, stop: int, device, seed: int = 0,
                   n_slow: int = 8, noise: float = 0.5, out: torch.Tensor = None) -> torch.Tensor:
    """Rows [start, stop) of the synthetic (n_total x f) matrix, as a float32 tensor on ``device``.

    Every row is a pure function of (seed, f, global frame index): the feature noise is drawn in
    blocks of NOISE_CHUNK rows on a global grid, block g from a generator seeded with (seed, g), and a
    shard takes the rows of the blocks it overlaps -- so any sharding of the series (and any
    ``start``) reproduces the same rows bit for bit on the same device type."""
    z = latent_chains(n_total, n_slow, seed)[start:stop]
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    A = torch.randn((n_slow, f), generator=g, dtype=torch.float32).to(device)
    s = (torch.rand(f, generator=g) * 0.45 + 0.05).to(device)
    m = (torch.rand(f, generator=g) * 2.5 + 0.5).to(device)
    X = out if out is not None else torch.empty((stop - start, f), dtype=torch.float32, device=device)
    dg = torch.Generator(device=device)
    for blk in range(start // NOISE_CHUNK, (max(stop, start + 1) - 1) // NOISE_CHUNK + 1):
        b0 = blk * NOISE_CHUNK
        g0, g1 = max(start, b0), min(stop, b0 + NOISE_CHUNK)
        if g1 <= g0:
            continue
        dg.manual_seed(seed * 1_000_003 + blk)                 # keyed by the global block index
        e = torch.randn((NOISE_CHUNK, f), generator=dg, dtype=torch.float32, device=device)[g0 - b0:g1 - b0]
        zc = torch.from_numpy(z[g0 - start:g1 - start]).to(device=device, dtype=torch.float32)
        # zc @ A as n_slow elementwise rank-1 updates in a fixed order (not a GEMM, whose summation
        # order depends on the shape heuristics): the rows must not depend on the shard shape
        e *= noise
        for k in range(n_slow):
            e.addcmul_(zc[:, k:k + 1], A[k:k + 1])
        X[g0 - start:g1 - start] = e * s + m
    return X


def cluster_points(n: int, d: int, k: int, device, seed: int = 2, sigma: float = 0.03,
                   dtype=torch.float32, start: int = 0) -> torch.Tensor:
    """C5: mixture of k Gaussians in [-1, 1]^d (centres U[-0.9, 0.9]^d, seed ``seed``)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cent = (torch.rand((k, d), generator=g, dtype=torch.float64) * 1.8 - 0.9).to(device)
    dg = torch.Generator(device=device).manual_seed(seed * 7919 + start)
    idx = torch.randint(k, (n,), generator=dg, device=device)
    Y = cent[idx] + sigma * torch.randn((n, d), generator=dg, dtype=torch.float64, device=device)
    return Y.to(dtype).contiguous()
