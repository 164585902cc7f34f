"""The reference's CPU path for the hot path, in its own float32 arithmetic, for timing.

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Used by ``bench.py``'s
``cpu_baseline`` leg and ``bench.py --impl reference``.

This is the "reference-faithful" mode described in BASELINE.md section 3: the same library calls
the reference makes on its linear-CV path, on the same inputs --

  pandas ``.agg(['mean','std','min','max'])``          (cv_calculator.py:295-297)
  in-place torch ``sub_`` / ``div_``                    (cv_calculator.py:834-835)
  pairs ``x_t = X[:N-lag]``, ``x_lag = X[lag:]``        (mlcolvar create_timelagged_dataset, as VIEWS)
  ``mu = mean(x_t)``; GEMM-form C0 / C_tau, symmetrise  (mlcolvar TICA.compute; the library's
      ``einsum('ij,ik,i->jk')`` materialises M x F x F without opt_einsum and cannot run at
      these sizes -- SURVEY section 6 -- so the mathematically identical GEMM form is timed)
  ``+1e-6 I``, ``torch.linalg.cholesky / inv / eigh``   (mlcolvar cholesky_eigh)
  ``X @ W``, min / max, ``[-1, 1]`` normalisation       (cv_calculator.py:974-991, 918-972)
  ``sklearn.cluster.KMeans(init=ndarray, n_init=1)``    (statistics.py:189-195)

mlcolvar itself is not installed in this image; the reference package cannot be imported
end-to-end (MDAnalysis / lightning / mlcolvar missing), so ``kind`` is "port".
"""
from __future__ import annotations

import os
import time
import warnings
from typing import Dict

import numpy as np
import pandas as pd
import torch


def cpu_threads() -> int:
    return os.cpu_count() or 1


def run_reference_pipeline(X: np.ndarray, lag: int, d: int, k: int, kmeans_iters: int,
                           timings: bool = True) -> Dict:
    """One pass of the hot path on the host cores.  ``X`` is (n x f) float32 (modified in place,
    like the reference).  Returns per-stage seconds and the results."""
    torch.set_num_threads(cpu_threads())
    t = {}
    n, f = X.shape
    t0 = time.perf_counter()
    df = pd.DataFrame(X, copy=False)
    stats = df.agg(["mean", "std", "min", "max"]).T
    mean = stats["mean"].to_numpy().astype(np.float32)
    std = stats["std"].to_numpy().astype(np.float32)
    std[np.abs(std) < 1e-8] = 1.0
    t["stats"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    data = torch.from_numpy(X)
    data.sub_(torch.from_numpy(mean))
    data.div_(torch.from_numpy(std))
    t["standardize"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    x_t, x_lag = data[: n - lag], data[lag:]
    M = n - lag
    mu = x_t.mean(dim=0)
    # centred Gram in GEMM form without materialising centred copies
    S0 = x_t.T @ x_t
    St = x_t.T @ x_lag
    sb = x_lag.sum(dim=0)
    C0 = S0 / M - torch.outer(mu, mu)
    Ct = St / M - torch.outer(mu, sb / M)
    C0 = 0.5 * (C0 + C0.T)
    Ct = 0.5 * (Ct + Ct.T)
    t["covariance"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    L = torch.linalg.cholesky(C0 + 1e-6 * torch.eye(f))
    Li = torch.linalg.inv(L)
    A = Li @ Ct @ Li.T
    evals, U = torch.linalg.eigh(0.5 * (A + A.T))
    evals, U = evals.flip(0), U.flip(1)
    V = Li.T @ U[:, :d]
    V = V / torch.linalg.norm(V, dim=0, keepdim=True)
    V = V * torch.sign(V[0:1, :])
    t["eigen"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    P = data @ V                                        # normalize_cv (cv_calculator.py:981)
    mn, mx = P.min(dim=0).values, P.max(dim=0).values
    P = data @ V                                        # project_data runs the GEMM again (:958)
    P.sub_((mx + mn) / 2).div_((mx - mn) / 2)
    t["projection"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    from sklearn.cluster import KMeans
    Y = P.numpy().astype(np.float64)                    # CSV hand-off is float64 (traj_cluster :202)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=k, random_state=0, init=Y[:k].copy(), n_init=1,
                    max_iter=kmeans_iters, tol=0.0).fit(Y)
    t["kmeans"] = time.perf_counter() - t0
    t["kmeans_iters"] = int(km.n_iter_)
    t["total"] = sum(v for key, v in t.items() if key not in ("kmeans_iters",))
    return {"timings": t, "evals": evals[:d].numpy(), "weights": V.numpy(), "labels": km.labels_}
