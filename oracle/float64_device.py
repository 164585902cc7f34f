"""TEST INFRASTRUCTURE ONLY -- float64 checker of the CV hot path in plain torch, on whatever device
its inputs live on (so the BASELINE-size configurations can be checked on the GPU box in seconds).

Same semantics as ``oracle/cv_oracle.py`` (numpy float64, pinned against the reference's golden
artefacts; ``tests/test_oracle_golden.py::test_float64_device_checker_equals_numpy_oracle`` pins this
file to it): float32 IEEE standardisation ``(x - mean) / range`` (reference cv_calculator.py:834-835),
M = N - lag pairs (mlcolvar create_timelagged_dataset, call site :2244-2247), float64 raw sums,
mlcolvar ``TICA.compute`` / ``cholesky_eigh`` post-processing (call site :2257-2261), hTICA
composition (:2311-2384), projection + CV min-max normalisation (:918-991).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (as the CHECKER of its parity block and
never as the thing timed) import this module; the product package never does.
"""
from __future__ import annotations

from typing import Optional

import torch


def standardize_f32(X: torch.Tensor, mean: Optional[torch.Tensor], rng: Optional[torch.Tensor]) -> torch.Tensor:
    """float32 ``(x - mean) / range`` exactly as the reference's in-place sub_/div_ (:834-835)."""
    if mean is None:
        return X
    return (X - mean.to(X.dtype)) / rng.to(X.dtype)


def lagged_sums(X: torch.Tensor, lag: int, mean=None, rng=None, chunk: int = 100_000, cols=None) -> dict:
    """float64 raw sums over the M = n - lag pairs of the float32-standardised rows, chunked over
    frames: S0 = sum z_t z_t^T, St = sum z_t z_{t+lag}^T, a = sum_{t<M} z_t, b = sum_{t>=lag} z_t.
    ``cols`` = (c0, c1) restricts to a column block."""
    n, f = X.shape
    M = n - lag
    if cols is not None:
        X = X[:, cols[0]:cols[1]]
        mean = None if mean is None else mean[cols[0]:cols[1]]
        rng = None if rng is None else rng[cols[0]:cols[1]]
        f = X.shape[1]
    dev = X.device
    S0 = torch.zeros((f, f), dtype=torch.float64, device=dev)
    St = torch.zeros((f, f), dtype=torch.float64, device=dev)
    a = torch.zeros(f, dtype=torch.float64, device=dev)
    b = torch.zeros(f, dtype=torch.float64, device=dev)
    for s0 in range(0, M, chunk):
        e0 = min(M, s0 + chunk)
        Z = standardize_f32(X[s0:e0 + lag], mean, rng).double()
        zt, zl = Z[:e0 - s0], Z[lag:lag + e0 - s0]
        S0.addmm_(zt.T, zt)
        St.addmm_(zt.T, zl)
        a += zt.sum(0)
        b += zl.sum(0)
    return {"S0": S0, "St": St, "a": a, "b": b, "M": M}


def tica_from_sums(S0, St, a, b, M: int, out: int, reg: float = 1e-6):
    """mlcolvar TICA.compute (remove_average=True) + cholesky_eigh, dense, float64."""
    mu, nu = a / M, b / M
    C0 = S0 / M - torch.outer(mu, mu)
    C0 = 0.5 * (C0 + C0.T)
    Ct = St / M - torch.outer(mu, nu)
    Ct = 0.5 * (Ct + Ct.T)
    F = C0.shape[0]
    L = torch.linalg.cholesky(C0 + reg * torch.eye(F, dtype=C0.dtype, device=C0.device))
    Y = torch.linalg.solve_triangular(L, Ct, upper=False)
    A = torch.linalg.solve_triangular(L, Y.T.contiguous(), upper=False).T
    ev, U = torch.linalg.eigh(0.5 * (A + A.T))
    out = min(out, F)
    ev, U = ev.flip(0)[:out], U.flip(1)[:, :out]
    V = torch.linalg.solve_triangular(L.T.contiguous(), U.contiguous(), upper=True)
    V = V / torch.linalg.norm(V, dim=0, keepdim=True)
    V = V * torch.sign(V[0:1, :])
    return ev, V


def htica_chunks(F: int, num_subspaces: int):
    w = F // num_subspaces
    return [] if w == 0 else [(s, min(s + w, F)) for s in range(0, F, w)]


def htica(X: torch.Tensor, lag: int, mean, rng, num_subspaces: int, sub_dim: int, d: int,
          reg: float = 1e-6, chunk: int = 100_000):
    """Reference hTICA (:2311-2384) without ever forming the full F x F Gram: per-block float64
    sums -> level-1 TICA -> T1; level-2 sums of the UNCENTRED level-1 projections p = z T1 (float64
    projections of the float32-standardised rows) -> TICA -> W = T1 V2.  Returns (W, T1, V2)."""
    n, F = X.shape
    chunks = htica_chunks(F, num_subspaces)
    blocks = []
    for (c0, c1) in chunks:
        s = lagged_sums(X, lag, mean, rng, chunk, cols=(c0, c1))
        blocks.append(tica_from_sums(s["S0"], s["St"], s["a"], s["b"], s["M"], sub_dim, reg)[1])
    S1 = sum(v.shape[1] for v in blocks)
    T1 = torch.zeros((F, S1), dtype=torch.float64, device=X.device)
    c = 0
    for (c0, c1), Vb in zip(chunks, blocks):
        T1[c0:c1, c:c + Vb.shape[1]] = Vb
        c += Vb.shape[1]
    M = n - lag
    S0 = torch.zeros((S1, S1), dtype=torch.float64, device=X.device)
    St = torch.zeros_like(S0)
    a = torch.zeros(S1, dtype=torch.float64, device=X.device)
    b = torch.zeros_like(a)
    for s0 in range(0, M, chunk):
        e0 = min(M, s0 + chunk)
        P = standardize_f32(X[s0:e0 + lag], mean, rng).double() @ T1
        pt, pl = P[:e0 - s0], P[lag:lag + e0 - s0]
        S0.addmm_(pt.T, pt); St.addmm_(pt.T, pl); a += pt.sum(0); b += pl.sum(0)
    _, V2 = tica_from_sums(S0, St, a, b, M, d, reg)
    return T1 @ V2, T1, V2


def project_normalized(X: torch.Tensor, mean, rng, W: torch.Tensor, chunk: int = 100_000):
    """float64 ``P = Z W`` of the float32-standardised rows, its per-column min / max and the
    [-1, 1] normalised projection (:974-991, :918-972).  Returns (Pn, cmin, cmax)."""
    W = W.double()
    P = torch.empty((X.shape[0], W.shape[1]), dtype=torch.float64, device=X.device)
    for s0 in range(0, X.shape[0], chunk):
        P[s0:s0 + chunk] = standardize_f32(X[s0:s0 + chunk], mean, rng).double() @ W
    mn, mx = P.min(0).values, P.max(0).values
    return (P - (mx + mn) / 2) / ((mx - mn) / 2), mn, mx


def eigvec_error(V: torch.Tensor, Vref: torch.Tensor) -> float:
    """max over columns of ||v - v_ref||_2 up to sign (columns have unit L2 norm)."""
    V, Vref = V.double(), Vref.double()
    sgn = torch.sign((V * Vref).sum(0, keepdim=True))
    return float(torch.linalg.norm(V * sgn - Vref, dim=0).max().item())


def sums_rel_error(s: dict, ref: dict) -> dict:
    """Normwise (max-abs relative to max |ref|) errors of S0 (upper triangle) and St."""
    out = {}
    S0, R0 = torch.triu(s["S0"]), torch.triu(ref["S0"])
    out["S0"] = float(((S0 - R0).abs().max() / R0.abs().max()).item())
    if s.get("St") is not None and ref.get("St") is not None:
        out["St"] = float(((s["St"] - ref["St"]).abs().max() / ref["St"].abs().max()).item())
        sy, ry = s["St"] + s["St"].T, ref["St"] + ref["St"].T
        out["St_sym"] = float(((sy - ry).abs().max() / ry.abs().max()).item())
    return out
