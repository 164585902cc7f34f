"""Small F x F eigen stage, FP64 on one device (torch.linalg -> cuSOLVER).

Takes the raw FP64 sums produced by ``ops.lagged_covariance`` and reproduces the
post-processing of mlcolvar's ``TICA.compute`` / ``cholesky_eigh`` and sklearn's PCA as the
reference calls them (cv_calculator.py:2194-2215, 2249-2267, 2311-2384).  SURVEY appendix A.1/A.2.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def _cholesky_eigh(C0: torch.Tensor, Ct: torch.Tensor, reg: float, out: int):
    """B = C0 + reg I; L = chol(B); A = L^-1 Ct L^-T; eigh; descending; V = L^-T U;
    unit-L2 columns; sign(row 0) >= 0; first ``out`` (mlcolvar cholesky_eigh + TICA)."""
    F = C0.shape[0]
    B = C0 + reg * torch.eye(F, dtype=C0.dtype, device=C0.device)
    L, info = torch.linalg.cholesky_ex(B)
    if int(info.item()) != 0:
        raise RuntimeError("TICA: C0 + reg*I is not positive definite")
    # A = L^-1 Ct L^-T by two triangular solves
    Y = torch.linalg.solve_triangular(L, Ct, upper=False)             # L^-1 Ct
    A = torch.linalg.solve_triangular(L, Y.T, upper=False).T          # (L^-1 (L^-1 Ct)^T)^T
    A = 0.5 * (A + A.T)
    evals, U = torch.linalg.eigh(A)
    evals = evals.flip(0)
    U = U.flip(1)
    out = min(out, F)
    V = torch.linalg.solve_triangular(L.T, U[:, :out], upper=True)    # L^-T U
    V = V / torch.linalg.norm(V, dim=0, keepdim=True)
    V = V * torch.sign(V[0:1, :])
    return evals[:out], V


def tica_from_sums(S0, St, a, b, M: int, out: int, reg: float = 1e-6):
    """TICA eigenpairs from raw sums.  ``S0`` must be fully symmetric.  mu = a/M is subtracted
    from BOTH series, covariances are divided by M, C_tau is symmetrised."""
    mu = a / M
    nu = b / M
    C0 = S0 / M - torch.outer(mu, mu)
    C0 = 0.5 * (C0 + C0.T)
    Ct = St / M - torch.outer(mu, nu)
    Ct = 0.5 * (Ct + Ct.T)
    return _cholesky_eigh(C0, Ct, reg, out)


def pca_from_sums(S_all, s_all, N: int, d: int):
    """sklearn covariance-eigh PCA + the reference's sign rule (cv_calculator.py:2204-2215)."""
    mu = s_all / N
    Cm = (S_all - N * torch.outer(mu, mu)) / (N - 1)
    Cm = 0.5 * (Cm + Cm.T)
    evals, V = torch.linalg.eigh(Cm)
    evals = evals.flip(0)[:d]
    W = V.flip(1)[:, :d].clone()
    neg = W[0, :] < 0
    W[:, neg] = -W[:, neg]
    return evals, W


def htica_chunks(F: int, num_subspaces: int) -> List[Tuple[int, int]]:
    """Column chunks exactly as torch.split(x, F // num_subspaces, dim=1) makes them
    (cv_calculator.py:2331-2334)."""
    w = F // num_subspaces
    if w == 0:
        return []
    return [(s, min(s + w, F)) for s in range(0, F, w)]


def htica_level1(S0, St, a, b, M: int, chunks, sub_dim: int, reg: float = 1e-6) -> torch.Tensor:
    """Per-block TICA (level 1); returns the block-diagonal transform T1 (F x S1)."""
    F = S0.shape[0]
    blocks = []
    for (s, e) in chunks:
        _, Vb = tica_from_sums(S0[s:e, s:e], St[s:e, s:e], a[s:e], b[s:e], M, sub_dim, reg)
        blocks.append(Vb)
    S1 = sum(v.shape[1] for v in blocks)
    T1 = torch.zeros((F, S1), dtype=S0.dtype, device=S0.device)
    c = 0
    for (s, e), Vb in zip(chunks, blocks):
        T1[s:e, c:c + Vb.shape[1]] = Vb
        c += Vb.shape[1]
    return T1


def htica_from_full_sums(S0, St, a, b, M: int, num_subspaces: int, sub_dim: int, d: int,
                         reg: float = 1e-6):
    """hTICA from the FULL F x F sums (single data pass): level-2 sums are T1^T S T1
    (level-1 projections are uncentred; level 2 re-centres).  Returns (W, T1, V2)."""
    chunks = htica_chunks(S0.shape[0], num_subspaces)
    if not chunks:
        raise ValueError("num_subspaces larger than number of features")
    T1 = htica_level1(S0, St, a, b, M, chunks, sub_dim, reg)
    _, V2 = tica_from_sums(T1.T @ S0 @ T1, T1.T @ St @ T1, T1.T @ a, T1.T @ b, M, d, reg)
    return T1 @ V2, T1, V2
