"""CPU oracle: numpy restatement of deep_cartograph's CV hot path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The product package
never imports this module.

Every function cites the reference lines it restates (paths relative to
``/root/reference/deep_cartograph/``).  Two third-party libraries carry the
arithmetic in the reference and are NOT vendored there:

* ``mlcolvar==1.2.2`` (``environment_detailed.yml:305``) -- absent from this
  image.  ``create_timelagged_dataset``, ``TICA.compute`` (``correlation_matrix``,
  ``cholesky_eigh``) and ``ReduceEigenvaluesLoss`` are restated from the
  published algorithm and pinned through the reference's golden artefacts
  (``tests/data/input/models/{tica,htica}_model.zip``), see
  ``tests/test_oracle_golden.py``.
* ``scikit-learn`` (pinned 1.6.1, 1.9.0 installed here) -- ``PCA`` and ``KMeans``
  are restated AND cross-checked against the installed library.

Pinning status:
  stats / standardise / projection / CV normalisation : pinned (golden npy + csv)
  PCA, TICA, hTICA weights                            : pinned (golden cv_weights.npy)
  KMeans                                              : parity UNPINNED by the reference's
      own tests (no reference test selects kmeans); pinned here against
      scikit-learn's KMeans called exactly as ``statistics.kmeans_clustering`` does.
  find_centroids                                      : pinned (golden traj_cluster csv)
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "column_stats", "prepare_normalization", "standardize", "lagged_pairs",
    "lagged_sums", "tica_from_sums", "tica", "pca_from_sums", "pca", "htica_chunks",
    "htica", "project", "cv_normalization", "project_normalized", "kmeans_lloyd",
    "kmeans_assign", "find_centroids", "deeptica_loss", "deeptica_cov",
    "calinski_harabasz", "davies_bouldin",
]


# ----------------------------------------------------------------------------
# A2/A3/A4: statistics and standardisation
# ----------------------------------------------------------------------------
def column_stats(X: np.ndarray) -> dict:
    """Per-feature mean, std (ddof=1), min, max.

    Restates ``modules/cv_learning/cv_calculator.py:295-297``
    (``training_df.agg(['mean','std','min','max'])``).  Evaluated in float64
    and returned as float64; the reference stores float32 (pandas on float32
    columns), the golden ``features_norm_{mean,range}.npy`` pin ddof=1.
    """
    X64 = np.asarray(X, dtype=np.float64)
    n = X64.shape[0]
    mean = X64.mean(axis=0)
    if n > 1:
        std = np.sqrt(((X64 - mean) ** 2).sum(axis=0) / (n - 1))
    else:
        std = np.full(X64.shape[1], np.nan)
    return {"mean": mean, "std": std, "min": X64.min(axis=0), "max": X64.max(axis=0)}


def prepare_normalization(stats: dict, mode):
    """(mean, range) per normalisation mode; |range| < 1e-8 -> 1.0.

    Restates ``cv_calculator.py:308-363``.
    """
    if mode is None:
        means = np.zeros(len(stats["mean"]))
        ranges = np.ones(len(stats["mean"]))
    elif mode == "mean_std":
        means = np.array(stats["mean"], copy=True)
        ranges = np.array(stats["std"], copy=True)
    elif mode == "min_max_range1":
        means = np.array(stats["min"], copy=True)
        ranges = stats["max"] - stats["min"]
    elif mode == "min_max_range2":
        means = (stats["min"] + stats["max"]) / 2
        ranges = (stats["max"] - stats["min"]) / 2
    else:
        raise ValueError(f"Normalization mode {mode} not recognized.")
    ranges = np.array(ranges, copy=True)
    ranges[np.abs(ranges) < 1e-8] = 1.0
    return means, ranges


def standardize(X: np.ndarray, mean, rng) -> np.ndarray:
    """float32 ``(x - mean) / range`` with IEEE sub then div.

    Restates ``cv_calculator.py:806-837`` (``data.sub_(mean); data.div_(range)``
    on float32 tensors).  Returns a new float32 array.
    """
    X32 = np.asarray(X, dtype=np.float32)
    m32 = np.asarray(mean, dtype=np.float32)
    r32 = np.asarray(rng, dtype=np.float32)
    return ((X32 - m32) / r32).astype(np.float32)


# ----------------------------------------------------------------------------
# A5/A6: time-lagged pairs and TICA (mlcolvar restatement)
# ----------------------------------------------------------------------------
def lagged_pairs(Z: np.ndarray, lag: int):
    """``x_t = Z[:N-lag]``, ``x_lag = Z[lag:]`` (M = N - lag pairs, unit weights).

    Restates mlcolvar ``create_timelagged_dataset`` as called at
    ``cv_calculator.py:2247, 2309`` (pinned against golden TICA weights:
    N-lag pairs).
    """
    n = Z.shape[0]
    if lag < 1 or lag >= n:
        raise ValueError(f"lag {lag} out of range for {n} frames")
    return Z[: n - lag], Z[lag:]


def lagged_sums(Z: np.ndarray, lag: int):
    """Raw float64 sums  S0 = sum z_t z_t^T,  St = sum z_t z_{t+lag}^T,
    a = sum z_t, b = sum z_{t+lag}  over the M = N - lag pairs."""
    x_t, x_lag = lagged_pairs(np.asarray(Z, dtype=np.float64), lag)
    S0 = x_t.T @ x_t
    St = x_t.T @ x_lag
    return S0, St, x_t.sum(axis=0), x_lag.sum(axis=0), x_t.shape[0]


def _cholesky_eigh(C0, Ct, reg, out):
    """mlcolvar ``cholesky_eigh`` + TICA post-processing, float64.

    B = C0 + reg*I ; L = chol(B) ; A = L^-1 Ct L^-T ; eigh ; descending ;
    V = L^-T U ; unit-L2 columns ; sign so that row 0 >= 0 ; first ``out``.
    (Called from ``cv_calculator.py:2257-2261, 2350-2354, 2374-2378``.)
    """
    F = C0.shape[0]
    B = C0 + reg * np.eye(F)
    L = np.linalg.cholesky(B)
    Linv = np.linalg.inv(L)
    A = Linv @ Ct @ Linv.T
    A = 0.5 * (A + A.T)
    evals, U = np.linalg.eigh(A)
    order = np.argsort(evals)[::-1]
    evals = evals[order]
    U = U[:, order]
    V = Linv.T @ U
    V = V / np.linalg.norm(V, axis=0, keepdims=True)
    V = V * np.sign(V[0:1, :])
    out = min(out, F)
    return evals[:out], V[:, :out]


def tica_from_sums(S0, St, a, b, M, out, reg=1e-6):
    """TICA from the raw sums.

    mu = mean(x_t) is subtracted from BOTH series (mlcolvar ``TICA.compute``
    with ``remove_average=True``), covariances divided by M, C_tau symmetrised.
    """
    mu = a / M
    nu = b / M
    C0 = S0 / M - np.outer(mu, mu)
    C0 = 0.5 * (C0 + C0.T)
    # sum (x_t-mu)(x_lag-mu)^T = St - a mu^T - mu b^T + M mu mu^T = St - mu b^T  (a = M mu)
    Ct_raw = St / M - np.outer(mu, nu)
    Ct = 0.5 * (Ct_raw + Ct_raw.T)
    return _cholesky_eigh(C0, Ct, reg, out)


def tica(Z: np.ndarray, lag: int, out: int, reg: float = 1e-6):
    """Direct (two-pass, centred) float64 TICA on standardised data Z.

    Restates ``TICACalculator.compute_cv`` (``cv_calculator.py:2249-2267``) +
    mlcolvar ``TICA.compute(data=[x_t, x_lag], remove_average=True)``.
    """
    x_t, x_lag = lagged_pairs(np.asarray(Z, dtype=np.float64), lag)
    M = x_t.shape[0]
    mu = x_t.mean(axis=0)
    xt = x_t - mu
    xl = x_lag - mu
    C0 = xt.T @ xt / M
    C0 = 0.5 * (C0 + C0.T)
    Ct = 0.5 * (xt.T @ xl + xl.T @ xt) / M
    return _cholesky_eigh(C0, Ct, reg, out)


# ----------------------------------------------------------------------------
# A7: PCA (sklearn restatement)
# ----------------------------------------------------------------------------
def pca_from_sums(S_all, s_all, N, d):
    """PCA weights from full-data raw sums: C = (S - N mu mu^T)/(N-1), eigh,
    descending, sign rule ``W[0,i] >= 0`` (``cv_calculator.py:2204-2215``;
    sklearn ``_pca.py`` covariance_eigh solver)."""
    mu = s_all / N
    C = (S_all - N * np.outer(mu, mu)) / (N - 1)
    C = 0.5 * (C + C.T)
    evals, V = np.linalg.eigh(C)
    order = np.argsort(evals)[::-1]
    evals = evals[order][:d]
    W = V[:, order][:, :d].copy()
    for i in range(W.shape[1]):
        if W[0, i] < 0:
            W[:, i] = -W[:, i]
    return evals, W


def pca(Z: np.ndarray, d: int):
    """float64 PCA on standardised data (``PCACalculator.compute_cv``, ``:2194-2215``)."""
    Z64 = np.asarray(Z, dtype=np.float64)
    return pca_from_sums(Z64.T @ Z64, Z64.sum(axis=0), Z64.shape[0], d)


# ----------------------------------------------------------------------------
# A8: hierarchical TICA
# ----------------------------------------------------------------------------
def htica_chunks(F: int, num_subspaces: int):
    """Column chunks as ``torch.split(data, F // num_subspaces, dim=1)`` makes them
    (``cv_calculator.py:2331-2334``): chunk SIZE is F//ns, so there may be one more,
    narrower chunk.  Returns list of (start, stop); empty if F < num_subspaces."""
    w = F // num_subspaces
    if w == 0:
        return []
    return [(s, min(s + w, F)) for s in range(0, F, w)]


def htica(Z: np.ndarray, lag: int, num_subspaces: int, sub_dim: int, d: int, reg: float = 1e-6):
    """float64 hTICA (``HTICACalculator.compute_cv``, ``cv_calculator.py:2311-2384``).

    Level 1: TICA per column chunk, ``out = sub_dim``; project the UNCENTRED
    standardised data (``:2363-2364``); level 2: TICA on the concatenation;
    ``W = blockdiag(V_b) @ V2`` (``:2367, 2384``).
    Returns (W, T1, V2).
    """
    Z64 = np.asarray(Z, dtype=np.float64)
    F = Z64.shape[1]
    chunks = htica_chunks(F, num_subspaces)
    if not chunks:
        raise ValueError("num_subspaces larger than number of features")
    x_t, x_lag = lagged_pairs(Z64, lag)
    blocks, proj_t, proj_l = [], [], []
    for (s, e) in chunks:
        _, Vb = tica(Z64[:, s:e], lag, sub_dim, reg)
        blocks.append(Vb)
        proj_t.append(x_t[:, s:e] @ Vb)
        proj_l.append(x_lag[:, s:e] @ Vb)
    S1 = sum(b.shape[1] for b in blocks)
    T1 = np.zeros((F, S1))
    c = 0
    for (s, e), Vb in zip(chunks, blocks):
        T1[s:e, c:c + Vb.shape[1]] = Vb
        c += Vb.shape[1]
    P_t = np.concatenate(proj_t, axis=1)
    P_l = np.concatenate(proj_l, axis=1)
    # level-2 TICA on the pairs (P_t, P_l)
    M = P_t.shape[0]
    mu = P_t.mean(axis=0)
    xt = P_t - mu
    xl = P_l - mu
    C0 = xt.T @ xt / M
    C0 = 0.5 * (C0 + C0.T)
    Ct = 0.5 * (xt.T @ xl + xl.T @ xt) / M
    _, V2 = _cholesky_eigh(C0, Ct, reg, d)
    return T1 @ V2, T1, V2


# ----------------------------------------------------------------------------
# A9/A10: projection and CV normalisation
# ----------------------------------------------------------------------------
def project(Z: np.ndarray, W: np.ndarray, dtype=np.float64) -> np.ndarray:
    """``P = Z @ W`` (``cv_calculator.py:958, 981``)."""
    return np.asarray(Z, dtype=dtype) @ np.asarray(W, dtype=dtype)


def cv_normalization(P: np.ndarray):
    """``cv_norm_mean = (max+min)/2``, ``cv_norm_range = (max-min)/2``
    (``cv_calculator.py:984-991``)."""
    mn = P.min(axis=0)
    mx = P.max(axis=0)
    return (mx + mn) / 2, (mx - mn) / 2


def project_normalized(X_raw, f_mean, f_range, W, cv_mean, cv_range) -> np.ndarray:
    """Full ``project_data(normalize_data=True)`` in float32, as the reference runs it
    (``cv_calculator.py:946-970``): standardise, ``@ W``, ``(P - cv_mean)/cv_range``."""
    Z = standardize(X_raw, f_mean, f_range)
    P = (Z @ np.asarray(W, dtype=np.float32)).astype(np.float32)
    return ((P - np.asarray(cv_mean, np.float32)) / np.asarray(cv_range, np.float32)).astype(np.float32)


# ----------------------------------------------------------------------------
# K1/K3: KMeans (sklearn Lloyd restatement) and find_centroids
# ----------------------------------------------------------------------------
def kmeans_assign(X, centers):
    """E-step: argmin_j ||c_j||^2 - 2 x.c_j with strict ``<`` (lowest index wins ties)
    (``sklearn/cluster/_k_means_lloyd.pyx:193-214``).  Returns (labels, best, second)
    where best/second are the two smallest full squared distances (for tie reporting)."""
    X = np.asarray(X, dtype=np.float64)
    C = np.asarray(centers, dtype=np.float64)
    n = X.shape[0]
    labels = np.empty(n, dtype=np.int32)
    best = np.empty(n)
    second = np.empty(n)
    xsq = (X * X).sum(axis=1)
    csq = (C * C).sum(axis=1)
    step = max(1, (1 << 22) // max(1, C.shape[0]))
    for s in range(0, n, step):
        D = csq[None, :] - 2.0 * (X[s:s + step] @ C.T)
        lab = D.argmin(axis=1)
        labels[s:s + step] = lab
        rows = np.arange(D.shape[0])
        b = D[rows, lab].copy()
        if C.shape[0] > 1:
            D[rows, lab] = np.inf
            sec = D.min(axis=1)
        else:
            sec = np.full_like(b, np.inf)
        best[s:s + step] = b + xsq[s:s + step]
        second[s:s + step] = sec + xsq[s:s + step]
    return labels, best, second


def kmeans_lloyd(X, init_centers, max_iter: int = 300, tol: float = 1e-4):
    """sklearn ``KMeans(init=ndarray, n_init=1, algorithm='lloyd')`` control flow in float64.

    Restates what ``statistics.kmeans_clustering`` (``modules/statistics/statistics.py:159-197``)
    triggers inside scikit-learn (``sklearn/cluster/_kmeans.py``): centre X by its column
    mean (``:1486-1493``); ``tol_eff = tol * mean(var(X, axis=0))`` (``:285-294``); Lloyd loop
    with label-equality ("strict") convergence first, then ``sum ||dc||^2 <= tol_eff``
    (``:703-740``); a final E-step when not strictly converged (``:742-754``); empty-cluster
    relocation to the farthest points (``_k_means_common.pyx:167-211``); centres = sums *
    (1/count); centres returned un-centred.
    Returns dict(labels, centers, n_iter, strict, best, second).
    """
    X = np.array(X, dtype=np.float64, copy=True)
    C = np.array(init_centers, dtype=np.float64, copy=True)
    n, d = X.shape
    k = C.shape[0]
    tol_eff = float(np.mean(np.var(X, axis=0)) * tol)
    x_mean = X.mean(axis=0)
    X -= x_mean
    C -= x_mean
    labels_old = np.full(n, -1, dtype=np.int32)
    labels = labels_old.copy()
    strict = False
    n_iter = 0
    best = second = None
    for it in range(max_iter):
        labels, best, second = kmeans_assign(X, C)
        sums = np.zeros((k, d))
        np.add.at(sums, labels, X)
        counts = np.bincount(labels, minlength=k).astype(np.float64)
        empty = np.where(counts == 0)[0]
        if empty.size:
            dist = ((X - C[labels]) ** 2).sum(axis=1)
            far = np.argpartition(dist, -empty.size)[: -empty.size - 1: -1]
            for idx, new_id in enumerate(empty):
                fi = far[idx]
                old_id = labels[fi]
                sums[old_id] -= X[fi]
                sums[new_id] = X[fi]
                counts[new_id] = 1.0
                counts[old_id] -= 1.0
        C_new = C.copy()
        nz = counts > 0
        C_new[nz] = sums[nz] * (1.0 / counts[nz])[:, None]
        shift_tot = float(((C_new - C) ** 2).sum())
        C = C_new
        n_iter = it + 1
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift_tot <= tol_eff:
            break
        labels_old = labels.copy()
    if not strict:
        labels, best, second = kmeans_assign(X, C)
    return {"labels": labels, "centers": C + x_mean, "n_iter": n_iter, "strict": strict,
            "best": best, "second": second}


def find_centroids(X, centers) -> np.ndarray:
    """Per centre: index of the first arg-min sample by Euclidean distance
    (``modules/statistics/statistics.py:370-377``).  Returns int64 array (k,)."""
    X = np.asarray(X, dtype=np.float64)
    out = np.empty(len(centers), dtype=np.int64)
    for i, c in enumerate(np.asarray(centers, dtype=np.float64)):
        out[i] = int(np.argmin(np.linalg.norm(X - c, axis=1)))
    return out


# ----------------------------------------------------------------------------
# N1: cluster-validity scores that need one pass given labels + centres
# ----------------------------------------------------------------------------
def calinski_harabasz(X, labels) -> float:
    """sklearn ``calinski_harabasz_score`` (called at ``statistics.py:73``)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    ks = np.unique(labels)
    mean = X.mean(axis=0)
    extra = intra = 0.0
    for c in ks:
        Xc = X[labels == c]
        mc = Xc.mean(axis=0)
        extra += len(Xc) * ((mc - mean) ** 2).sum()
        intra += ((Xc - mc) ** 2).sum()
    k = len(ks)
    return 1.0 if intra == 0.0 else float(extra * (n - k) / (intra * (k - 1.0)))


def davies_bouldin(X, labels) -> float:
    """sklearn ``davies_bouldin_score`` (called at ``statistics.py:74``)."""
    X = np.asarray(X, dtype=np.float64)
    ks = np.unique(labels)
    k = len(ks)
    cent = np.zeros((k, X.shape[1]))
    intra = np.zeros(k)
    for i, c in enumerate(ks):
        Xc = X[labels == c]
        cent[i] = Xc.mean(axis=0)
        intra[i] = np.mean(np.linalg.norm(Xc - cent[i], axis=1))
    D = np.linalg.norm(cent[:, None, :] - cent[None, :, :], axis=2)
    if np.allclose(intra, 0) or np.allclose(D, 0):
        return 0.0
    D[D == 0] = np.inf
    comb = intra[:, None] + intra[None, :]
    return float(np.mean(np.max(comb / D, axis=1)))


# ----------------------------------------------------------------------------
# A11: DeepTICA minibatch covariance + loss (mlcolvar restatement)
# ----------------------------------------------------------------------------
def deeptica_cov(f, g, w=None, wl=None):
    """Weighted mean-free C0 / C_tau of network outputs.

    mu = sum w_n f_n (weights normalised to 1); C0 = sum w f~ f~^T (symmetrised);
    Ct = 1/2 sum w' (f~ g~^T + g~ f~^T).  mlcolvar ``TICA.compute`` with
    ``remove_average=True`` as driven from ``cv_calculator.py:1508-1524``.
    """
    f = np.asarray(f, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    B = f.shape[0]
    w = np.ones(B) if w is None else np.asarray(w, dtype=np.float64)
    wl = np.ones(B) if wl is None else np.asarray(wl, dtype=np.float64)
    wn = w / w.sum()
    wln = wl / wl.sum()
    mu = (wn[:, None] * f).sum(axis=0)
    ft = f - mu
    gt = g - mu
    C0 = (ft * wn[:, None]).T @ ft
    C0 = 0.5 * (C0 + C0.T)
    Ct = (ft * wln[:, None]).T @ gt
    Ct = 0.5 * (Ct + Ct.T)
    return C0, Ct, mu


def deeptica_loss(f, g, w=None, wl=None, reg: float = 1e-6, n_eig: int = 0):
    """DeepTICA loss  -sum lambda_i^2  (mlcolvar ``ReduceEigenvaluesLoss(mode='sum2')``).
    Returns (loss, evals descending)."""
    C0, Ct, _ = deeptica_cov(f, g, w, wl)
    evals, _ = _cholesky_eigh(C0, Ct, reg, C0.shape[0])
    if n_eig and n_eig > 0:
        evals = evals[:n_eig]
    return float(-(evals ** 2).sum()), evals
