"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the free-energy-surface estimate the reference
draws from the projected trajectory.

Reference path: ``modules/figures/figures.py:24-110`` calls
``mlcolvar.utils.fes.compute_fes(data, temp, backend="KDEpy", num_samples=num_bins, bandwidth,
blocks=num_blocks, eps=1e-10, bounds=get_ranges(data))`` (:95), driven from
``tools/train_colvars/train_colvars_workflow.py:146-182`` (1-D per CV component with 100 blocks, 2-D per
CV pair with 1 block).  mlcolvar 1.2.2 and KDEpy are absent from this image; the published algorithm is
restated here:

  * evaluation grid: ``num_samples`` equidistant nodes per axis spanning ``bounds``; 2-D grids are
    ``numpy.meshgrid`` ('xy': fes[iy, ix]);
  * KDEpy ``FFTKDE(bw=bandwidth, kernel='gaussian')``: linear binning of the samples onto the grid,
    then a convolution with the Gaussian of standard deviation ``bandwidth`` sampled at the grid
    offsets (here over its full support; KDEpy truncates it where it is negligible);
  * ``fes = -kbT log(density + eps)``, ``kbT = 0.00831441 * temp`` kJ/mol, shifted so that min = 0;
  * ``blocks`` > 1: the frames are split with ``numpy.array_split``, every block gets its own FES,
    and the result is the weighted block average with the standard error of the weighted mean
    (effective number of blocks ``(sum W)^2 / sum W^2``).

PINNED (single block, 2-D): the reference's own legacy output
``data/calpha_transitions/reference/1rcs_B-3ssx_R-3/train_colvars/pca/fes/{fes,grid,bounds}.npy``
(bandwidth 0.025, 200 bins, 300 K: ``data/calpha_transitions/input/distances_config.yml:93-95``) is
reproduced from its ``projected_trajectory.csv`` to 1e-3 kJ/mol wherever fes < 20
(``tests/golden/fes_legacy_pca.npz``, ``tests/test_oracle_golden.py``).
UNPINNED: the block average / error for blocks > 1 (no reference-held output has an error array).
"""
from __future__ import annotations

import numpy as np

KB = 0.00831441          # kJ / (mol K), as mlcolvar


def get_ranges(X: np.ndarray):
    """Range of the data along each dimension +/- 0.5 % (reference figures.py:399-470, without the
    supplementary data)."""
    X = np.asarray(X)
    if X.ndim == 1:
        lo, hi = float(X.min()), float(X.max())
        off = 0.005 * (hi - lo)
        return (lo - off, hi + off)
    out = []
    for i in range(X.shape[1]):
        lo, hi = float(X[:, i].min()), float(X[:, i].max())
        off = 0.005 * (hi - lo)
        out.append((lo - off, hi + off))
    return out


def linear_binning(X: np.ndarray, bounds, G: int) -> np.ndarray:
    """KDEpy linear binning: weights on the grid nodes around every sample; sums to len(X).
    1-D -> (G,), 2-D -> (G, G) indexed [iy, ix]."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        lo, hi = bounds
        f = (X - lo) / (hi - lo) * (G - 1)
        i = np.minimum(np.floor(f).astype(np.int64), G - 2)
        a = f - i
        w = np.zeros(G)
        np.add.at(w, i, 1 - a)
        np.add.at(w, i + 1, a)
        return w
    (lo0, hi0), (lo1, hi1) = bounds
    fx = (X[:, 0] - lo0) / (hi0 - lo0) * (G - 1)
    fy = (X[:, 1] - lo1) / (hi1 - lo1) * (G - 1)
    ix = np.minimum(np.floor(fx).astype(np.int64), G - 2)
    iy = np.minimum(np.floor(fy).astype(np.int64), G - 2)
    ax, ay = fx - ix, fy - iy
    w = np.zeros((G, G))
    np.add.at(w, (iy, ix), (1 - ax) * (1 - ay))
    np.add.at(w, (iy, ix + 1), ax * (1 - ay))
    np.add.at(w, (iy + 1, ix), (1 - ax) * ay)
    np.add.at(w, (iy + 1, ix + 1), ax * ay)
    return w


def _gauss_matrix(G: int, step: float, h: float) -> np.ndarray:
    d = (np.arange(G)[:, None] - np.arange(G)[None, :]) * step
    return np.exp(-0.5 * (d / h) ** 2) / (np.sqrt(2 * np.pi) * h)


def binned_density(X, bounds, G: int, bandwidth: float) -> np.ndarray:
    X = np.asarray(X)
    w = linear_binning(X, bounds, G) / X.shape[0]
    if X.ndim == 1:
        step = (bounds[1] - bounds[0]) / (G - 1)
        return _gauss_matrix(G, step, bandwidth) @ w
    sx = (bounds[0][1] - bounds[0][0]) / (G - 1)
    sy = (bounds[1][1] - bounds[1][0]) / (G - 1)
    return _gauss_matrix(G, sy, bandwidth) @ w @ _gauss_matrix(G, sx, bandwidth).T


def compute_fes(X, temp: float = 300.0, num_samples: int = 100, bounds=None, bandwidth: float = 0.01,
                blocks: int = 1, eps: float = 0.0):
    """Returns (fes, grid, bounds, error) like mlcolvar's compute_fes (error None for one block)."""
    X = np.asarray(X, dtype=np.float64)
    dim = 1 if X.ndim == 1 else X.shape[1]
    if dim == 2 and X.ndim == 2 and X.shape[1] == 1:
        X, dim = X[:, 0], 1
    kbt = KB * temp
    if bounds is None:
        off = 1e-3
        bounds = (X.min() - off, X.max() + off) if dim == 1 else [(X[:, i].min() - off, X[:, i].max() + off) for i in range(dim)]
    G = int(num_samples)
    if dim == 1:
        grid = np.linspace(bounds[0], bounds[1], G)
    else:
        grid = np.meshgrid(*[np.linspace(b[0], b[1], G) for b in bounds])
    O, W = [], []
    for Xb in np.array_split(X, blocks):
        f = -kbt * np.log(binned_density(Xb, bounds, G, bandwidth) + eps)
        O.append(f - f.min())
        W.append(float(Xb.shape[0]))
    if blocks == 1:
        return O[0], grid, bounds, None
    O, W = np.asarray(O), np.asarray(W)
    Wb = W.reshape((-1,) + (1,) * (O.ndim - 1))
    fes = np.nansum(O * Wb, axis=0) / np.nansum(W)
    dev = O - fes
    blocks_eff = W.sum() ** 2 / (W ** 2).sum()
    variance = blocks_eff / (blocks_eff - 1) * np.nansum(dev ** 2 * Wb, axis=0) / np.nansum(W)
    error = np.sqrt(variance / blocks_eff)
    return fes, grid, bounds, error
