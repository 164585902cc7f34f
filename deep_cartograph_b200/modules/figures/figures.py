"""Free-energy surfaces of the projected trajectory on the B200 kernels.

Host-side mirror of the FES part of the reference's ``modules/figures/figures.py``: ``plot_fes``
(:24-192, the computation and the ``fes*.npy`` files; matplotlib / seaborn are not in this image, so
nothing is drawn) and ``get_ranges`` (:399-470).  ``compute_fes`` has the meaning of
``mlcolvar.utils.fes.compute_fes(..., backend="KDEpy")`` as the reference calls it (:95): linear binning
onto ``num_samples`` equidistant nodes per axis + Gaussian kernel of standard deviation ``bandwidth``,
``fes = -kbT log(density + eps)`` shifted to min 0, weighted block average and its standard error for
``blocks`` > 1.  The pass over the frames and the convolution run in ``dcg_fes_bin_f32`` /
``dcg_fes_smooth_f64``; the O(blocks x grid) remainder is torch FP64 on the device.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from ... import ops

logger = logging.getLogger(__name__)

KB = 0.00831441      # kJ / (mol K), as mlcolvar


def get_ranges(X, X_ref: Optional[List[np.ndarray]] = None) -> List:
    """Range of the data along each dimension +/- 0.5 % of it (reference figures.py:399-470)."""
    X = X.detach().cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
    cols = [X] if X.ndim == 1 else [X[:, i] for i in range(X.shape[1])]
    out = []
    for i, c in enumerate(cols):
        lo, hi = float(np.min(c)), float(np.max(c))
        for r in X_ref or []:
            r = np.asarray(r)
            ri = r if r.ndim == 1 else r[:, i]
            lo, hi = min(lo, float(np.min(ri))), max(hi, float(np.max(ri)))
        off = 0.005 * (hi - lo)
        out.append((lo - off, hi + off))
    return out[0] if X.ndim == 1 else out


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("deep_cartograph_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def compute_fes(X, temp: float = 300.0, num_samples: int = 100, bounds=None, bandwidth: float = 0.01,
                blocks: int = 1, eps: float = 0.0, cols=None):
    """(fes, grid, bounds, error) of a 1-D or 2-D data set; numpy arrays like mlcolvar's.  ``X`` may be
    a CUDA tensor (n, d) with ``cols`` selecting one or two columns, so the projection of the training
    data is used where it lies."""
    if isinstance(X, torch.Tensor) and X.is_cuda:
        P = X if X.dim() == 2 else X.reshape(-1, 1)
    else:
        A = np.asarray(X, dtype=np.float32)
        P = torch.from_numpy(np.ascontiguousarray(A if A.ndim == 2 else A.reshape(-1, 1))).to(_device())
    if P.dtype != torch.float32:
        P = P.to(torch.float32)
    cols = list(range(P.shape[1])) if cols is None else list(cols)
    dim = len(cols)
    if dim not in (1, 2):
        raise ValueError("FES is computed along one or two variables")
    if bounds is None:
        off = 1e-3        # mlcolvar's default offset when no bounds are given
        mn, mx = P[:, cols].amin(dim=0).tolist(), P[:, cols].amax(dim=0).tolist()
        bounds = [(a - off, b + off) for a, b in zip(mn, mx)]
        bounds = bounds[0] if dim == 1 else bounds
    bl = [tuple(bounds)] if dim == 1 else [tuple(b) for b in bounds]
    G = int(num_samples)
    kbt = KB * temp
    dens, frames, outside = ops.fes_density(P, cols, bounds, G, bandwidth, blocks)
    fes_b = -kbt * torch.log(dens + eps)
    fes_b = fes_b - fes_b.reshape(blocks, -1).min(dim=1).values.reshape((blocks,) + (1,) * dim)
    n_out = int(outside.item())
    if n_out:
        logger.warning(f"{n_out} frames lie outside the FES bounds and were left out")
    if dim == 1:
        grid = np.linspace(bl[0][0], bl[0][1], G)
    else:
        grid = np.meshgrid(*[np.linspace(b[0], b[1], G) for b in bl])
    if blocks == 1:
        return fes_b[0].cpu().numpy(), grid, bounds, None
    W = frames.reshape((blocks,) + (1,) * dim)
    fes = (fes_b * W).nansum(dim=0) / frames.sum()
    dev2 = (fes_b - fes) ** 2
    blocks_eff = frames.sum() ** 2 / (frames ** 2).sum()
    variance = blocks_eff / (blocks_eff - 1) * (dev2 * W).nansum(dim=0) / frames.sum()
    error = torch.sqrt(variance / blocks_eff)
    return fes.cpu().numpy(), grid, bounds, error.cpu().numpy()


def plot_fes(data, cv_labels: List[str], settings: Dict, output_path: str, num_blocks: int = 1,
             sup_data=None, sup_data_labels=None, legend_cutoff: int = 10, cols=None):
    """The computation and the files of the reference's ``plot_fes`` (figures.py:24-110): FES along the
    given variables with ``settings`` (compute, save, temperature, bandwidth, num_bins), block count
    reduced when a block would hold fewer than 100 samples (:81-87); ``fes.npy``, ``fes_grid.npy``,
    ``fes_bounds.npy``, ``fes_error.npy`` written to ``output_path`` when ``save``.  No figure is drawn."""
    s = {"compute": True, "save": True, "temperature": 300, "bandwidth": 0.05, "num_fes_levels": 10,
         "num_bins": 150, "max_fes": 30}
    s.update(settings or {})
    if not s["compute"]:
        return None
    logger.info("Computing FES(" + ", ".join(cv_labels) + ")...")
    n = data.shape[0]
    min_block_size = 100
    if int(n / num_blocks) < min_block_size:
        old = num_blocks
        num_blocks = max(1, int(n / min_block_size))
        logger.warning(f"Block size too small with {old} blocks and {n} samples. Reducing the number of blocks to {num_blocks}")
    if isinstance(data, torch.Tensor):
        sel = data if cols is None else data[:, list(cols)]
        bounds = get_ranges(sel if sel.shape[1] > 1 else sel[:, 0])
    else:
        bounds = get_ranges(np.asarray(data))
    fes, grid, bounds, err = compute_fes(data, temp=s["temperature"], num_samples=s["num_bins"], bounds=bounds,
                                         bandwidth=s["bandwidth"], blocks=num_blocks, eps=1e-10, cols=cols)
    if s.get("save", False):
        os.makedirs(output_path, exist_ok=True)
        np.save(os.path.join(output_path, "fes.npy"), fes)
        np.save(os.path.join(output_path, "fes_grid.npy"), np.asarray(grid))
        np.save(os.path.join(output_path, "fes_bounds.npy"), np.asarray(bounds))
        np.save(os.path.join(output_path, "fes_error.npy"), np.asarray(err, dtype=object) if err is None else err)
    return fes, grid, bounds, err
