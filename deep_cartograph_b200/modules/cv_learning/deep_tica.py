"""DeepTICA minibatch loss on the fused correlation kernel.

Replaces the covariance part of mlcolvar's ``DeepTICA.training_step`` as driven from the
reference's ``NonLinear.train`` (cv_calculator.py:1456-1553, regularisation at :1508-1509):
weighted mean, mean-free C0 / C_tau of the network outputs, Cholesky-reduced eigenvalues,
``loss = -sum lambda_i^2`` (mlcolvar ``ReduceEigenvaluesLoss(mode='sum2')``).

The B x d reduction runs in one CUDA kernel (``dcg_ticacov_f32``, FP64 sums); its backward is
analytic; the d x d Cholesky / eigh stay in ``torch.linalg`` (autograd supplies dL/dC0, dL/dCt).

Frame-sharded training ("exact all-reduce"): C0 / C_tau are statistics of
the WHOLE minibatch, so with the minibatch split over ranks the raw sums are all-reduced before
the eigen step (``shards``: 2 + 3d + 2d^2 doubles) and every rank evaluates the same loss; the
backward needs one more all-reduce of d doubles (the gradient through the batch mean), and the
parameter gradients are SUMMED over ranks (``allreduce_gradients_``) -- the result equals the
single-process gradient of the loss on the concatenated minibatch.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from ... import ops


class _TicaCov(torch.autograd.Function):
    """(f, g, w, wl) -> (C0, Ct) in float64, mean-free with mu = sum_n w_n f_n / sum w."""

    @staticmethod
    def forward(ctx, f, g, w, wl, shards=None):
        s = ops.ticacov_sums(f.detach(), g.detach(), w, wl)
        if shards is not None:
            shards.allreduce_sum_(s["flat"])          # the views below see the global sums
        ctx.shards = shards
        sw, swl = s["sw"], s["swl"]
        mu = s["swf"] / sw
        C0 = s["sff"] / sw - torch.outer(mu, mu)
        C0 = 0.5 * (C0 + C0.T)
        # sum wl (f-mu)(g-mu)^T / swl
        Ct = s["sfg"] / swl - torch.outer(mu, s["slg"] / swl) - torch.outer(s["slf"] / swl, mu) + torch.outer(mu, mu)
        Ct = 0.5 * (Ct + Ct.T)
        ctx.save_for_backward(f, g, w if w is not None else torch.empty(0, device=f.device),
                              wl if wl is not None else torch.empty(0, device=f.device), mu, sw, swl)
        return C0, Ct

    @staticmethod
    def backward(ctx, G0, Gt):
        f, g, w, wl, mu, sw, swl = ctx.saved_tensors
        dt = f.dtype
        B = f.shape[0]
        wn = (w.to(torch.float64) / sw) if w.numel() else torch.full((B,), 1.0 / float(sw), dtype=torch.float64, device=f.device)
        wln = (wl.to(torch.float64) / swl) if wl.numel() else torch.full((B,), 1.0 / float(swl), dtype=torch.float64, device=f.device)
        ft = f.to(torch.float64) - mu
        gt = g.to(torch.float64) - mu
        H0 = G0 + G0.T                      # d/dA of sym(A), times the factor 2 of the quadratic form
        Ht = 0.5 * (Gt + Gt.T)
        d_ft = wn[:, None] * (ft @ H0) + wln[:, None] * (gt @ Ht)
        d_gt = wln[:, None] * (ft @ Ht)
        # mean subtraction: f~ = f - mu, g~ = g - mu, mu = sum wn f
        tot = d_ft.sum(dim=0) + d_gt.sum(dim=0)
        if ctx.shards is not None:
            ctx.shards.allreduce_sum_(tot)            # mu is the mean of the WHOLE minibatch
        d_f = d_ft - wn[:, None] * tot
        return d_f.to(dt), d_gt.to(dt), None, None, None


def tica_covariances(f: torch.Tensor, g: torch.Tensor, w: Optional[torch.Tensor] = None,
                     wl: Optional[torch.Tensor] = None, shards=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Mean-free, symmetrised C0 and C_tau (float64, d x d) of the outputs f = nn(x_t),
    g = nn(x_{t+lag}) (float32, B x d), differentiable w.r.t. f and g.  With ``shards``
    (parallel.FrameShards) f and g are this rank's part of the minibatch and C0 / C_tau are those
    of the whole minibatch (identical on every rank)."""
    return _TicaCov.apply(f.contiguous(), g.contiguous(), w, wl, shards)


def allreduce_gradients_(module: nn.Module, shards) -> None:
    """SUM the parameter gradients over ranks in one fused collective (the loss is a function of
    the global minibatch: dL/dtheta = sum over ranks of the local chain-rule terms)."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads or shards is None:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    shards.allreduce_sum_(flat)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def reduced_eigenvalues(C0: torch.Tensor, Ct: torch.Tensor, reg: float) -> torch.Tensor:
    """Eigenvalues of L^-1 Ct L^-T with L = chol(C0 + reg I), descending (mlcolvar cholesky_eigh)."""
    d = C0.shape[0]
    L = torch.linalg.cholesky(C0 + reg * torch.eye(d, dtype=C0.dtype, device=C0.device))
    Y = torch.linalg.solve_triangular(L, Ct, upper=False)
    A = torch.linalg.solve_triangular(L, Y.T, upper=False).T
    A = 0.5 * (A + A.T)
    return torch.linalg.eigvalsh(A).flip(0)


class _TicaLoss(torch.autograd.Function):
    """(f, g, w, wl) -> (loss, eigenvalues) with the whole d x d part -- mean-free C0 / C_tau,
    Cholesky reduction, eigenvalues, ``-sum lambda^2`` and dLoss/dC0, dLoss/dC_tau -- in ONE kernel
    (``dcg_ticaloss_f64``) instead of ~40 tiny torch.linalg launches and their autograd backward.
    The backward with respect to f and g is the analytic one of ``_TicaCov``."""

    @staticmethod
    def forward(ctx, f, g, w, wl, reg, n_eig, shards):
        d = f.shape[1]
        s = ops.ticacov_sums(f.detach(), g.detach(), w, wl)
        if shards is not None:
            shards.allreduce_sum_(s["flat"])
        r = ops.ticaloss(s["flat"], d, reg, n_eig)
        ctx.shards = shards
        ctx.save_for_backward(f, g, w if w is not None else torch.empty(0, device=f.device),
                              wl if wl is not None else torch.empty(0, device=f.device),
                              r["mu"], r["sw"], r["swl"], r["G0"], r["Gt"])
        ctx.mark_non_differentiable(r["evals"])
        return r["loss"].clone(), r["evals"]

    @staticmethod
    def backward(ctx, g_loss, _g_evals):
        f, g, w, wl, mu, sw, swl, G0, Gt = ctx.saved_tensors
        dt = f.dtype
        B = f.shape[0]
        wn = (w.to(torch.float64) / sw) if w.numel() else (torch.ones((), dtype=torch.float64, device=f.device) / sw).expand(B)
        wln = (wl.to(torch.float64) / swl) if wl.numel() else (torch.ones((), dtype=torch.float64, device=f.device) / swl).expand(B)
        ft = f.to(torch.float64) - mu
        gt = g.to(torch.float64) - mu
        # a batch whose C0 + reg I was not positive definite (NaN loss) contributes no gradient
        H0 = torch.nan_to_num((G0 + G0.T) * g_loss, nan=0.0, posinf=0.0, neginf=0.0)
        Ht = torch.nan_to_num((0.5 * g_loss) * (Gt + Gt.T), nan=0.0, posinf=0.0, neginf=0.0)
        d_ft = wn[:, None] * (ft @ H0) + wln[:, None] * (gt @ Ht)
        d_gt = wln[:, None] * (ft @ Ht)
        tot = d_ft.sum(dim=0) + d_gt.sum(dim=0)
        if ctx.shards is not None:
            ctx.shards.allreduce_sum_(tot)
        d_f = d_ft - wn[:, None] * tot
        return d_f.to(dt), d_gt.to(dt), None, None, None, None, None


def tica_loss(f: torch.Tensor, g: torch.Tensor, w: Optional[torch.Tensor] = None,
              wl: Optional[torch.Tensor] = None, reg: float = 1e-6, n_eig: int = 0, shards=None):
    """DeepTICA loss ``-sum lambda_i^2`` and the eigenvalues (descending).  A batch whose
    ``C0 + reg I`` is not positive definite gives a NaN loss (the reference's Cholesky raises)."""
    return _TicaLoss.apply(f.contiguous(), g.contiguous(), w, wl, float(reg), int(n_eig or 0), shards)


def tica_loss_reference(f: torch.Tensor, g: torch.Tensor, w: Optional[torch.Tensor] = None,
                        wl: Optional[torch.Tensor] = None, reg: float = 1e-6, n_eig: int = 0, shards=None):
    """The same loss through ``tica_covariances`` + torch.linalg autograd (validation of the fused
    kernel; ~40 launches)."""
    C0, Ct = tica_covariances(f, g, w, wl, shards)
    evals = reduced_eigenvalues(C0, Ct, reg)
    used = evals[:n_eig] if n_eig and n_eig > 0 else evals
    return -(used ** 2).sum(), evals


class DeepTICA(nn.Module):
    """Feed-forward DeepTICA CV with the graph of the reference's exported model:
    ``norm_in: (x - mean)/range -> nn -> tica: (y - mu) @ evecs -> postprocessing: (z - m)/r``."""

    def __init__(self, layers: List[int], mean: torch.Tensor, rng: torch.Tensor,
                 activation: str = "tanh", reg: float = 1e-6):
        super().__init__()
        acts = {"tanh": nn.Tanh, "relu": nn.ReLU, "elu": nn.ELU, "leaky_relu": nn.LeakyReLU,
                "softplus": nn.Softplus}
        mods: List[nn.Module] = []
        for i in range(len(layers) - 1):
            mods.append(nn.Linear(layers[i], layers[i + 1]))
            if i < len(layers) - 2:
                mods.append(acts[activation]())
        self.nn = nn.Sequential(*mods)
        d = layers[-1]
        self.reg = reg
        self.register_buffer("in_mean", mean.clone().to(torch.float32))
        self.register_buffer("in_range", rng.clone().to(torch.float32))
        self.register_buffer("tica_mean", torch.zeros(d))
        self.register_buffer("tica_evecs", torch.eye(d))
        self.register_buffer("out_mean", torch.zeros(d))
        self.register_buffer("out_range", torch.ones(d))

    def features(self, x: torch.Tensor) -> torch.Tensor:
        return self.nn((x - self.in_mean) / self.in_range)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        z = (self.features(x) - self.tica_mean) @ self.tica_evecs
        return (z - self.out_mean) / self.out_range

    def loss(self, x_t: torch.Tensor, x_lag: torch.Tensor, w=None, wl=None, shards=None):
        return tica_loss(self.features(x_t), self.features(x_lag), w, wl, reg=self.reg, shards=shards)

    def loss_indexed(self, X: torch.Tensor, idx: torch.Tensor, lag: int, w=None, wl=None, shards=None):
        """``loss(X[idx], X[idx + lag])`` with the two minibatches gathered and passed through
        ``norm_in`` by one kernel each (same float32 values as ``features``)."""
        z_t = ops.gather_standardize(X, idx, self.in_mean, self.in_range, 0)
        z_l = ops.gather_standardize(X, idx, self.in_mean, self.in_range, lag)
        return tica_loss(self.nn(z_t), self.nn(z_l), w, wl, reg=self.reg, shards=shards)

    @torch.no_grad()
    def fit_tica_layer(self, x_t: torch.Tensor, x_lag: torch.Tensor):
        """Freeze the linear TICA read-out from the current network outputs (unit-L2 columns,
        sign of row 0, as mlcolvar's TICA.compute(save_params=True))."""
        from ... import linalg
        f, g = self.features(x_t), self.features(x_lag)
        s = ops.ticacov_sums(f, g)
        M = f.shape[0]
        evals, V = linalg.tica_from_sums(s["sff"], s["sfg"], s["swf"], s["slg"], M, f.shape[1], self.reg)
        self.tica_mean.copy_((s["swf"] / M).to(torch.float32))
        self.tica_evecs.copy_(V.to(torch.float32))
        return evals
