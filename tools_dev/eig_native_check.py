"""Hand-written eigen-stage kernels against torch.linalg: Cholesky + inverse, the iteration, end result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.synthetic import feature_matrix
dev = torch.device("cuda:0")
for F in (1000, 495, 50, 333):
    n, lag, out = 100000, 10, 4
    X = feature_matrix(n, F, 0, n, dev)
    st = ops.column_stats(X)
    mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
    s = ops.lagged_covariance(X, lag, mean, rng, xmin=st["min"], xmax=st["max"])
    S0 = ops.symmetrize_upper(s["S0"]); M = s["M"]
    mu = s["a"] / M; nu = s["b"] / M
    C0 = S0 / M - torch.outer(mu, mu); C0 = 0.5 * (C0 + C0.T)
    Ct = s["St"] / M - torch.outer(mu, nu); Ct = 0.5 * (Ct + Ct.T)
    B = (C0 + 1e-6 * torch.eye(F, dtype=torch.float64, device=dev)).contiguous(); Ct = Ct.contiguous()
    K, Li, LiT, status = ops.eig_factor(B, Ct, 1.05)
    torch.cuda.synchronize()
    Lref = torch.linalg.cholesky(K)
    Liref = torch.linalg.solve_triangular(Lref, torch.eye(F, dtype=torch.float64, device=dev), upper=False)
    print(F, "status", status.item(), "Li err", float((torch.tril(Li) - Liref).abs().max() / Liref.abs().max()),
          "LiT err", float((torch.triu(LiT) - Liref.T).abs().max() / Liref.abs().max()),
          "|Li K Li^T - I|", float((torch.tril(Li) @ K @ torch.tril(Li).T - torch.eye(F, dtype=torch.float64, device=dev)).abs().max()), flush=True)
    b = min(F, out + 8)
    X0 = linalg._start_block(F, b, dev).clone().contiguous()
    Xn = X0.clone()
    BX, CX, Gb, H = ops.eig_iterate(B, K, Ct, Li, LiT, Xn, 8, 3)
    torch.cuda.synchronize()
    # torch restatement of the same 8 steps
    Xt = X0.clone()
    def solve(Z):
        Y = Liref.mT @ (Liref @ Z)
        return Y + Liref.mT @ (Liref @ (Z - K @ Y))
    for i in range(8):
        Xt = solve(B @ Xt)
        if i & 1 and i != 3:
            Xt = Xt / torch.linalg.norm(Xt, dim=0, keepdim=True)
        if i == 3:
            G = Xt.T @ Xt
            Lg = torch.linalg.cholesky(0.5 * (G + G.T))
            Xt = torch.linalg.solve_triangular(Lg.T, Xt, upper=True, left=False)
    # compare the subspaces (columns may differ by conditioning): principal angles via projector difference
    Qn, _ = torch.linalg.qr(Xn); Qt, _ = torch.linalg.qr(Xt)
    print("   subspace diff", float((Qn @ (Qn.T @ Qt) - Qt).abs().max()), "col diff", float((Xn - Xt).abs().max() / Xt.abs().max()),
          "Gb err", float((Gb - Xn.T @ B @ Xn).abs().max()), "H err", float((H - Xn.T @ Ct @ Xn).abs().max()),
          "BX err", float((BX - B @ Xn).abs().max()), flush=True)
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(); K, Li, LiT, status = ops.eig_factor(B, Ct, 1.05); e1.record()
        Xn = X0.clone(); ops.eig_iterate(B, K, Ct, Li, LiT, Xn, 8, 3); e2.record(); torch.cuda.synchronize()
    print("   factor ms", e0.elapsed_time(e1), "iterate ms", e1.elapsed_time(e2), flush=True)
