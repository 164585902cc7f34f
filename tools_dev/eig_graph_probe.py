"""Probe: is the fixed part of the shift-and-invert eigen stage (factor, 8 steps, Cholesky QR,
Rayleigh-Ritz) faster as one CUDA graph replay than as eager torch ops?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.synthetic import feature_matrix

F, out = 1000, 4
dev = torch.device("cuda:0")
n, lag = 200000, 10
X = feature_matrix(n, F, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
s = ops.lagged_covariance(X, lag, mean, rng)
S0 = ops.symmetrize_upper(s["S0"]); M = s["M"]
mu = s["a"] / M; nu = s["b"] / M
C0 = S0 / M - torch.outer(mu, mu); C0 = 0.5 * (C0 + C0.T)
Ct = s["St"] / M - torch.outer(mu, nu); Ct = 0.5 * (Ct + Ct.T)
B = (C0 + 1e-6 * torch.eye(F, dtype=torch.float64, device=dev)).unsqueeze(0).contiguous()
Ct = Ct.unsqueeze(0).contiguous()
b = out + 8
eye = torch.eye(F, dtype=torch.float64, device=dev)
eye_b = torch.eye(b, dtype=torch.float64, device=dev)
X0 = linalg._start_block(F, b, dev).expand(1, F, b).contiguous()
nrm = torch.linalg.matrix_norm(Ct).unsqueeze(-1)

def fixed(Bm, Cm):
    Kmat = 1.05 * Bm - Cm
    Lk, info = torch.linalg.cholesky_ex(Kmat)
    Li = torch.linalg.solve_triangular(Lk, eye.expand(1, F, F), upper=False)
    Xc = X0
    for i in range(8):
        Z = Bm @ Xc
        Y = Li.mT @ (Li @ Z)
        Xc = torch.baddbmm(Y, Li.mT, Li @ torch.baddbmm(Z, Kmat, Y, alpha=-1.0))
        if i & 1:
            Xc = Xc / torch.linalg.norm(Xc, dim=-2, keepdim=True)
        if i == 3:
            G = Xc.mT @ Xc
            Lg, _ = torch.linalg.cholesky_ex(0.5 * (G + G.mT))
            Xc = torch.linalg.solve_triangular(Lg.mT, Xc, upper=True, left=False)
    BX = Bm @ Xc; CX = Cm @ Xc
    Gb = Xc.mT @ BX
    theta, S, info2 = ops.gen_eig_small(Xc.mT @ CX, Gb)
    Xr = Xc @ S
    res = torch.linalg.norm(CX @ S[..., :out] - (BX @ S[..., :out]) * theta[:, None, :out], dim=-2)
    rel = res / (nrm * torch.linalg.norm(Xr[..., :out], dim=-2))
    return torch.cat([rel.max().reshape(1), theta[0, :out], info.double().reshape(-1), info2.double().reshape(-1)]), Xr

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        r = fn()
        r[0].tolist()          # the one host read
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3, r

ms_e, r_e = timeit(lambda: fixed(B, Ct))
print(f"eager fixed sequence: {ms_e:.3f} ms  out={r_e[0].tolist()[:5]}")
try:
    Bs, Cs = B.clone(), Ct.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fixed(Bs, Cs)
    torch.cuda.current_stream().wait_stream(side)
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        outs = fixed(Bs, Cs)
    def replay():
        Bs.copy_(B); Cs.copy_(Ct)
        gph.replay()
        return outs
    ms_g, r_g = timeit(replay)
    print(f"graph replay:         {ms_g:.3f} ms  out={r_g[0].tolist()[:5]}")
except Exception as e:  # noqa: BLE001
    print("graph capture failed:", type(e).__name__, str(e)[:300])
