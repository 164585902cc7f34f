"""Time the pieces of the eigen stage on the GPU. usage: python tools_dev/eig_time.py [F out]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.synthetic import feature_matrix

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
out = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
n, lag = 200000, 10
X = feature_matrix(n, F, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
s = ops.lagged_covariance(X, lag, mean, rng)
S0 = ops.symmetrize_upper(s["S0"])

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3, r

ms, (ev, V) = timeit(lambda: linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], out))
print(f"F={F} out={out}: tica_from_sums {ms:.3f} ms  stats={linalg.EIG_STATS} evals={ev.tolist()}")
linalg._PARTIAL_MIN_F = 10 ** 9
ms2, (ev2, V2) = timeit(lambda: linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], out))
print(f"dense route {ms2:.3f} ms; evals diff {float((ev - ev2).abs().max()):.2e} vec diff {float((V - V2).abs().max()):.2e}")
B = S0 / s["M"] + 1e-6 * torch.eye(F, dtype=torch.float64, device=dev)
Xb = torch.randn(F, 12, dtype=torch.float64, device=dev)
L = torch.linalg.cholesky(B)
for name, fn in [("cholesky_ex", lambda: torch.linalg.cholesky_ex(B)),
                 ("B@X", lambda: B @ Xb),
                 ("cholesky_solve", lambda: torch.cholesky_solve(Xb, L)),
                 ("eigh 12", lambda: torch.linalg.eigh(Xb.T @ Xb)),
                 ("item sync", lambda: int(torch.linalg.cholesky_ex(Xb.T @ Xb)[1].item())),
                 ("trsm small", lambda: torch.linalg.solve_triangular(torch.linalg.cholesky(Xb.T @ Xb + torch.eye(12, dtype=torch.float64, device=dev)), Xb, upper=False, left=False)),
                 ("eigh F", lambda: torch.linalg.eigh(B))]:
    print(f"  {name:16s} {timeit(fn)[0]:.3f} ms")
Bm = torch.randn(F, F, dtype=torch.float64, device=dev)
print(f"  {'F^3 GEMM':16s} {timeit(lambda: Bm @ B)[0]:.3f} ms")
print(f"  {'trsm inverse':16s} {timeit(lambda: torch.linalg.solve_triangular(L, torch.eye(F, dtype=torch.float64, device=dev), upper=False))[0]:.3f} ms")
print(f"  {'cholesky_inverse':16s} {timeit(lambda: torch.cholesky_inverse(L))[0]:.3f} ms")
