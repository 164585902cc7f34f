"""cuSOLVER / cuBLAS FP64 building blocks of the eigen stage by size: potrf, triangular inverse (trsm with n
right-hand sides), n^3 GEMM; and a recursive triangular inverse built from them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def tri_inv(L, base):
    F = L.shape[-1]
    if F <= base:
        return torch.linalg.solve_triangular(L, torch.eye(F, dtype=L.dtype, device=L.device), upper=False)
    h = (F // 2 + 7) // 8 * 8
    Ai, Ci = tri_inv(L[:h, :h], base), tri_inv(L[h:, h:], base)
    out = torch.zeros_like(L)
    out[:h, :h] = Ai; out[h:, h:] = Ci
    out[h:, :h] = -(Ci @ (L[h:, :h] @ Ai))
    return out
for n in (125, 250, 500, 1000):
    A = torch.randn(n, n, dtype=torch.float64, device=dev); A = A @ A.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    L = torch.linalg.cholesky(A)
    I = torch.eye(n, dtype=torch.float64, device=dev)
    print(n, "potrf %.3f" % timeit(lambda: torch.linalg.cholesky_ex(A)), "trsm-inv %.3f" % timeit(lambda: torch.linalg.solve_triangular(L, I, upper=False)),
          "gemm %.3f" % timeit(lambda: A @ A), flush=True)
A = torch.randn(1000, 1000, dtype=torch.float64, device=dev); A = A @ A.T + 1000 * torch.eye(1000, dtype=torch.float64, device=dev)
L = torch.linalg.cholesky(A)
ref = torch.linalg.solve_triangular(L, torch.eye(1000, dtype=torch.float64, device=dev), upper=False)
for base in (128, 256, 512):
    got = tri_inv(L, base)
    print("recursive inverse base", base, "%.3f ms" % timeit(lambda: tri_inv(L, base)), "err", float((got - ref).abs().max() / ref.abs().max()), flush=True)
