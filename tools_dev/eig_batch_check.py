"""Determinism / timing check of the batched level-1 eigen stage (10 pencils of 495) and of the
small Ritz-pencil kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.synthetic import feature_matrix

dev = torch.device("cuda:0")
n, f, lag = 200000, 4950, 10
ld = 4952
buf = torch.empty((n, ld), dtype=torch.float32, device=dev)
for s0 in range(0, n, 50000):
    buf[s0:s0 + 50000, :f] = feature_matrix(n, f, s0, min(n, s0 + 50000), dev, n_slow=14)
X = buf[:, :f]
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
s = ops.lagged_covariance(X, lag, mean, rng, block=495)
S0 = ops.symmetrize_upper(s["S0"])
chunks = linalg.htica_chunks(f, 10)
outs = []
for rep in range(8):
    before = dict(linalg.EIG_STATS)
    torch.cuda.synchronize(); t = time.perf_counter()
    T1 = linalg.htica_level1(S0, s["St"], s["a"], s["b"], s["M"], chunks, 5)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t) * 1e3
    after = dict(linalg.EIG_STATS)
    outs.append(T1)
    print(f"rep {rep}: {ms:.2f} ms fast+{after['fast'] - before['fast']} dense+{after['dense'] - before['dense']} iters {after['last_iters']} "
          f"same_as_rep0 {bool(torch.equal(T1, outs[0]))} maxdiff {float((T1 - outs[0]).abs().max()):.2e}")
g0 = torch.Generator(device=dev).manual_seed(1)
R = torch.randn((10, 13, 16), generator=g0, device=dev, dtype=torch.float64)
G = R @ R.mT + 0.1 * torch.eye(13, device=dev, dtype=torch.float64)
Q = torch.randn((10, 13, 13), generator=g0, device=dev, dtype=torch.float64); H = 0.5 * (Q + Q.mT)
th0, S0_, _ = ops.gen_eig_small(H, G)
same = all(torch.equal(ops.gen_eig_small(H, G)[1], S0_) for _ in range(200))
print("gen_eig_small deterministic over 200 runs:", same)
