"""torch.profiler breakdown of the eigen stage (tica_from_sums) on the GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.synthetic import feature_matrix

F, out = 1000, 4
dev = torch.device("cuda:0")
n, lag = 200000, 10
X = feature_matrix(n, F, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
s = ops.lagged_covariance(X, lag, mean, rng)
S0 = ops.symmetrize_upper(s["S0"])
for _ in range(3):
    linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], out)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], out)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
import time
t = time.perf_counter()
for _ in range(10):
    linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], out)
torch.cuda.synchronize()
print("wall per call ms", (time.perf_counter() - t) * 100)
