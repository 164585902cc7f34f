"""One call of the fused tc_i8x3 kernel (for ncu).  usage: python tools_dev/i8_fused_one.py [n f block]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.synthetic import feature_matrix
dev = torch.device("cuda:0")
a = [int(v) for v in sys.argv[1:]]
n = a[0] if len(a) > 0 else 1_000_000
f = a[1] if len(a) > 1 else 1000
block = a[2] if len(a) > 2 else 0
lag = 10
ld = (f + 3) // 4 * 4
buf = torch.empty((n, ld), dtype=torch.float32, device=dev)
for s0 in range(0, n, 100_000):
    e0 = min(n, s0 + 100_000)
    buf[s0:e0, :f] = feature_matrix(n, f, s0, e0, dev)
X = buf[:, :f]
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
for rep in range(2):
    ops.lagged_covariance(X, lag, mean, rng, block=block, engine="tc_i8x3", xmin=st["min"], xmax=st["max"])
torch.cuda.synchronize()
