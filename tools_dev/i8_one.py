import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.synthetic import feature_matrix
dev = torch.device("cuda:0")
n, f, lag = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 1000, 10
X = feature_matrix(n, f, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
for _ in range(2):
    s = ops.lagged_covariance(X, lag, mean, rng, engine="tc_i8x3", xmin=st["min"], xmax=st["max"])
torch.cuda.synchronize()
