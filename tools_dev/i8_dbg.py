import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.synthetic import feature_matrix
from oracle import float64_device as f64
dev = torch.device("cuda:0")
n, f, lag = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X = feature_matrix(n, f, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
torch.cuda.synchronize()
s = ops.lagged_covariance(X, lag, mean, rng, engine="tc_i8x3")
torch.cuda.synchronize()
ref = f64.lagged_sums(X, lag, mean, rng)
ref["St"] = 0.5 * (ref["St"] + ref["St"].T)
print(f64.sums_rel_error(s, ref))
