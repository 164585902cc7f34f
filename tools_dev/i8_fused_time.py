"""Timing of the fused tc_i8x3 kernel at C2 under the DCG_I8_DBG diagnostic modes (see ParamsF::dbg)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.synthetic import feature_matrix
dev = torch.device("cuda:0")
n, f, lag = 1_000_000, 1000, 10
X = feature_matrix(n, f, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
def t(label, **env):
    for k, v in env.items(): os.environ[k] = str(v)
    best = 1e9
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.lagged_covariance(X, lag, mean, rng, engine="tc_i8x3", xmin=st["min"], xmax=st["max"])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    for k in env: os.environ.pop(k)
    print(f"{label}: {best:.2f} ms", flush=True)
for spec in sys.argv[1:] or ["fused"]:
    env = dict(kv.split("=") for kv in spec.split(",") if "=" in kv)
    t(spec, **env)
