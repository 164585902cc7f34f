import os, sys, json
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
        s = ops.lagged_covariance(X, lag, mean, rng, engine="tc_i8x3", xmin=st["min"], xmax=st["max"])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    for k in env: os.environ.pop(k)
    print(f"{label}: {best:.2f} ms", flush=True)
t("full N=auto")
t("full N=64", DCG_I8_N=64)
t("full N=128", DCG_I8_N=128)
t("full N=96", DCG_I8_N=96)
t("no MMA (quantize + TMA + epilogue)", DCG_I8_DBG=2)
t("no TMA, no MMA (quantize + epilogue)", DCG_I8_DBG=3)
t("no TMA, no MMA, no LDTM (quantize)", DCG_I8_DBG=7)
t("item 8192", DCG_I8_ITEM_FRAMES=8192)
t("item 16384", DCG_I8_ITEM_FRAMES=16384)
