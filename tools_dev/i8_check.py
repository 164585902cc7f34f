"""tc_i8x3 engine against the float64 device checker on a few shapes, then its time at C2."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.synthetic import feature_matrix
from oracle import float64_device as f64
dev = torch.device("cuda:0")

def check(n, f, lag, block=0, norm=True, ld=None, seed=0):
    X = feature_matrix(n, f, 0, n, dev, seed=seed)
    if ld:
        buf = torch.zeros((n, ld), dtype=torch.float32, device=dev); buf[:, :f] = X; X = buf[:, :f]
    mean = rng = None
    if norm:
        st = ops.column_stats(X)
        mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
    s = ops.lagged_covariance(X, lag, mean, rng, block=block, engine="tc_i8x3")
    torch.cuda.synchronize()
    ref = f64.lagged_sums(X, lag, mean, rng)
    ref["St"] = 0.5 * (ref["St"] + ref["St"].T)            # the i8 engine returns the symmetric part
    if lag == 0:
        s["St"] = None
    if block:
        mask = torch.zeros((f, f), dtype=torch.bool, device=dev)
        for b0 in range(0, f, block):
            mask[b0:b0 + block, b0:b0 + block] = True
        for k in ("S0", "St"):
            if s[k] is not None:
                s[k] = torch.where(mask, s[k], torch.zeros_like(s[k])); ref[k] = torch.where(mask, ref[k], torch.zeros_like(ref[k]))
    err = f64.sums_rel_error(s, ref)
    err["a"] = float((s["a"] - ref["a"]).abs().max() / max(1.0, float(ref["a"].abs().max())))
    err["b"] = float((s["b"] - ref["b"]).abs().max() / max(1.0, float(ref["b"].abs().max())))
    err["clamped"] = int(s["clamped"].item())
    print(json.dumps({"n": n, "f": f, "lag": lag, "block": block, "norm": norm, "ld": ld, **err}), flush=True)

check(164, 54, 1)
check(5000, 300, 7)
check(5000, 300, 7, norm=False)
check(20000, 1000, 10)
check(20011, 1003, 33, block=100, ld=1004)
check(9000, 331, 0)
check(70000, 4950 // 5, 10, block=99)
# C2 timing
n, f, lag = 1_000_000, 1000, 10
X = feature_matrix(n, f, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
for eng in ("tc_i8x3", "tc_3xf16"):
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        s = ops.lagged_covariance(X, lag, mean, rng, engine=eng, xmin=st["min"], xmax=st["max"])
        e1.record(); torch.cuda.synchronize()
        print(eng, "C2 ms", e0.elapsed_time(e1), flush=True)
ref = f64.lagged_sums(X, lag, mean, rng)
ref["St"] = 0.5 * (ref["St"] + ref["St"].T)
s = ops.lagged_covariance(X, lag, mean, rng, engine="tc_i8x3", xmin=st["min"], xmax=st["max"])
print("C2 err", f64.sums_rel_error(s, ref), flush=True)
