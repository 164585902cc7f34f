"""Fused quantise + contraction kernel of the tc_i8x3 engine: parity against the float64 device checker
and against the two-kernel path on a few shapes (one and several rounds, block mode, lag 0, unaligned
rows), then timings at C2 and at the C3 level-1 shape.  usage: python tools_dev/i8_fused_check.py [quick]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.synthetic import feature_matrix
from oracle import float64_device as f64
dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"

def run(X, lag, mean, rng, block, fused, **kw):
    os.environ["DCG_I8_FUSED"] = "2" if fused else "0"
    s = ops.lagged_covariance(X, lag, mean, rng, block=block, engine="tc_i8x3", **kw)
    torch.cuda.synchronize()
    return s

def check(n, f, lag, block=0, norm=True, ld=None, seed=0):
    X = feature_matrix(n, f, 0, n, dev, seed=seed)
    if ld:
        buf = torch.zeros((n, ld), dtype=torch.float32, device=dev); buf[:, :f] = X; X = buf[:, :f]
    mean = rng = None
    if norm:
        st = ops.column_stats(X)
        mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
    s = run(X, lag, mean, rng, block, True)
    u = run(X, lag, mean, rng, block, False)
    ref = f64.lagged_sums(X, lag, mean, rng)
    ref["St"] = 0.5 * (ref["St"] + ref["St"].T)
    if lag == 0:
        s["St"] = None; u["St"] = None
    mask = torch.ones((f, f), dtype=torch.bool, device=dev).triu()
    if block:
        bm = torch.zeros((f, f), dtype=torch.bool, device=dev)
        for b0 in range(0, f, block):
            bm[b0:b0 + block, b0:b0 + block] = True
        mask &= bm
    vs = {}
    for k in ("S0", "St"):
        if s[k] is not None:
            sk, uk, rk = (torch.where(mask, t, torch.zeros_like(t)) for t in (s[k], u[k], ref[k]))
            vs[k] = float((sk - rk).norm() / rk.norm())
            vs[k + "_vs_unfused"] = float((sk - uk).abs().max() / uk.abs().max())
    vs["a"] = float((s["a"] - ref["a"]).abs().max() / max(1.0, float(ref["a"].abs().max())))
    vs["b"] = float((s["b"] - ref["b"]).abs().max() / max(1.0, float(ref["b"].abs().max())))
    vs["a_vs_unfused"] = float((s["a"] - u["a"]).abs().max())
    vs["b_vs_unfused"] = float((s["b"] - u["b"]).abs().max())
    vs["clamped"] = int(s["clamped"].item())
    print(json.dumps({"n": n, "f": f, "lag": lag, "block": block, "norm": norm, "ld": ld, **vs}), flush=True)

os.environ["DCG_I8_DEBUG"] = "1"
check(5000, 300, 7)
check(20000, 1000, 10)
check(20011, 1003, 33, block=100, ld=1004)
check(9000, 331, 0)
check(70000, 990, 10, block=99)
check(40000, 2000, 10)                      # 136 tiles: two rounds
check(30000, 4950, 10, block=495)           # hTICA level 1 of C3: 100 tiles, two rounds
check(3001, 1000, 3, norm=False)
os.environ.pop("DCG_I8_DEBUG")
if quick:
    sys.exit(0)

def timeit(label, X, lag, mean, rng, st, block=0, fused=True, reps=3, **env):
    for k, v in env.items(): os.environ[k] = str(v)
    best = 1e9
    for rep in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        os.environ["DCG_I8_FUSED"] = "2" if fused else "0"
        e0.record()
        ops.lagged_covariance(X, lag, mean, rng, block=block, engine="tc_i8x3", xmin=st["min"], xmax=st["max"])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    for k in env: os.environ.pop(k)
    print(f"{label}: {best:.2f} ms", flush=True)

n, f, lag = 1_000_000, 1000, 10
X = feature_matrix(n, f, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
timeit("C2 unfused", X, lag, mean, rng, st, fused=False)
timeit("C2 fused (defaults)", X, lag, mean, rng, st)
for ring, mb in ((2, 24), (3, 24), (3, 48), (4, 48), (4, 64), (6, 64)):
    timeit(f"C2 fused ring {ring} budget {mb} MB", X, lag, mean, rng, st, DCG_I8_RING=ring, DCG_I8_RING_BYTES=mb << 20)
for k in (2, 4, 16, 32):
    timeit(f"C2 fused window stages {k}", X, lag, mean, rng, st, DCG_I8_WINDOW_STAGES=k, DCG_I8_RING_BYTES=200 << 20)
ref = f64.lagged_sums(X, lag, mean, rng)
ref["St"] = 0.5 * (ref["St"] + ref["St"].T)
s = run(X, lag, mean, rng, 0, True, xmin=st["min"], xmax=st["max"])
print("C2 err fused", f64.sums_rel_error(s, ref), flush=True)
del X, ref, s
torch.cuda.empty_cache()
n, f = 1_250_000, 4950
ld = 4952
buf = torch.empty((n, ld), dtype=torch.float32, device=dev)
for s0 in range(0, n, 100_000):
    e0 = min(n, s0 + 100_000)
    buf[s0:e0, :f] = feature_matrix(n, f, s0, e0, dev, n_slow=14)
X = buf[:, :f]
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
timeit("C3 level-1 unfused", X, lag, mean, rng, st, block=495, fused=False, reps=2)
timeit("C3 level-1 fused (defaults)", X, lag, mean, rng, st, block=495, reps=2)
for ring, mb in ((3, 64), (4, 64), (4, 96)):
    timeit(f"C3 level-1 fused ring {ring} budget {mb} MB", X, lag, mean, rng, st, block=495, reps=2, DCG_I8_RING=ring, DCG_I8_RING_BYTES=mb << 20)
