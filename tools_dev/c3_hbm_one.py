"""One statistics pass and one projection pass at the C3 per-GPU footprint (1.25M x 4950, 24.75 GB) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
dev = torch.device("cuda:0")
n, f = 1_250_000, 4950
ld = 4952
buf = torch.empty((n, ld), dtype=torch.float32, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
for s0 in range(0, n, 125_000):
    buf[s0:s0 + 125_000].normal_(generator=g)
X = buf[:, :f]
W = torch.randn((f, 10), generator=g, device=dev)
mean = torch.zeros(f, device=dev); rng = torch.ones(f, device=dev)
for _ in range(2):
    st = ops.column_stats(X)
    P, pmin, pmax = ops.project(X, W, mean, rng)
torch.cuda.synchronize()
for name, fn in (("stats", lambda: ops.column_stats(X)), ("project", lambda: ops.project(X, W, mean, rng))):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(name, "%.3f ms  %.1f GB/s" % (ms, 4.0 * ld * n / ms / 1e6), flush=True)
