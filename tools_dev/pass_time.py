"""Time the HBM-bound passes alone on the GPU (CUDA events): stats, projection, KMeans step.
usage: python tools_dev/pass_time.py proj|stats|kmeans [n f d k reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import ops

what = sys.argv[1]
a = [int(v) for v in sys.argv[2:]]
dev = torch.device("cuda:0")
HBM = 6556.2

def timeit(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return ms[len(ms) // 2]

if what in ("proj", "stats"):
    n, f, d = (a + [1000000, 1000, 4])[:3] if len(a) >= 3 else (1000000, 1000, 4)
    reps = a[3] if len(a) > 3 else 7
    ld = (f + 3) // 4 * 4
    X = (torch.randn((n, ld), device=dev) * 0.3 + 2.0)[:, :f]
    mean = X.mean(0); rng = X.std(0)
    W = torch.randn((f, d), device=dev) / f ** 0.5
    if what == "proj":
        ms = timeit(lambda: ops.project(X, W, mean, rng), reps)
        by = (4.0 * f + 4.0 * d) * n
    else:
        ms = timeit(lambda: ops.column_stats(X), reps)
        by = 4.0 * f * n
    print(f"{what} n={n} f={f} d={d}: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s  {by / ms / 1e6 / HBM * 100:.1f}% of measured HBM")
else:
    n, d, k = (a + [12500000, 10, 1000])[:3] if len(a) >= 3 else (12500000, 10, 1000)
    reps = a[3] if len(a) > 3 else 5
    g = torch.Generator(device=dev).manual_seed(2)
    cen = torch.rand((k, d), generator=g, device=dev) * 1.8 - 0.9
    idx = torch.randint(0, k, (n,), generator=g, device=dev)
    Y = (cen[idx] + 0.03 * torch.randn((n, d), generator=g, device=dev)).contiguous()
    C = Y[:k].to(torch.float64).clone()
    labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
    bound = Y.abs().amax().to(torch.float64).reshape(1) if os.environ.get("KM_FIXED", "1") == "1" else None
    ms = timeit(lambda: ops.kmeans_step(Y, C, labels, absmax=bound), reps)
    by = (4.0 * d + 4.0) * n
    fl = 2.0 * k * d * n
    print(f"kmeans n={n} d={d} k={k}: {ms:.3f} ms  {n / ms / 1e3:.1f} Mframes/s  {by / ms / 1e6:.0f} GB/s ({by / ms / 1e6 / HBM * 100:.1f}% HBM)  {fl / ms / 1e9:.1f} TFLOP/s fp32-equivalent")
