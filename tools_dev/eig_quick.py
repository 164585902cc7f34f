"""Eigen stage: wall time, device time (CUDA events) and iteration count on two synthetic spectra."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.synthetic import feature_matrix
dev = torch.device("cuda:0")
for n in (200_000, 1_000_000):
    F, lag, out = 1000, 10, 4
    X = feature_matrix(n, F, 0, n, dev)
    st = ops.column_stats(X)
    mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
    s = ops.lagged_covariance(X, lag, mean, rng, xmin=st["min"], xmax=st["max"])
    S0 = ops.symmetrize_upper(s["S0"])
    del X
    for rep in range(3):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        e0.record()
        ev, V = linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], out)
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t) * 1e3
    linalg_stats = dict(linalg.EIG_STATS)
    old = linalg._PARTIAL_MIN_F; linalg._PARTIAL_MIN_F = 10 ** 9
    ev2, V2 = linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], out)
    linalg._PARTIAL_MIN_F = old
    print(f"n={n}: wall {wall:.3f} ms, device {e0.elapsed_time(e1):.3f} ms, {linalg_stats}, vs dense: evals {float((ev - ev2).abs().max()):.1e} vecs {float((V - V2).abs().max()):.1e}", flush=True)
