"""Time the covariance engines on the GPU (CUDA events) and check them against torch float64.
usage: python tools_dev/cov_time.py [n f lag engine reps]   (env DCG_TC_KC etc. select variants)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import ops


def main():
    a = sys.argv[1:]
    n, f, lag = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (200000, 1000, 10)
    engine = a[3] if len(a) > 3 else "tc_3xtf32"
    reps = int(a[4]) if len(a) > 4 else 5
    check = os.environ.get("COV_CHECK", "1") == "1"
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn((n, f), generator=g, device=dev) * 0.3 + 2.0
    # correlate the columns a little so that off-diagonal sums are not all ~0
    X += X.roll(1, dims=1) * 0.5
    mean = X.mean(0); rng = X.std(0)
    for _ in range(2):
        s = ops.lagged_covariance(X, lag, mean, rng, engine=engine)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        e0.record(); s = ops.lagged_covariance(X, lag, mean, rng, engine=engine); e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    med = ms[len(ms) // 2]
    M = n - lag
    alg = 3.0 * f * f * M
    msg = (f"n={n} f={f} lag={lag} {engine} kc={os.environ.get('DCG_TC_KC', 'dflt')}: median {med:.3f} ms (min {ms[0]:.3f}) "
           f"{n / med * 1e3 / 1e6:.2f} Mframes/s  alg {alg / med / 1e9:.1f} TF/s  issued(x3) {3 * alg / med / 1e9:.1f} TF/s")
    if check:
        Z = ((X - mean) / rng).double()
        S0 = Z[:M].T @ Z[:M]; St = Z[:M].T @ Z[lag:]
        e0 = (torch.triu(s["S0"] - S0)).abs().max().item() / S0.abs().max().item()
        et = (s["St"] - St).abs().max().item() / St.abs().max().item()
        d0 = ((torch.diagonal(s["S0"]) - torch.diagonal(S0)) / torch.diagonal(S0))
        msg += f" | rel err S0 {e0:.2e} St {et:.2e} diag bias {d0.mean().item():+.2e}"
    print(msg, flush=True)


if __name__ == "__main__":
    main()
