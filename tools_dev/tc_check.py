"""Quick on-GPU check of the tcgen05 covariance engine against a torch float64 evaluation.
usage: python tools_dev/tc_check.py [n f lag] ; env DCG_TC_A_TMEM, DCG_TC_KC select variants."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import ops

def main():
    n, f, lag = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4000, 256, 10)
    engine = sys.argv[4] if len(sys.argv) > 4 else "tc_3xtf32"
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn((n, f), generator=g, device=dev) * 0.3 + 2.0
    mean = X.mean(0); rng = X.std(0)
    Z = ((X - mean) / rng).double()
    M = n - lag
    S0 = Z[:M].T @ Z[:M]; St = Z[:M].T @ Z[lag:]
    for std in (True, False):
        t0 = time.time()
        if std:
            s = ops.lagged_covariance(X, lag, mean, rng, engine=engine)
        else:
            s = ops.lagged_covariance(((X - mean) / rng).contiguous(), lag, engine=engine)
        torch.cuda.synchronize()
        e0 = (torch.triu(s["S0"] - S0)).abs().max().item() / S0.abs().max().item()
        et = (s["St"] - St).abs().max().item() / St.abs().max().item()
        print(f"n={n} f={f} lag={lag} engine={engine} std={std} A_TMEM={os.environ.get('DCG_TC_A_TMEM','1')} "
              f"KC={os.environ.get('DCG_TC_KC','1024')}: rel err S0 {e0:.3e} St {et:.3e}  ({time.time()-t0:.3f}s)", flush=True)

if __name__ == "__main__":
    main()
