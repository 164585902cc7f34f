"""Phase timing of the e2e path (public calculator API from pinned host memory)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200.modules.cv_learning.cv_calculator import TICACalculator
from deep_cartograph_b200.modules.statistics import statistics
from deep_cartograph_b200.synthetic import feature_matrix

n, f, lag, d, k, iters = 1_000_000, 1000, 10, 4, 100, 10
dev = torch.device("cuda:0")
host = torch.empty((n, f), dtype=torch.float32, pin_memory=True)
host.copy_(feature_matrix(n, f, 0, n, dev))
cfg = {"dimension": d, "lag_time": lag, "features_normalization": "mean_std"}
out = os.path.join(ROOT, "gpurun_out", "e2e_tmp")
os.makedirs(out, exist_ok=True)

def tick(t, name, acc):
    torch.cuda.synchronize()
    now = time.perf_counter()
    acc.append((name, (now - t) * 1e3))
    return now

for rep in range(3):
    acc = []
    torch.cuda.synchronize()
    t = t0 = time.perf_counter()
    calc = TICACalculator(configuration=cfg, output_path=out)
    t = tick(t, "ctor", acc)
    calc.load_training_tensor(host)
    t = tick(t, "load(H2D+stats+spec sums)", acc)
    calc.create_output_folders()
    t = tick(t, "folders", acc)
    calc.compute_cv()
    t = tick(t, "compute_cv", acc)
    calc.set_labels()
    Pn = calc.normalize_cv()
    t = tick(t, "normalize_cv", acc)
    init = Pn[:k].to(torch.float64).clone()
    res = statistics.kmeans_lloyd(Pn, init, max_iter=iters, tol=0.0)
    t = tick(t, "kmeans_lloyd", acc)
    lab = res["labels"].cpu(); cen = res["centers"].cpu()
    t = tick(t, "d2h", acc)
    print(f"rep {rep}: total {(t - t0) * 1e3:.1f} ms | " + " | ".join(f"{a} {b:.1f}" for a, b in acc), flush=True)
