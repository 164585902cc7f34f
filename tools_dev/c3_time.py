"""hTICA on one GPU's shard of the C3 configuration (synthetic n x 4950 features, 10 subspaces of
495, subspace dimension 5, d = 10, lag 10) through HTICACalculator + kmeans_lloyd, timed per
stage with CUDA events.  usage: python tools_dev/c3_time.py [n f k kmeans_iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.modules.cv_learning.cv_calculator import HTICACalculator
from deep_cartograph_b200.modules.statistics import statistics
from deep_cartograph_b200.synthetic import feature_matrix

a = [int(v) for v in sys.argv[1:]]
n = a[0] if len(a) > 0 else 500_000
f = a[1] if len(a) > 1 else 4950
k = a[2] if len(a) > 2 else 1000
iters = a[3] if len(a) > 3 else 5
dev = torch.device("cuda:0")
lag, d = 10, 10
ld = (f + 3) // 4 * 4
buf = torch.empty((n, ld), dtype=torch.float32, device=dev)
step = 100_000
for s0 in range(0, n, step):                       # generate in chunks (bounded temporaries)
    e0 = min(n, s0 + step)
    buf[s0:e0, :f] = feature_matrix(n, f, s0, e0, dev, n_slow=14)
X = buf[:, :f]
torch.cuda.synchronize()
cfg = {"dimension": d, "lag_time": lag, "features_normalization": "mean_std",
       "num_subspaces": 10, "subspaces_dimension": 5}

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for rep in range(2):
    t = [ev()]
    calc = HTICACalculator(configuration=cfg, output_path=os.path.join(ROOT, "gpurun_out", "c3_tmp"))
    calc.load_training_tensor(X)                                     # stats
    t.append(ev())
    calc.create_output_folders()
    calc.cv_dimension = d
    launches0 = ops.KERNEL_LAUNCHES
    calc.compute_cv()                                                # level 1 + level 2
    t.append(ev())
    calc.set_labels()
    Pn = calc.normalize_cv()                                         # projection + CV min/max
    t.append(ev())
    init = Pn[:k].to(torch.float64).clone()
    res = statistics.kmeans_lloyd(Pn, init, max_iter=iters, tol=0.0)
    t.append(ev())
    torch.cuda.synchronize()
    ms = [t[i].elapsed_time(t[i + 1]) for i in range(len(t) - 1)]
    tot = sum(ms)
    print(f"rep {rep}: n={n} f={f} (ld {ld}): stats {ms[0]:.1f} ms | hTICA sums+eig {ms[1]:.1f} ms | projection {ms[2]:.1f} ms | "
          f"KMeans k={k} x {res['n_iter']} iters {ms[3]:.1f} ms | total {tot:.1f} ms = {n / tot / 1e3:.2f} Mframes/s; "
          f"eig stats {linalg.EIG_STATS}; launches {ops.KERNEL_LAUNCHES - launches0}", flush=True)
# finer breakdown of compute_cv (block-diagonal level 1, eigen stage, level 2)
chunks = linalg.htica_chunks(f, 10)
for rep in range(2):
    t = [ev()]
    s1 = calc._lagged_sums(lag, block=f // 10)
    t.append(ev())
    T1 = linalg.htica_level1(s1["S0"], s1["St"], s1["a"], s1["b"], s1["M"], chunks, 5)
    t.append(ev())
    mean, rng = calc._norm_on_device()
    T1f = T1.to(torch.float32)
    Wc = torch.zeros((f, 5), dtype=torch.float32, device=dev)
    c = 0
    for (s0, e0) in chunks:
        Wc[s0:e0] = T1f[s0:e0, c:c + 5]; c += 5
    P = ops.project_blocks(calc.training_data, Wc, f // 10, mean, rng)      # all blocks, one pass over X
    t.append(ev())
    s2 = ops.lagged_covariance(P, lag)
    t.append(ev())
    _, V2 = linalg.tica_from_sums(ops.symmetrize_upper(s2["S0"]), s2["St"], s2["a"], s2["b"], s2["M"], d)
    t.append(ev())
    torch.cuda.synchronize()
    ms = [t[i].elapsed_time(t[i + 1]) for i in range(len(t) - 1)]
    print(f"compute_cv breakdown rep {rep}: level-1 sums {ms[0]:.1f} | level-1 eig (10 x 495, batched) {ms[1]:.1f} | "
          f"level-1 projection {ms[2]:.1f} | level-2 sums (50 feat.) {ms[3]:.1f} | level-2 eig {ms[4]:.1f} ms", flush=True)
W = torch.as_tensor(calc.cv)
print("cv weights", tuple(W.shape), "finite", bool(torch.isfinite(W).all()))
