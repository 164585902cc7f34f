"""hTICA block-diagonal path (level-1 block sums + block projection + level-2 pass) against the
float64 device checker and against the full-Gram path."""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200.modules.cv_learning.cv_calculator import HTICACalculator
from deep_cartograph_b200.synthetic import feature_matrix
from oracle import float64_device as f64

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
f = int(sys.argv[2]) if len(sys.argv) > 2 else 1003
ns, sd, d, lag = 10, 5, 10, 10
dev = torch.device("cuda:0")
X = feature_matrix(n, f, 0, n, dev, n_slow=14)
res = {}
for name, full_max in (("block", 0), ("full", 1 << 20)):
    cfg = {"dimension": d, "lag_time": lag, "features_normalization": "mean_std", "num_subspaces": ns,
           "subspaces_dimension": sd, "backend": {"htica_full_gram_max_features": full_max}}
    calc = HTICACalculator(configuration=cfg, output_path=tempfile.mkdtemp())
    calc.load_training_tensor(X)
    calc.create_output_folders(); calc.compute_cv(); calc.set_labels()
    P = calc.normalize_cv()
    res[name] = (torch.from_numpy(calc.cv).to(dev), P)
    mean, rng = calc._norm_on_device()
W_ref, T1, V2 = f64.htica(X, lag, mean, rng, ns, sd, d)
Pn_ref, _, _ = f64.project_normalized(X, mean, rng, W_ref)
for name, (W, P) in res.items():
    sgn = torch.sign((W.double() * W_ref).sum(0, keepdim=True))
    print(json.dumps({"path": name, "W_maxabs": float((W.double() * sgn - W_ref).abs().max()),
                      "W_col_l2": float(torch.linalg.norm(W.double() * sgn - W_ref, dim=0).max()),
                      "W_ref_col_norms": torch.linalg.norm(W_ref, dim=0).tolist()[:3],
                      "proj": float((P.double() * sgn - Pn_ref).abs().max())}), flush=True)
print("block vs full W:", float((res["block"][0] - res["full"][0]).abs().max()))
