"""C2 (1M x 1000, lag 10, d 4): every covariance engine against the float64 device checker --
full S0 / St, eigenvalues, eigenvectors, projections.  Usage: python tools_dev/c2_parity.py [n] [f]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import linalg, ops
from deep_cartograph_b200.synthetic import feature_matrix
from oracle import float64_device as f64

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
f = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
lag, d = 10, 4
dev = torch.device("cuda:0")
X = feature_matrix(n, f, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float()
rng = torch.sqrt(st["m2"] / (n - 1)).float()
t0 = time.time()
ref = f64.lagged_sums(X, lag, mean, rng)
ev_ref, V_ref = f64.tica_from_sums(ref["S0"], ref["St"], ref["a"], ref["b"], ref["M"], d + 2)
Pn_ref, _, _ = f64.project_normalized(X, mean, rng, V_ref[:, :d])
torch.cuda.synchronize()
print("float64 reference: %.1f s; eigenvalues %s" % (time.time() - t0, ev_ref.tolist()), flush=True)
for engine, kc in (("tc_i8x3", 0), ("tc_3xf16", 256), ("tc_3xtf32", 256), ("simt_f32", 0)):
    if kc:
        os.environ["DCG_TC_KC"] = str(kc)
    s = ops.lagged_covariance(X, lag, mean, rng, engine=engine, xmin=st["min"], xmax=st["max"])
    err = f64.sums_rel_error(s, ref)
    if engine == "tc_i8x3":
        err["St"] = err["St_sym"]
    S0 = ops.symmetrize_upper(s["S0"])
    # eigen stage of the product on the product's sums ...
    ev, V = linalg.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], d)
    # ... and the dense float64 checker on the product's sums (separates sum error from solver error)
    ev2, V2 = f64.tica_from_sums(S0, s["St"], s["a"], s["b"], s["M"], d)
    P, pmin, pmax = ops.project(X, V.float(), mean, rng)
    ops.standardize_(P, (pmax + pmin) / 2, (pmax - pmin) / 2)
    out = {"engine": engine, "kc": kc, **err,
           "a": float((s["a"] - ref["a"]).abs().max() / n), "b": float((s["b"] - ref["b"]).abs().max() / n),
           "eval_rel": float(((ev - ev_ref[:d]).abs() / ev_ref[:d].abs()).max()),
           "evec": f64.eigvec_error(V, V_ref[:, :d]), "evec_dense_solver": f64.eigvec_error(V2, V_ref[:, :d]),
           "evec_per_col": [f64.eigvec_error(V[:, i:i + 1], V_ref[:, i:i + 1]) for i in range(d)],
           "proj": float((P.double() - Pn_ref).abs().max())}
    print(json.dumps(out), flush=True)
