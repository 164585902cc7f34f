import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.synthetic import feature_matrix
from oracle import float64_device as f64
dev = torch.device("cuda:0")
n, f, lag = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5000, 300, 7)
X = feature_matrix(n, f, 0, n, dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
s = ops.lagged_covariance(X, lag, mean, rng, engine="tc_i8x3", xmin=st["min"], xmax=st["max"])
ref = f64.lagged_sums(X, lag, mean, rng)
# emulate the quantisation in torch (float64 arithmetic on exact integers)
KQ = 4161536.0
amax = torch.maximum((st["max"] - mean).abs(), (st["min"] - mean).abs())
e = torch.floor(torch.log2(KQ / (amax.double() * 1.0000002)))
mul = torch.pow(2.0, e).float()
d = (X - mean)                                   # float32 subtraction, as the kernel
q = torch.round(d.double() * mul.double())       # exact: d * 2^e is exact in float32
print("max |q|", float(q.abs().max()), "KQ", KQ)
scale = (torch.pow(2.0, -e) / rng.double())
zq = q * scale                                   # what the kernel contracts, exactly
M = n - lag
S0q = torch.zeros((f, f), dtype=torch.float64, device=dev); Stq = torch.zeros_like(S0q)
for c0 in range(0, M, 100000):
    c1 = min(M, c0 + 100000)
    S0q.addmm_(zq[c0:c1].T, zq[c0:c1]); Stq.addmm_(zq[c0:c1].T, zq[c0 + lag:c1 + lag])
Stq = 0.5 * (Stq + Stq.T)
Sr = 0.5 * (ref["St"] + ref["St"].T)
sc = float(ref["S0"].abs().max())
print("kernel vs emulated-exact   S0 %.3e  St %.3e" % (float((torch.triu(s["S0"]) - torch.triu(S0q)).abs().max()) / sc, float((s["St"] - Stq).abs().max()) / sc))
print("emulated-exact vs reference S0 %.3e  St %.3e" % (float((S0q - ref["S0"]).abs().max()) / sc, float((Stq - Sr).abs().max()) / sc))
E = (torch.triu(s["S0"]) - torch.triu(S0q)).abs()
i = int(E.argmax()); print("worst entry", i // f, i % f, float(E.max()), "diag err", float(torch.diagonal(E).max()))
E2 = (S0q - ref["S0"]).abs()
i2 = int(E2.argmax()); print("worst entry emulated-vs-ref", i2 // f, i2 % f, float(E2.max()), "diag", float(torch.diagonal(E2).max()))
Eo = E.clone(); Eo.fill_diagonal_(0); print("kernel-vs-exact off-diagonal max", float(Eo.max()) / sc, "rms", float((torch.triu(s["S0"]) - torch.triu(S0q)).pow(2).mean().sqrt()) / sc)
