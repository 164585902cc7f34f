"""DeepTICA minibatch step at the C4 shape: forward of both time-lagged batches through the MLP,
fused C0 / C_tau sums (dcg_ticacov_f32), Cholesky-reduced eigenvalues, loss = -sum lambda^2, backward,
Adam step.  usage: python tools_dev/deeptica_time.py [n f d batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import ops
from deep_cartograph_b200.modules.cv_learning.deep_tica import DeepTICA, tica_loss
from deep_cartograph_b200.synthetic import feature_matrix

a = [int(v) for v in sys.argv[1:]]
n = a[0] if len(a) > 0 else 2_000_000
f = a[1] if len(a) > 1 else 1000
d = a[2] if len(a) > 2 else 4
B = a[3] if len(a) > 3 else 65536
lag = 10
dev = torch.device("cuda:0")
X = torch.empty((n, f), dtype=torch.float32, device=dev)
for s0 in range(0, n, 250_000):
    X[s0:s0 + 250_000] = feature_matrix(n, f, s0, min(n, s0 + 250_000), dev)
st = ops.column_stats(X)
mean = st["mean"].float(); rng = torch.sqrt(st["m2"] / (n - 1)).float()
model = DeepTICA([f, 64, 32, d], mean, rng, activation="tanh").to(dev)
opt = torch.optim.Adam(model.nn.parameters(), lr=1e-3)
g = torch.Generator(device=dev).manual_seed(0)

def step():
    idx = torch.randint(0, n - lag, (B,), generator=g, device=dev)
    loss, ev = model.loss_indexed(X, idx, lag)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss, ev

for _ in range(5):
    step()
torch.cuda.synchronize()
reps = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    loss, ev = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# the covariance part alone
fo = torch.randn((B, d), device=dev); go = torch.randn((B, d), device=dev)
for _ in range(3): ops.ticacov_sums(fo, go)
torch.cuda.synchronize()
e0.record()
for _ in range(100): ops.ticacov_sums(fo, go)
e1.record(); torch.cuda.synchronize()
cov_ms = e0.elapsed_time(e1) / 100
print(f"DeepTICA step n={n} f={f} layers=[{f},64,32,{d}] batch={B}: {ms:.3f} ms/step = {2 * B / ms / 1e3:.1f} M frames/s through the network "
      f"(gather {2 * B * f * 4 / 1e6:.0f} MB/step); fused C0/C_tau sums alone {cov_ms * 1e3:.1f} us; loss {float(loss):.4f} evals {ev.tolist()}")
