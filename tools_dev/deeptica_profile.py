"""torch.profiler breakdown of one DeepTICA minibatch step at the C4 shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from deep_cartograph_b200 import ops
from deep_cartograph_b200.modules.cv_learning.deep_tica import DeepTICA
from deep_cartograph_b200.synthetic import feature_matrix

n, f, d, B, lag = 1_000_000, 1000, 4, 65536, 10
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

for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
