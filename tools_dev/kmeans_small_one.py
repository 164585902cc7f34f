"""A few Lloyd iterations at the C2 KMeans shape (1M x 4 float32, k = 100) for ncu / timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
dev = torch.device("cuda:0")
n, d, k = 1_000_000, 4, 100
g = torch.Generator(device=dev).manual_seed(0)
# a slow random walk squashed into [-1, 1]: time-ordered frames share labels, like projected MD frames
Y = torch.tanh(torch.cumsum(torch.randn((n, d), generator=g, device=dev) * 0.01, 0)).float().contiguous()
C = Y[:k].double().clone()
labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
work = ops.kmeans_work(k, d, dev)
b = torch.full((1,), 1.001, dtype=torch.float64, device=dev)
for it in range(3):
    ops.kmeans_iterate_(Y, C, labels, work, absmax=b)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    ops.kmeans_iterate_(Y, C, labels, work, absmax=b)
e1.record(); torch.cuda.synchronize()
print(f"10 iterations: {e0.elapsed_time(e1):.3f} ms")
