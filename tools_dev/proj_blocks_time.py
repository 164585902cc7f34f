"""Time the hTICA level-1 block projection (one pass over X) on the GPU.
usage: python tools_dev/proj_blocks_time.py [n f block s reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import ops

a = [int(v) for v in sys.argv[1:]]
n, f, block, s = (a + [500000, 4950, 495, 5])[:4] if len(a) >= 4 else (500000, 4950, 495, 5)
reps = a[4] if len(a) > 4 else 7
dev = torch.device("cuda:0")
ld = (f + 3) // 4 * 4
buf = torch.randn((n, ld), device=dev) * 0.3 + 2.0
X = buf[:, :f]
mean = X.mean(0); rng = X.std(0)
W = torch.randn((f, s), device=dev) / block ** 0.5
for _ in range(2): ops.project_blocks(X, W, block, mean, rng)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for e0, e1 in ev:
    e0.record(); ops.project_blocks(X, W, block, mean, rng); e1.record()
torch.cuda.synchronize()
ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[reps // 2]
by = 4.0 * f * n
print(f"proj_blocks n={n} f={f} block={block} s={s}: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s  {by / ms / 1e6 / 6556.2 * 100:.1f}% of measured HBM")
