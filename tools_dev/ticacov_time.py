import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deep_cartograph_b200 import ops
dev = torch.device("cuda:0")
for B, d in [(512, 4), (4096, 4), (65536, 4), (65536, 10), (1048576, 4), (65536, 32)]:
    f = torch.randn((B, d), device=dev); g = torch.randn((B, d), device=dev)
    for _ in range(3): ops.ticacov_sums(f, g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): s = ops.ticacov_sums(f, g)
    e1.record(); torch.cuda.synchronize()
    ref = f.double().T @ g.double()
    err = float((s["sfg"] - ref).abs().max() / ref.abs().max())
    print(f"B={B} d={d}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call, rel err {err:.1e}")
