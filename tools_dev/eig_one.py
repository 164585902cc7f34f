import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import linalg, ops
dev = torch.device("cuda:0")
F = 1000
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn((F, F + 5), generator=g, dtype=torch.float64, device=dev)
B = (A @ A.T / F + 0.1 * torch.eye(F, dtype=torch.float64, device=dev)).contiguous()
Ct = (0.5 * B).contiguous()
for _ in range(2):
    K, Li, LiT, status = ops.eig_factor(B, Ct, 1.05)
    X = linalg._start_block(F, 12, dev).clone().contiguous()
    ops.eig_iterate(B, K, Ct, Li, LiT, X, 8, 3)
torch.cuda.synchronize()
Liref = torch.linalg.inv(torch.linalg.cholesky(K))
print("Li err", float((torch.tril(Li) - Liref).abs().max() / Liref.abs().max()))
