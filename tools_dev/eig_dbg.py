import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops
dev = torch.device("cuda:0")
F = int(sys.argv[1]); mode = sys.argv[2]
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn((F, F + 5), generator=g, dtype=torch.float64, device=dev)
B = (A @ A.T / F + 0.1 * torch.eye(F, dtype=torch.float64, device=dev)).contiguous()
Ct = (0.5 * B).contiguous()
print("start", F, mode, flush=True)
K, Li, LiT, status = ops.eig_factor(B, Ct, 1.05)
torch.cuda.synchronize()
Lref = torch.linalg.cholesky(K)
Liref = torch.linalg.inv(Lref)
print("factor ok; status", status.item(), "Li err", float((torch.tril(Li) - Liref).abs().max() / Liref.abs().max()), "LiT err", float((torch.triu(LiT) - Liref.T).abs().max() / Liref.abs().max()), flush=True)
if mode == "iter":
    b = 12
    X = torch.randn((F, b), generator=g, dtype=torch.float64, device=dev).contiguous()
    X0 = X.clone()
    BX, CX, Gb, H = ops.eig_iterate(B, K, Ct, Liref.contiguous(), Liref.T.contiguous(), X, 2, -1)
    torch.cuda.synchronize()
    print("iterate ok", float((BX - B @ X).abs().max()), flush=True)
