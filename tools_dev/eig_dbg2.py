import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_cartograph_b200 import ops, _lib
dev = torch.device("cuda:0")
F = int(sys.argv[1])
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn((F, F + 5), generator=g, dtype=torch.float64, device=dev)
B = (A @ A.T / F + 0.1 * torch.eye(F, dtype=torch.float64, device=dev)).contiguous()
Ct = (0.5 * B).contiguous()
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
K = torch.empty_like(B)
print("shift", flush=True)
_lib.call("dcg_eig_shift_matrix_f64", B.data_ptr(), Ct.data_ptr(), F, 1.05, K.data_ptr(), st)
torch.cuda.synchronize(); print("shift done", float(K.abs().max()), flush=True)
L = K.clone(); Li = torch.empty_like(B); LiT = torch.empty_like(B); status = torch.empty(1, dtype=torch.float64, device=dev)
nbytes = lib.dcg_eig_chol_inv_workspace_bytes(F); print("ws", nbytes, flush=True)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
print("chol launch", flush=True)
import time
torch.cuda.synchronize(); t0 = time.time()
rc = lib.dcg_eig_chol_inv_f64(L.data_ptr(), F, Li.data_ptr(), LiT.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), st)
print("chol returned", rc, flush=True)
torch.cuda.synchronize(); print("chol done", status.item(), "seconds", time.time() - t0, flush=True)
bar = ws[nbytes - 256:nbytes - 256 + 32].view(torch.int32)
print("barrier words [count, abort, epoch, seen, cta]", bar[:5].tolist(), flush=True)
Lref = torch.linalg.cholesky(K)
print("L err", float((torch.tril(L) - Lref).abs().max()), flush=True)
