"""GPU-side correctness of the frame-sharded path: NCCL all-gather / P2P halo / all-reduce
with the real kernels, against the single-GPU result and the float64 checker.  Needs >= 2 GPUs
(skipped otherwise); the work is in tests/_nccl_worker.py, launched with torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_workers(world: int, timeout: int = 600) -> str:
    port = 29000 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-4000:] + "\n" + res.stderr[-4000:]
    assert "NCCL_WORKER_OK" in res.stdout, res.stdout[-2000:]
    return res.stdout


@pytest.mark.timeout(900)
def test_sharded_path_equals_single_gpu_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    run_workers(min(torch.cuda.device_count(), 4))
