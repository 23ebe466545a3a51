"""Multi-rank parity of the SURVEY.md 8e rows beyond LLT / LU (needs >= 2 GPUs; skipped on a single-GPU box): torchrun over NCCL, the
distributed QR (broadcast of the factored panel and its T block) on 2 / 4 / 8 ranks against the same driver run locally — factors and T
blocks to rounding, reconstruction probe through the library's block-Householder sequence — and the column-split GEMM against the
single-GPU product (tools/dist_parity.py ... qr-gemm). The schedules are covered on CPU by tests/test_dist_cpu.py (gloo, world 2-4)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_qr_and_gemm_match_single_gpu(cuda_dev, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + world), os.path.join(ROOT, "tools", "dist_parity.py"), "3072", "256", "qr-gemm"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("-> OK") == 2, out.stdout[-2000:]
