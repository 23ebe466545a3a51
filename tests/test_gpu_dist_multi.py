"""Multi-rank parity on hardware (needs >= 2 GPUs; skipped on a single-GPU box): torchrun over NCCL, the distributed LLT and LU on
2 (and 4 / 8 when visible) ranks against the single-GPU run of the same matrices — permutations / status bit-exact, factors to
rounding, reconstruction probes (tools/dist_parity.py; the distributed QR and the column-split GEMM: test_gpu_zzzzzzzz_dist_multi_qr_gemm.py). The world-size-2 logic of the layout is covered on CPU by
tests/test_dist_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_factorizations_match_single_gpu(cuda_dev, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "tools", "dist_parity.py"), "3072", "256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("-> OK") == 2, out.stdout[-2000:]
