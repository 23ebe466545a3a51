"""Dev aid for ncu: three LLT-style trailing updates (lower destination, Add) on one GEMM kernel.
usage: python tools/run_syrk_once.py <gemm_ws mode> <n> <k> [rect]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

mode, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rect = len(sys.argv) > 4
dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
lib.faer_b200_set_option(b"gemm_ws", mode)
P = torch.randn((k, n), dtype=torch.float64, device=dev).T
C = torch.randn((n, n), dtype=torch.float64, device=dev).T
for _ in range(3):
    if rect:
        la.matmul(C, la.Accum.Add, P, P.T, -1.0)
    else:
        la.matmul_triangular(C, 1, la.Accum.Add, P, 0, P.T, 0, -1.0)
torch.cuda.synchronize()
print("done", mode, n, k, rect)
