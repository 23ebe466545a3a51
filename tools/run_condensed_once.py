"""Dev aid for ncu: one bidiag or tridiag call. usage: python tools/run_condensed_once.py bidiag|tridiag n [f64|f32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

what, n = sys.argv[1], int(sys.argv[2])
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
A0 = torch.randn((n, n), dtype=dt, device=dev)
if what == "bidiag":
    A = A0.T
    Hl = torch.zeros((n, 64), dtype=dt, device=dev).T
    Hr = torch.zeros((n - 1, 64), dtype=dt, device=dev).T
    la.bidiag_in_place(A, Hl, Hr)
else:
    S = (A0 + A0.T).T
    H = torch.zeros((n - 1, 64), dtype=dt, device=dev).T
    la.tridiag_in_place(S, H)
torch.cuda.synchronize()
print("done")
