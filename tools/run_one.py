"""Run ONE hot-path call on device-resident data (for ncu): python tools/run_one.py {gemm|llt|lu} N [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200
from faer_b200 import linalg as la
op, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
if op == "gemm":
    A = torch.randn((n, n), dtype=torch.float64, device=dev).T
    B = torch.randn((n, n), dtype=torch.float64, device=dev).T
    Cm = torch.empty((n, n), dtype=torch.float64, device=dev).T
    for _ in range(reps):
        la.matmul(Cm, la.Accum.Replace, A, B, 1.0)
elif op == "llt":
    G = torch.randn((n, n), dtype=torch.float64, device=dev)
    A0 = (G @ G.T + n * torch.eye(n, dtype=torch.float64, device=dev)).T.contiguous().T
    for _ in range(reps):
        A = A0.clone()
        la.cholesky_in_place(A)
elif op == "lu":
    A0 = torch.randn((n, n), dtype=torch.float64, device=dev).T
    p = torch.empty(n, dtype=torch.int64, device=dev); pi = torch.empty(n, dtype=torch.int64, device=dev)
    for _ in range(reps):
        A = A0.clone()
        la.lu_in_place(A, p, pi)
torch.cuda.synchronize()
print("done", op, n, "launches", lib.faer_b200_launch_count())
