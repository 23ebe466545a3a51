"""Device-resident timings of the complex factorizations (cplx.cu) — functional single-SM panels, see DESIGN.md 5 / 7:
c64 and c32 LLT, partial-pivoting LU and Householder QR at n = 2048 / 4096, one warm-up + one timed call each (CUDA events on the
library's stream), with a reconstruction probe. usage: python tools/time_cplx.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import faer_b200
from faer_b200 import linalg as la

dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)


def timed(f):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for dt, name in ((torch.complex128, "c64"), (torch.complex64, "c32")):
    for n in (2048, 4096):
        torch.manual_seed(n)
        G = torch.randn((n, n), dtype=dt, device=dev).T.contiguous().T
        x = torch.randn((n, 2), dtype=dt, device=dev)
        # LLT
        S0 = (G @ G.conj().T + n * torch.eye(n, dtype=dt, device=dev)).T.contiguous().T
        A = S0.clone(memory_format=torch.preserve_format)
        ms = timed(lambda: (A.copy_(S0), la.cholesky_in_place(A)))
        L = torch.tril(A)
        res = float((S0 @ x - L @ (L.conj().T @ x)).abs().max() / (S0.abs().max() * n))
        print(f"{name} LLT n={n}: {ms:8.2f} ms  {4 * n**3 / 3 / ms / 1e9:6.2f} TFLOP/s  probe {res:.1e}", flush=True)
        # LU
        A = G.clone(memory_format=torch.preserve_format)
        p = torch.zeros(n, dtype=torch.int64, device=dev); pi = torch.zeros(n, dtype=torch.int64, device=dev)
        ms = timed(lambda: (A.copy_(G), la.lu_in_place(A, p, pi)))
        Lm = torch.tril(A, -1) + torch.eye(n, dtype=dt, device=dev); U = torch.triu(A)
        res = float((G[p] @ x - Lm @ (U @ x)).abs().max() / (G.abs().max() * n))
        print(f"{name} LU  n={n}: {ms:8.2f} ms  {8 * n**3 / 3 / ms / 1e9:6.2f} TFLOP/s  probe {res:.1e}", flush=True)
        # QR
        bs = la.qr_recommended_block_size(n, n)
        A = G.clone(memory_format=torch.preserve_format)
        H = torch.zeros((n, bs), dtype=dt, device=dev).T
        ms = timed(lambda: (A.copy_(G), la.qr_in_place(A, H)))
        y = (torch.triu(A) @ x).T.contiguous().T
        la.apply_block_householder_sequence_on_the_left_in_place(A, H, y)
        res = float((G @ x - y).abs().max() / (G.abs().max() * n))
        print(f"{name} QR  n={n} bs={bs}: {ms:8.2f} ms  {16 * n**3 / 3 / ms / 1e9:6.2f} TFLOP/s  probe {res:.1e}", flush=True)
