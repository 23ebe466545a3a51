"""Dev aid: the ws f64 GEMM (gemm_ws = 2) against the cp.async kernel (0) on the trailing-update shapes of LLT / LU.
usage: [FAER_B200_WS_STAGGER=0|1] python tools/time_ws_shapes.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)


def best(f, reps=5):
    f(); torch.cuda.synchronize(); b = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b


tag = f"stagger={os.environ.get('FAER_B200_WS_STAGGER', '1')}"
for (n, k) in [(16128, 256), (16128, 512), (8192, 256), (4096, 256)]:
    P = torch.randn((k, n), dtype=torch.float64, device=dev).T
    Cm = torch.randn((n, n), dtype=torch.float64, device=dev).T
    line = f"{tag} SYRK(lower) n={n} k={k}:"
    for mode in (0, 2):
        lib.faer_b200_set_option(b"gemm_ws", mode)
        ms = best(lambda: la.matmul_triangular(Cm, 1, 1, P, 0, P.T, 0, -1.0))
        line += f"  mode {mode}: {ms:.3f} ms {n * n * k / ms / 1e9:.2f} TF"
    print(line, flush=True)
    del P, Cm
for (m, n, k) in [(16384, 16384, 256), (16384, 16384, 512), (8192, 8192, 512), (16384, 512, 512), (16384, 256, 256)]:
    L = torch.randn((k, m), dtype=torch.float64, device=dev).T
    U = torch.randn((n, k), dtype=torch.float64, device=dev).T
    Cm = torch.randn((n, m), dtype=torch.float64, device=dev).T
    line = f"{tag} update m={m} n={n} k={k}:"
    for mode in (0, 2):
        lib.faer_b200_set_option(b"gemm_ws", mode)
        ms = best(lambda: la.matmul(Cm, 1, L, U, -1.0))
        line += f"  mode {mode}: {ms:.3f} ms {2.0 * m * n * k / ms / 1e9:.2f} TF"
    print(line, flush=True)
    del L, U, Cm
lib.faer_b200_set_option(b"gemm_ws", 1)
