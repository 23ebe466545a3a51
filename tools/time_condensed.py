"""Dev aid: device-resident timings of bidiag / tridiag (HBM-bound stages). usage: python tools/time_condensed.py [n ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192]


def time_ms(f, reps=2):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best


for dt, name, sz in [(torch.float64, "f64", 8), (torch.float32, "f32", 4)]:
    for n in sizes:
        A0 = torch.randn((n, n), dtype=dt, device=dev)
        A = A0.clone().T
        Hl = torch.zeros((n, 64), dtype=dt, device=dev).T
        Hr = torch.zeros((n - 1, 64), dtype=dt, device=dev).T
        tc = time_ms(lambda: A.copy_(A0.T))

        def run_b():
            A.copy_(A0.T)
            la.bidiag_in_place(A, Hl, Hr)
        ms = time_ms(run_b) - tc
        byts = 3.0 * sz * n ** 3 / 3.0  # sum_k 3 * s * (n-k)^2
        print(f"bidiag {name} n={n}: {ms:.2f} ms  {8 / 3 * n**3 / ms / 1e9:.3f} TFLOP/s  {byts / ms / 1e6:.1f} GB/s algorithmic", flush=True)
        if n <= 8192:
            S0 = (A0 + A0.T)
            S = S0.clone().T
            H = torch.zeros((n - 1, 64), dtype=dt, device=dev).T

            def run_t():
                S.copy_(S0)
                la.tridiag_in_place(S, H)
            ms = time_ms(run_t) - tc
            byts = 2.0 * sz * n ** 3 / 6.0  # sum_k 2 * s * (n-k)^2 / 2
            print(f"tridiag {name} n={n}: {ms:.2f} ms  {4 / 3 * n**3 / ms / 1e9:.3f} TFLOP/s  {byts / ms / 1e6:.1f} GB/s algorithmic", flush=True)
        del A0, A
