"""Dev aid: f64 GEMM timings, cp.async DMMA kernel (gemm_ws=0) vs TMA / warp-specialised DMMA kernel (gemm_ws=2) vs the opt-in
int8-sliced tcgen05 product (f64_gemm_mode=1): square products, the LLT trailing update (lower dst, k = nb), the LU update.
usage: python tools/time_gemm_modes.py [quick]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)


def time_ms(f, reps=3):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best


def modes():
    for name, ws, sl in [("cp.async", 0, 0), ("tma-ws", 2, 0), ("sliced-i8", 1, 1)]:
        lib.faer_b200_set_option(b"gemm_ws", ws); lib.faer_b200_set_option(b"f64_gemm_mode", sl)
        yield name
    lib.faer_b200_set_option(b"gemm_ws", 1); lib.faer_b200_set_option(b"f64_gemm_mode", 0)


for n in ([4096, 8192] if quick else [4096, 8192, 16384]):
    A = torch.randn((n, n), dtype=torch.float64, device=dev).T
    B = torch.randn((n, n), dtype=torch.float64, device=dev).T
    C = torch.empty((n, n), dtype=torch.float64, device=dev).T
    for kind in ["NN", "NT", "TN"]:
        a = A if kind[0] == "N" else A.T
        b = B if kind[1] == "N" else B.T
        line = f"GEMM n={n} {kind}:"
        for name in modes():
            ms = time_ms(lambda: la.matmul(C, 0, a, b, 1.0), 2)
            line += f"  {name} {ms:.2f} ms {2 * n**3 / ms / 1e9:.2f} TF"
        print(line, flush=True)
    del A, B, C

for (n, k) in [(16128, 256), (16128, 512), (16128, 1024), (8192, 256), (8192, 512), (4096, 256)]:
    P = torch.randn((k, n), dtype=torch.float64, device=dev).T  # n x k column-major
    C = torch.randn((n, n), dtype=torch.float64, device=dev).T
    line = f"SYRK(lower) n={n} k={k}:"
    for name in modes():
        if name == "sliced-i8":
            continue
        ms = time_ms(lambda: la.matmul_triangular(C, 1, 1, P, 0, P.T, 0, -1.0))
        line += f"  {name} {ms:.3f} ms {n * n * k / ms / 1e9:.2f} TF"
    print(line, flush=True)
    del P, C

for (m, n, k) in [(16384, 16384, 512), (8192, 8192, 512), (16384, 16384, 256)]:
    L = torch.randn((k, m), dtype=torch.float64, device=dev).T  # m x k column-major
    Um = torch.randn((n, k), dtype=torch.float64, device=dev).T  # k x n column-major
    C = torch.randn((n, m), dtype=torch.float64, device=dev).T
    line = f"LU update m={m} n={n} k={k}:"
    for name in modes():
        ms = time_ms(lambda: la.matmul(C, 1, L, Um, -1.0))
        line += f"  {name} {ms:.3f} ms {2 * m * n * k / ms / 1e9:.2f} TF"
    print(line, flush=True)
    del L, Um, C
