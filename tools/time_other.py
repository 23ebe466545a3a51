"""Dev aid: device-resident timings of the other BASELINE.json configs: f64/f32/c64 GEMM, f32 QR 65536x4096 (configs[3]),
c64 GEMM n=8192 (configs[4]). usage: python tools/time_other.py [gemm|qr|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)


def time_ms(f, reps=3):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best


if what in ("gemm", "all"):
    for dt, name, fl in [(torch.float64, "f64", 2.0), (torch.float32, "f32 (3xTF32)", 2.0), (torch.complex128, "c64 (4M)", 8.0)]:
        for n in [4096, 8192] + ([16384] if dt == torch.float64 else []):
            A = torch.randn((n, n), dtype=dt, device=dev).T
            B = torch.randn((n, n), dtype=dt, device=dev).T
            Cm = torch.empty((n, n), dtype=dt, device=dev).T
            ms = time_ms(lambda: la.matmul(Cm, la.Accum.Replace, A, B, 1.0), 2)
            print(f"GEMM {name} n={n}: {ms:.2f} ms  {fl * n**3 / ms / 1e9:.2f} TFLOP/s", flush=True)
            del A, B, Cm

if what in ("qr", "all"):
    only = sys.argv[2] if len(sys.argv) > 2 else ""
    for (m, n, dt, name) in [(16384, 2048, torch.float64, "f64"), (65536, 4096, torch.float32, "f32")]:
        if only and only != name:
            continue
        A0 = torch.randn((n, m), dtype=dt, device=dev).T
        A = A0.clone(memory_format=torch.preserve_format)
        bs = la.qr_recommended_block_size(m, n)
        H = torch.zeros((n, bs), dtype=dt, device=dev).T
        tc = time_ms(lambda: A.copy_(A0))

        def run():
            A.copy_(A0)
            la.qr_in_place(A, H)
        l0 = lib.faer_b200_launch_count(); run(); nl = lib.faer_b200_launch_count() - l0
        ms = time_ms(run, 2) - tc
        flops = 2.0 * m * n * n - 2.0 * n ** 3 / 3.0
        print(f"QR {name} {m}x{n} (bs={bs}): {ms:.2f} ms  {flops / ms / 1e9:.2f} TFLOP/s  ({nl} launches)", flush=True)
        del A, A0, H
