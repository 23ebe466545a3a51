"""Dev aid: failure rate of the TMA / warp-specialised f64 GEMM (gemm_ws = 2) on Add-mode products under the bring-up knobs
FAER_B200_WS_VAR / FAER_B200_WS_EXTRA_SMEM (read once per process). usage: python tools/debug_ws_variants.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
lib.faer_b200_set_option(b"gemm_ws", 2)
torch.manual_seed(7)
tag = f"VAR={os.environ.get('FAER_B200_WS_VAR', '0')} EXTRA_SMEM={os.environ.get('FAER_B200_WS_EXTRA_SMEM', '0')}"
for (n, k, kind) in [(6016, 256, "rect"), (8192, 512, "rectB"), (6016, 256, "low")]:
    A = torch.randn((k, n), dtype=torch.float64, device=dev).T
    B = A.T if kind != "rectB" else torch.randn((n, k), dtype=torch.float64, device=dev).T
    C0 = torch.randn((n, n), dtype=torch.float64, device=dev).T
    full = C0 - A @ B
    ref = (torch.tril(full) + torch.triu(C0, 1)) if kind == "low" else full
    bads = []
    for rep in range(reps):
        Cm = C0.clone(memory_format=torch.preserve_format)
        if kind == "low":
            la.matmul_triangular(Cm, 1, 1, A, 0, B, 0, -1.0)
        else:
            la.matmul(Cm, 1, A, B, -1.0)
        torch.cuda.synchronize()
        bads.append(int(((Cm - ref).abs() > 1e-9).sum()))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    la.matmul(C0, 1, A, B, -1.0) if kind != "low" else la.matmul_triangular(C0, 1, 1, A, 0, B, 0, -1.0)
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: n={n} k={k} {kind}: bad entries per rep {bads}  ({e0.elapsed_time(e1):.3f} ms)", flush=True)
    del A, B, C0, full, ref
