import os, sys, torch
sys.path.insert(0, os.getcwd())
import faer_b200
from faer_b200 import linalg as la
dev = torch.device("cuda:0"); lib = faer_b200.load(); lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
for n in (4096, 8192):
    G = torch.randn((n, n), dtype=torch.complex128, device=dev)
    A0 = (G @ G.conj().T + n * torch.eye(n, dtype=torch.complex128, device=dev)).T.contiguous().T  # column-major
    A = A0.clone(memory_format=torch.preserve_format)
    def run():
        A.copy_(A0); la.cholesky_in_place(A)
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    L = torch.tril(A)
    x = torch.randn((n, 2), dtype=torch.complex128, device=dev)
    res = float((A0 @ x - L @ (L.conj().T @ x)).abs().max() / (A0.abs().max() * n))
    print(f"c64 LLT n={n}: {ms:.2f} ms  {4 * n**3 / 3 / ms / 1e9:.2f} TFLOP/s (4 n^3 / 3 real flop)  probe residual {res:.2e}", flush=True)
