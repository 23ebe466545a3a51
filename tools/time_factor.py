"""Dev aid: device-resident timings of LLT / LU through the C ABI. usage: python tools/time_factor.py [llt|lu|all] [sizes...]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200
from faer_b200 import linalg as la
what = sys.argv[1] if len(sys.argv) > 1 else "all"
sizes = [int(x) for x in sys.argv[2:]] or [4096, 8192, 16384]
dev = torch.device("cuda:0"); lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)

def time_ms(f, reps=3):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

out = {}
for n in sizes:
    if what in ("llt", "all"):
        G = torch.randn((n, n), dtype=torch.float64, device=dev)
        A0 = torch.addmm(n * torch.eye(n, dtype=torch.float64, device=dev), G, G.T).T; del G
        A = A0.clone(memory_format=torch.preserve_format)
        tc = time_ms(lambda: A.copy_(A0))
        def run():
            A.copy_(A0); la.cholesky_in_place(A)
        l0 = lib.faer_b200_launch_count(); run(); nl = lib.faer_b200_launch_count() - l0
        ms = time_ms(run) - tc
        print(f"LLT n={n}: {ms:.2f} ms  {n**3/3/ms/1e9:.2f} TFLOP/s  ({nl} launches)", flush=True)
        out[f"llt_{n}"] = n**3/3/ms/1e9
        del A, A0
    if what in ("lu", "all"):
        A0 = torch.randn((n, n), dtype=torch.float64, device=dev).T
        A = A0.clone(memory_format=torch.preserve_format)
        p = torch.zeros(n, dtype=torch.int64, device=dev); pi = torch.zeros(n, dtype=torch.int64, device=dev)
        tc = time_ms(lambda: A.copy_(A0))
        def run():
            A.copy_(A0); la.lu_in_place(A, p, pi)
        l0 = lib.faer_b200_launch_count(); run(); nl = lib.faer_b200_launch_count() - l0
        ms = time_ms(run, 2) - tc
        print(f"LU  n={n}: {ms:.2f} ms  {2*n**3/3/ms/1e9:.2f} TFLOP/s  ({nl} launches)", flush=True)
        out[f"lu_{n}"] = 2*n**3/3/ms/1e9
        del A, A0
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/time_factor_{what}.json", "w"))
