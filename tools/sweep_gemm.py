"""Dev aid: time the f64 GEMM for each forced tile configuration (FAER_B200_GEMM_CFG) in a subprocess."""
import os, subprocess, sys
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import faer_b200
from faer_b200 import linalg as la
lib = faer_b200.load(); dev = torch.device("cuda:0")
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
for n in [4096, 8192]:
    A = torch.randn((n, n), dtype=torch.float64, device=dev).T
    B = torch.randn((n, n), dtype=torch.float64, device=dev).T
    C = torch.empty((n, n), dtype=torch.float64, device=dev).T
    for kind in ["NN", "NT", "TN"]:
        a = A if kind[0] == "N" else A.T
        b = B if kind[1] == "N" else B.T
        la.matmul(C, 0, a, b, 1.0); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); la.matmul(C, 0, a, b, 1.0); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"cfg={os.environ.get('FAER_B200_GEMM_CFG')} n={n} {kind}: {best:.2f} ms {2*n**3/best/1e9:.2f} TF", flush=True)
    # SYRK-like lower update with k = 128 / 512
    for k in [128, 512]:
        P = torch.randn((k, n), dtype=torch.float64, device=dev).T  # n x k col-major
        la.matmul_triangular(C, 1, 1, P, 0, P.T, 0, -1.0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); la.matmul_triangular(C, 1, 1, P, 0, P.T, 0, -1.0); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"cfg={os.environ.get('FAER_B200_GEMM_CFG')} syrk n={n} k={k}: {ms:.3f} ms {n*n*k/ms/1e9:.2f} TF", flush=True)
'''
for cfg in ["5", "10"]:
    env = dict(os.environ, FAER_B200_GEMM_CFG=cfg)
    subprocess.run([sys.executable, "-c", code], env=env)
