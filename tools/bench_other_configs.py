"""Device-resident timings of the BASELINE.json configs other than configs[1] / [2]: f32 Householder QR 65536 x 4096
(configs[3]), f64 bidiagonalization n = 8192 and c64 GEMM n = 8192 (configs[4]). Prints ONE JSON object. Run by bench.py
in a child process (informational `also.other_configs`), or by hand. Same timing as tools/time_other.py /
tools/time_condensed.py: CUDA events around restore-copy + call, best of the repetitions, copy time subtracted."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)


def best_ms(f, reps=2):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best


out = {}
torch.manual_seed(99)
m3, n3 = 65536, 4096
Q0 = torch.randn((n3, m3), dtype=torch.float32, device=dev).T
Qw = Q0.clone(memory_format=torch.preserve_format)
bs3 = la.qr_recommended_block_size(m3, n3)
Hq = torch.zeros((n3, bs3), dtype=torch.float32, device=dev).T
t_c = best_ms(lambda: Qw.copy_(Q0))


def _qr():
    Qw.copy_(Q0)
    la.qr_in_place(Qw, Hq)


t = best_ms(_qr) - t_c
out["qr_f32_65536x4096_ms"] = t
out["qr_f32_tflops"] = (2.0 * m3 * n3 * n3 - 2.0 * n3 ** 3 / 3.0) / t / 1e9
del Q0, Qw, Hq

n4 = 8192
B0 = torch.randn((n4, n4), dtype=torch.float64, device=dev)
Bw = B0.clone().T
Hl = torch.zeros((n4, 64), dtype=torch.float64, device=dev).T
Hr = torch.zeros((n4 - 1, 64), dtype=torch.float64, device=dev).T
t_c = best_ms(lambda: Bw.copy_(B0.T))


def _bd():
    Bw.copy_(B0.T)
    la.bidiag_in_place(Bw, Hl, Hr)


t = best_ms(_bd, reps=1) - t_c
out["bidiag_f64_n8192_ms"] = t
out["bidiag_algorithmic_GBps"] = 8.0 * n4 ** 3 / t / 1e6  # sum_k 3 * 8 * (n - k)^2 bytes
del B0, Bw, Hl, Hr

Ca = torch.randn((n4, n4), dtype=torch.complex128, device=dev).T
Cb = torch.randn((n4, n4), dtype=torch.complex128, device=dev).T
Cc = torch.empty((n4, n4), dtype=torch.complex128, device=dev).T
t = best_ms(lambda: la.matmul(Cc, la.Accum.Replace, Ca, Cb, 1.0))
out["gemm_c64_n8192_ms"] = t
out["gemm_c64_tflops"] = 8.0 * n4 ** 3 / t / 1e9
print(json.dumps(out), flush=True)

# configs[4] end to end for the values: bidiagonalization + bisection (csrc/svd.cu). That driver had not run on hardware
# when this was written, so the line above is already out; if this part succeeds a second, richer line supersedes it
# (bench.py reads the last one).
try:
    A8 = torch.randn((n4, n4), dtype=torch.float64, device=dev).T
    t = best_ms(lambda: la.singular_values(A8), reps=1)
    out["singular_values_f64_n8192_ms"] = t
    print(json.dumps(out), flush=True)
except Exception as e:  # pragma: no cover
    out["singular_values_f64_n8192_ms"] = None
    out["singular_values_error"] = repr(e)
    print(json.dumps(out), flush=True)

# configs[4] with vectors: svd(A, Full U, Full V) at n = 8192 (bidiagonalization + divide and conquer on the Golub-Kahan form +
# QR stabilisation + back-transforms; csrc/svd_vectors.cu), and the self-adjoint EVD with vectors at the same n. Residual of the
# factorization checked on a probe (not timed).
try:
    S8 = torch.zeros(n4, dtype=torch.float64, device=dev)
    U8 = torch.zeros((n4, n4), dtype=torch.float64, device=dev).T
    V8 = torch.zeros((n4, n4), dtype=torch.float64, device=dev).T
    t = best_ms(lambda: la.svd(A8, S8, U8, V8), reps=1)
    out["svd_full_vectors_f64_n8192_ms"] = t
    x = torch.randn((n4, 2), dtype=torch.float64, device=dev)
    r = (A8 @ x - U8 @ (S8[:, None] * (V8.T @ x))).abs().max() / (A8.abs().max() * n4)
    out["svd_probe_residual"] = float(r)
    print(json.dumps(out), flush=True)
    del U8, V8
    H8 = (A8 + A8.T).T.contiguous().T
    W8 = torch.zeros((n4, n4), dtype=torch.float64, device=dev).T
    t = best_ms(lambda: la.self_adjoint_evd(H8, S8, W8), reps=1)
    out["self_adjoint_evd_vectors_f64_n8192_ms"] = t
    out["evd_probe_residual"] = float((H8 @ x - W8 @ (S8[:, None] * (W8.T @ x))).abs().max() / (H8.abs().max() * n4))
    print(json.dumps(out), flush=True)
except Exception as e:  # pragma: no cover
    out["svd_vectors_error"] = repr(e)
    print(json.dumps(out), flush=True)
