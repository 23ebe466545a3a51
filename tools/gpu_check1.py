"""First-contact GPU check (development aid, not a test): numerics vs numpy + rough timings."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

lib = faer_b200.load()
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
out = {}


def colmajor(a):
    return np.asfortranarray(a)


def t_dev(a):
    """device tensor sharing the numpy array's logical layout (col-major aware)"""
    if a.flags.f_contiguous and not a.flags.c_contiguous:
        return torch.from_numpy(np.ascontiguousarray(a.T)).to(dev).T
    return torch.from_numpy(a).to(dev)


fails = 0


def check(name, got, want, tol):
    global fails
    err = float(np.max(np.abs(got - want))) if want.size else 0.0
    ok = err <= tol
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max err {err:.3e} (tol {tol:.1e})", flush=True)
    if not ok:
        fails += 1


# ---- GEMM, host arrays, all layouts ----
for (m, n, k) in [(1, 1, 1), (7, 5, 3), (64, 64, 64), (127, 129, 65), (256, 256, 256), (300, 200, 500), (1, 300, 70), (300, 1, 70)]:
    for la_, lb_, lc_ in [("F", "F", "F"), ("C", "F", "F"), ("F", "C", "F"), ("C", "C", "C"), ("F", "F", "C")]:
        A = np.array(rng.standard_normal((m, k)), order=la_)
        B = np.array(rng.standard_normal((k, n)), order=lb_)
        Cm = np.array(rng.standard_normal((m, n)), order=lc_)
        want = Cm + 0.5 * (A @ B)
        la.matmul(Cm, la.Accum.Add, A, B, 0.5)
        check(f"gemm host {m}x{n}x{k} {la_}{lb_}{lc_} add", Cm, want, 1e-11 * max(1, k))
        C2 = np.full((m, n), np.nan, order=lc_)
        la.matmul(C2, la.Accum.Replace, A, B, -1.0)
        check(f"gemm host {m}x{n}x{k} {la_}{lb_}{lc_} replace", C2, -(A @ B), 1e-11 * max(1, k))

# strided / negative stride sub-views (host)
A = rng.standard_normal((200, 300)); B = rng.standard_normal((300, 150)); Cm = np.zeros((400, 300))
Av = A[::2, ::-1][:, :100]; Bv = B[:100:1, ::3]; Cv = Cm[1::4, ::6][:100, :50]
la.matmul(Cv, la.Accum.Replace, Av, Bv, 2.0)
check("gemm host strided views", Cv, 2.0 * (Av @ Bv), 1e-10)

# ---- GEMM device ----
for n in [512, 1000]:
    A = colmajor(rng.standard_normal((n, n))); B = colmajor(rng.standard_normal((n, n)))
    dA, dB = t_dev(A), t_dev(B)
    dC = torch.empty((n, n), dtype=torch.float64, device=dev).T
    la.matmul(dC, la.Accum.Replace, dA, dB, 1.0)
    check(f"gemm dev n={n}", dC.cpu().numpy(), A @ B, 1e-10 * n)

# ---- triangular matmul ----
S = la.BlockStructure


def mask(a, s):
    a = a.copy()
    if s == S.Rectangular: return a
    low = s in (S.TriangularLower, S.StrictTriangularLower, S.UnitTriangularLower)
    a = np.tril(a) if low else np.triu(a)
    if s in (S.StrictTriangularLower, S.StrictTriangularUpper): np.fill_diagonal(a, 0.0)
    if s in (S.UnitTriangularLower, S.UnitTriangularUpper): np.fill_diagonal(a, 1.0)
    return a


for n in [37, 150, 257]:
    for ds in range(7):
        for ls in range(7):
            for rs_ in range(7):
                if (ds * 7 + ls) * 7 + rs_ not in range(0, 343, 5) and n != 37: continue
                k = n if (ls != 0 or rs_ != 0) else 45
                m = n; nn = n
                A = colmajor(rng.standard_normal((m, k))); B = colmajor(rng.standard_normal((k, nn)))
                Cm = colmajor(rng.standard_normal((m, nn))); C0 = Cm.copy()
                full = C0 + 0.7 * (mask(A, ls) @ mask(B, rs_))
                if ds == 0: want = full
                else:
                    low = ds in (1, 3, 5)
                    sel = np.tril(np.ones((m, nn), bool)) if low else np.triu(np.ones((m, nn), bool))
                    if ds >= 3: np.fill_diagonal(sel, False)
                    want = np.where(sel, full, C0)
                la.matmul_triangular(Cm, ds, la.Accum.Add, A, ls, B, rs_, 0.7)
                err = np.max(np.abs(Cm - want))
                if err > 1e-10:
                    print(f"FAIL tri n={n} dst={ds} lhs={ls} rhs={rs_} err={err:.3e}"); fails += 1
print("tri matmul sweep done", flush=True)

# ---- TRSM ----
for n, k in [(5, 3), (33, 70), (128, 300), (257, 64)]:
    T = colmajor(rng.standard_normal((n, n)) + n * np.eye(n))
    for lower in (True, False):
        for unit in (True, False):
            Tm = np.tril(T) if lower else np.triu(T)
            if unit: np.fill_diagonal(Tm, 1.0)
            Bm = colmajor(rng.standard_normal((n, k)))
            X = Bm.copy()
            f = {(True, False): la.solve_lower_triangular_in_place, (True, True): la.solve_unit_lower_triangular_in_place,
                 (False, False): la.solve_upper_triangular_in_place, (False, True): la.solve_unit_upper_triangular_in_place}[(lower, unit)]
            f(T, X)
            check(f"trsm n={n} k={k} lower={lower} unit={unit}", Tm @ X, Bm, 1e-9)

# ---- LLT ----
for n in [1, 5, 64, 100, 129, 300, 1000]:
    G = rng.standard_normal((n, n)); A = colmajor(G @ G.T + n * np.eye(n))
    L = A.copy()
    info = la.cholesky_in_place(L)
    Lt = np.tril(L)
    check(f"llt host n={n}", Lt @ Lt.T, A, 1e-12 * n * np.max(np.abs(A)))
    check(f"llt host n={n} upper untouched", np.triu(L, 1), np.triu(A, 1), 0.0)
# failure index
A = colmajor(np.eye(200)); A[150, 150] = -1.0
try:
    la.cholesky_in_place(A); print("FAIL: no error"); fails += 1
except la.LltError as e:
    print("OK   llt error index", e.index, "(want 150)"); fails += (e.index != 150)

# ---- timings (device resident) ----
def time_ms(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
for n in [4096, 8192, 16384]:
    dA = torch.randn((n, n), dtype=torch.float64, device=dev).T
    dB = torch.randn((n, n), dtype=torch.float64, device=dev).T
    dC = torch.empty((n, n), dtype=torch.float64, device=dev).T
    ms = time_ms(lambda: la.matmul(dC, la.Accum.Replace, dA, dB, 1.0), 2 if n > 8192 else 3)
    tf = 2.0 * n ** 3 / ms / 1e9
    print(f"GEMM n={n}: {ms:.2f} ms, {tf:.2f} TFLOP/s", flush=True)
    out[f"gemm_{n}_tflops"] = tf
    if n == 4096:
        ref = (dA @ dB)
        print("   vs torch err", float((dC - ref).abs().max()))
    del dA, dB, dC

for n in [4096, 16384]:
    G = torch.randn((n, n), dtype=torch.float64, device=dev)
    A0 = (G @ G.T + n * torch.eye(n, dtype=torch.float64, device=dev)).T.contiguous().T
    del G
    A = A0.clone()
    def run():
        A.copy_(A0)
        la.cholesky_in_place(A)
    t_copy = time_ms(lambda: A.copy_(A0))
    ms = time_ms(run, 2) - t_copy
    tf = n ** 3 / 3 / ms / 1e9
    print(f"LLT n={n}: {ms:.2f} ms, {tf:.2f} TFLOP/s", flush=True)
    out[f"llt_{n}_tflops"] = tf
    if n == 4096:
        L = torch.tril(A)
        print("   recon err", float((L @ L.T - A0).abs().max()), "ref scale", float(A0.abs().max()))
    del A, A0

print("launches:", lib.faer_b200_launch_count())
print("FAILS:", fails)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/check1.json", "w"))
sys.exit(1 if fails else 0)
