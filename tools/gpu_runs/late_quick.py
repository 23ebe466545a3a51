"""A few seconds of hardware for the late additions (no torch import: numpy + ctypes only): one tiny call per new family, each
checked with numpy, progress flushed line by line so that a partial run still tells which families ran.
   gpurun --timeout 60 -- 'timeout 50 python tools/gpu_runs/late_quick.py > gpurun_out/late_quick.log 2>&1'"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
t0 = time.time()
import faer_b200  # noqa: E402

la = faer_b200.linalg
res = {}
rng = np.random.default_rng(0)


def crandn(shape, dtype=np.complex128):
    return np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype))


def step(name, f):
    t = time.time()
    try:
        err = float(f())
        res[name] = {"err": err, "s": round(time.time() - t, 3)}
    except Exception as e:  # noqa: BLE001
        res[name] = {"exception": repr(e)}
    print(name, res[name], flush=True)


def inv_tri():
    n = 40
    T = np.asfortranarray(np.tril(rng.standard_normal((n, n)) / 6 + 2 * np.eye(n)).astype(np.float32))
    out = np.zeros((n, n), np.float32, order="F")
    la.invert_triangular(out, T, True, False)
    return np.abs(np.tril(out).astype(np.float64) @ T.astype(np.float64) - np.eye(n)).max()


def llt_rec():
    n = 50
    G = crandn((n, n)); A = np.asfortranarray(G @ G.conj().T + n * np.eye(n))
    L = A.copy(order="F"); la.cholesky_in_place(L)
    out = np.zeros((n, n), np.complex128, order="F"); la.llt_reconstruct(out, L)
    return np.abs(np.tril(out) - np.tril(A)).max() / np.abs(A).max()


def ldlt_f32():
    n = 60
    G = rng.standard_normal((n, n)); s = np.where(rng.random(n) < 0.4, -1.0, 1.0)
    A = np.asfortranarray(((G + G.T) / np.sqrt(n) + np.diag(4 * s)).astype(np.float32))
    LD = A.copy(order="F"); la.ldlt_in_place(LD)
    out = np.zeros((n, n), np.float32, order="F"); la.ldlt_reconstruct(out, LD)
    return np.abs(np.tril(out) - np.tril(A)).max()


def hess():
    n = 30
    A = np.asfortranarray(rng.standard_normal((n, n)))
    W = A.copy(order="F"); H = np.zeros((4, n - 1), order="F"); la.hessenberg_in_place(W, H)
    ev_a = np.sort(np.abs(np.linalg.eigvals(A))); ev_h = np.sort(np.abs(np.linalg.eigvals(np.triu(W, -1))))
    return np.abs(ev_a - ev_h).max()


def svd_c():
    A = crandn((20, 12)); S = np.zeros(12, np.complex128)
    U = np.zeros((20, 12), np.complex128, order="F"); V = np.zeros((12, 12), np.complex128, order="F")
    la.svd(A, S, U, V)
    return np.abs((U * S.real[None, :]) @ V.conj().T - A).max()


def evd_c():
    n = 16
    G = crandn((n, n)); A = np.asfortranarray(G + G.conj().T)
    S = np.zeros(n, np.complex128); U = np.zeros((n, n), np.complex128, order="F")
    la.self_adjoint_evd(A, S, U)
    return np.abs((U * S.real[None, :]) @ U.conj().T - A).max()


def tridiag_rm():
    n = 24
    G = rng.standard_normal((n, n)); A = (G + G.T) / 2
    ref = np.asfortranarray(A); H0 = np.zeros((4, n - 1), order="F"); la.tridiag_in_place(ref, H0)
    rm = np.array(A, order="C", copy=True); H = np.zeros((4, n - 1), order="F"); la.tridiag_in_place(rm, H)
    return np.abs(np.tril(rm) - np.tril(ref)).max()


print("import", round(time.time() - t0, 2), "s", flush=True)
for name, f in (("inverse_triangular_f32", inv_tri), ("llt_reconstruct_c64", llt_rec), ("ldlt_f32", ldlt_f32), ("hessenberg_f64", hess),
                ("svd_c64", svd_c), ("self_adjoint_evd_c64", evd_c), ("tridiag_row_major_f64", tridiag_rm)):
    step(name, f)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/late_quick.json", "w"), indent=1)
print("done", round(time.time() - t0, 2), "s", flush=True)
