"""Next-round groundwork (CPU only): numerical model of the int8-sliced ("Ozaki") f64 GEMM planned for tcgen05 kind::i8.

Scheme (row/column block-floating-point, signed 7-bit slices, exact int32 accumulation per order):
  a_ij = 2^e_i * x_ij, |x| < 1;  x = A0/64 + A1/(64*128) + A2/(64*128^2) + ...   with A_p = rint(.) in [-64, 64] (int8)
  b_kj likewise with per-column exponents f_j.
  C_ij = 2^(e_i + f_j) * sum_{d=0}^{S-1} 2^(-12 - 7 d) * G_d[i, j],   G_d = sum_{p+q=d} A_p B_q   (integer, exact)
Pairs with p + q >= S are dropped (S(S+1)/2 int8 GEMMs instead of S^2). This script measures the error of the scheme against
an exact (integer / long double) product for several S, matrix kinds and k, next to the error of a plain f64 GEMM, so that the
CUDA implementation (faer-rs_b200/csrc/gemm_f64_ozaki.cuh, tools/next/ozaki_test.cu) has a reference to be checked against.
usage: python tools/next/ozaki_emulation.py
"""
import numpy as np


def slices(X, axis, S):
    """Split X (f64) into S int8 slice matrices and the per-row (axis=1) / per-column (axis=0) exponents."""
    amax = np.max(np.abs(X), axis=axis, keepdims=True)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0).astype(np.int64)
    x = np.ldexp(X, -e)  # |x| < 1, exact
    out = []
    r = x * 64.0
    for p in range(S):
        a = np.rint(r)
        out.append(a.astype(np.int8))
        r = (r - a) * 128.0  # exact: both are multiples of 2^-46.. within f64
    return out, e


def ozaki_gemm(A, B, S):
    As, e = slices(A, 1, S)
    Bs, f = slices(B, 0, S)
    m, n = A.shape[0], B.shape[1]
    C = np.zeros((m, n))
    for d in range(S - 1, -1, -1):  # small terms first
        G = np.zeros((m, n), dtype=np.int64)
        for p in range(d + 1):
            G += As[p].astype(np.int64) @ Bs[d - p].astype(np.int64)
        assert np.abs(G).max() < 2 ** 31, "int32 accumulator would overflow"
        C += np.ldexp(G.astype(np.float64), -12 - 7 * d)
    return np.ldexp(C, e + f)


def exact(A, B):
    return (A.astype(np.longdouble) @ B.astype(np.longdouble))


def report(name, A, B):
    ref = exact(A, B)
    scale = (np.abs(A).astype(np.longdouble) @ np.abs(B).astype(np.longdouble))  # the |a||b| bound of SURVEY appendix B
    e64 = float(np.max(np.abs(A @ B - ref) / scale))
    row = [f"{name:34s} k={A.shape[1]:5d}  f64 GEMM {e64:.2e}"]
    for S in (6, 7, 8, 9):
        err = float(np.max(np.abs(ozaki_gemm(A, B, S) - ref) / scale))
        row.append(f"S={S}: {err:.2e}")
    print("  ".join(row), flush=True)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    u = 2.0 ** -53
    print(f"errors are max |C - exact| / (|A||B|)_ij; u = {u:.2e}")
    for k in (64, 1024, 4096):
        m = n = 96
        report("gaussian", rng.standard_normal((m, k)), rng.standard_normal((k, n)))
        report("uniform(0,1) (no cancellation)", rng.uniform(0, 1, (m, k)), rng.uniform(0, 1, (k, n)))
        A = rng.standard_normal((m, k)) * np.exp(rng.uniform(-8, 8, (m, 1)))          # row scaling: absorbed by e_i
        B = rng.standard_normal((k, n)) * np.exp(rng.uniform(-8, 8, (1, n)))
        report("row/col scaled by e^[-8,8]", A, B)
        A = rng.standard_normal((m, k)) * np.exp(rng.uniform(-6, 6, (m, k)))          # wide range INSIDE rows: small entries lose bits
        B = rng.standard_normal((k, n))
        report("entrywise range e^[-6,6] in A", A, B)
