"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes binding of the CPU oracle (oracle/liboracle.so, built by oracle/Makefile). Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The oracle restates the reference's algorithms (see oracle.hpp for file:line citations and for the
"parity unpinned at the bit level, pinned at tolerance level" statement).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

_DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}
RECT, TRI_LOWER, TRI_UPPER, STRICT_LOWER, STRICT_UPPER, UNIT_LOWER, UNIT_UPPER = range(7)


class OMat(C.Structure):
    _fields_ = [("p", C.c_void_p), ("m", C.c_longlong), ("n", C.c_longlong), ("rs", C.c_longlong), ("cs", C.c_longlong)]


_lib = None


def build() -> None:
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_set_num_threads.argtypes = [C.c_int]
    lib.oracle_matmul.argtypes = [C.c_int, OMat, C.c_int, OMat, C.c_int, OMat, C.c_int, C.c_void_p]
    lib.oracle_matmul.restype = C.c_longlong
    lib.oracle_matmul_triangular.argtypes = [C.c_int, OMat, C.c_int, C.c_int, OMat, C.c_int, C.c_int, OMat, C.c_int,
                                             C.c_int, C.c_void_p]
    lib.oracle_matmul_triangular.restype = C.c_longlong
    lib.oracle_solve_triangular.argtypes = [C.c_int, C.c_int, C.c_int, OMat, C.c_int, OMat]
    lib.oracle_solve_triangular.restype = C.c_longlong
    lib.oracle_llt.argtypes = [C.c_int, OMat, C.c_double, C.c_double, C.c_longlong, C.c_longlong,
                               C.POINTER(C.c_longlong)]
    lib.oracle_llt.restype = C.c_longlong
    lib.oracle_ldlt.argtypes = [C.c_int, OMat, C.c_double, C.c_double, C.c_void_p, C.c_longlong, C.c_longlong,
                                C.POINTER(C.c_longlong)]
    lib.oracle_ldlt.restype = C.c_longlong
    lib.oracle_lu.argtypes = [C.c_int, OMat, C.c_void_p, C.c_void_p, C.c_longlong]
    lib.oracle_lu.restype = C.c_longlong
    lib.oracle_qr.argtypes = [C.c_int, OMat, OMat, C.c_longlong]
    lib.oracle_qr.restype = C.c_longlong
    lib.oracle_qr_recommended_block_size.argtypes = [C.c_longlong, C.c_longlong]
    lib.oracle_qr_recommended_block_size.restype = C.c_longlong
    lib.oracle_bidiag.argtypes = [C.c_int, OMat, OMat, OMat]
    lib.oracle_bidiag.restype = C.c_longlong
    lib.oracle_tridiag.argtypes = [C.c_int, OMat, OMat]
    lib.oracle_tridiag.restype = C.c_longlong
    lib.oracle_apply_block_householder_left.argtypes = [C.c_int, OMat, OMat, C.c_int, OMat, C.c_int]
    lib.oracle_apply_block_householder_left.restype = C.c_longlong
    lib.oracle_norm_l2.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_longlong]
    lib.oracle_norm_l2.restype = C.c_double
    _lib = lib
    return lib


def _om(a: np.ndarray) -> OMat:
    assert a.ndim == 2
    return OMat(a.ctypes.data, a.shape[0], a.shape[1], a.strides[0] // a.itemsize, a.strides[1] // a.itemsize)


def _scalar(dtype, v):
    return np.array([v], dtype=dtype)


def num_threads() -> int:
    return load().oracle_num_threads()


def set_num_threads(n: int) -> None:
    load().oracle_set_num_threads(int(n))


def matmul(dst, add: bool, lhs, rhs, alpha, conj_lhs=False, conj_rhs=False) -> None:
    """dst = [dst +] alpha * conj?(lhs) * conj?(rhs)   (matmul/mod.rs:1909-1947)"""
    assert dst.shape == (lhs.shape[0], rhs.shape[1]) and lhs.shape[1] == rhs.shape[0]
    assert dst.dtype == lhs.dtype == rhs.dtype
    a = _scalar(dst.dtype, alpha)
    r = load().oracle_matmul(_DT[dst.dtype], _om(dst), int(add), _om(lhs), int(conj_lhs), _om(rhs), int(conj_rhs),
                             a.ctypes.data)
    assert r == 0


def matmul_triangular(dst, dst_s, add, lhs, lhs_s, rhs, rhs_s, alpha, conj_lhs=False, conj_rhs=False) -> None:
    assert dst.shape == (lhs.shape[0], rhs.shape[1]) and lhs.shape[1] == rhs.shape[0]
    a = _scalar(dst.dtype, alpha)
    r = load().oracle_matmul_triangular(_DT[dst.dtype], _om(dst), dst_s, int(add), _om(lhs), lhs_s, int(conj_lhs),
                                        _om(rhs), rhs_s, int(conj_rhs), a.ctypes.data)
    assert r == 0


def solve_triangular(tri, rhs, lower: bool, unit: bool, conj=False) -> None:
    assert tri.shape[0] == tri.shape[1] == rhs.shape[0]
    r = load().oracle_solve_triangular(_DT[rhs.dtype], int(lower), int(unit), _om(tri), int(conj), _om(rhs))
    assert r == 0


def llt(A, delta=0.0, eps=0.0, recursion_threshold=64, block_size=128):
    """In-place LLT of the lower triangle. Returns (fail_index or -1, regularisation count)."""
    assert A.shape[0] == A.shape[1]
    cnt = C.c_longlong(0)
    r = load().oracle_llt(_DT[A.dtype], _om(A), float(delta), float(eps), recursion_threshold, block_size, C.byref(cnt))
    assert r != -100
    return int(r), int(cnt.value)


def ldlt(A, delta=0.0, eps=0.0, signs=None, recursion_threshold=64, block_size=128):
    """In-place LDLT of the lower triangle (cholesky/ldlt/factor.rs:725-767): D on the diagonal, unit-lower L strictly
    below it. Returns (ZeroPivot index or -1, regularisation count). signs: optional int8 array of expected pivot signs."""
    assert A.shape[0] == A.shape[1]
    cnt = C.c_longlong(0)
    sp = None
    if signs is not None:
        signs = np.ascontiguousarray(signs, dtype=np.int8)
        assert signs.size == A.shape[0]
        sp = signs.ctypes.data
    r = load().oracle_ldlt(_DT[A.dtype], _om(A), float(delta), float(eps), sp, recursion_threshold, block_size,
                           C.byref(cnt))
    assert r != -100
    return int(r), int(cnt.value)


def ldlt_solve(LD, rhs, conj_lhs=False) -> None:
    """cholesky::ldlt::solve::solve_in_place_with_conj (ldlt/solve.rs:11-49): unit-lower solve with L, rows scaled by
    recip(Re d_i), unit-upper solve with L^T under conj composed with Yes."""
    n = LD.shape[0]
    assert LD.shape[1] == n and rhs.shape[0] == n
    solve_triangular(LD, rhs, lower=True, unit=True, conj=conj_lhs)
    d = (1.0 / np.real(np.diagonal(LD))).astype(np.real(LD[:1, :1]).dtype)
    rhs *= d[:, None]
    solve_triangular(LD.T, rhs, lower=False, unit=True, conj=not conj_lhs)


def lu(A, recursion_threshold=16):
    """In-place P A = L U. Returns (perm, perm_inv, transposition_count); (P A)[i] = A[perm[i]]."""
    m = A.shape[0]
    perm = np.zeros(m, dtype=np.int64)
    perm_inv = np.zeros(m, dtype=np.int64)
    r = load().oracle_lu(_DT[A.dtype], _om(A), perm.ctypes.data, perm_inv.ctypes.data, recursion_threshold)
    assert r != -100
    return perm, perm_inv, int(r)


def qr_recommended_block_size(nrows: int, ncols: int) -> int:
    """qr/no_pivoting/factor.rs:91-116"""
    return int(load().oracle_qr_recommended_block_size(nrows, ncols))


def qr(A, block_size=None, blocking_threshold=48 * 48):
    """In-place Householder QR without pivoting (qr/no_pivoting/factor.rs:258-301).
    Returns (Q_coeff [block_size x min(m,n)], rank). R is the upper triangle of A, V strictly below the diagonal."""
    m, n = A.shape
    bs = block_size or qr_recommended_block_size(m, n)
    H = np.zeros((bs, min(m, n)), dtype=A.dtype, order="F")
    r = load().oracle_qr(_DT[A.dtype], _om(A), _om(H), blocking_threshold)
    assert r != -100
    return H, int(r)


def apply_block_householder_on_the_left(V, T, M, forward: bool, conj_lhs=False) -> None:
    """M <- (I - V T^-1 V^H) M (forward=False) or (I - V T^-H V^H) M (forward=True)  (householder.rs:370-620)."""
    r = load().oracle_apply_block_householder_left(_DT[M.dtype], _om(V), _om(T), int(conj_lhs), _om(M), int(forward))
    assert r == 0


def apply_q_transpose_sequence(QR, H, M, conj_lhs=True) -> None:
    """apply_block_householder_sequence_transpose_on_the_left_in_place_with_conj (householder.rs:768-808): M <- Q^H M."""
    bs, size = H.shape
    j = 0
    while j < size:
        b = min(bs, size - j)
        # transpose_on_the_left composes conj_lhs with Conj::Yes and uses forward = true (householder.rs:697-719)
        apply_block_householder_on_the_left(QR[j:, j:j + b], H[:b, j:j + b], M[j:, :], True, conj_lhs=not conj_lhs)
        j += b


def apply_q_sequence(QR, H, M, conj_lhs=False) -> None:
    """apply_block_householder_sequence_on_the_left_in_place_with_conj (householder.rs:724-765): M <- Q M."""
    bs, size = H.shape
    j = size
    b = size % bs or bs
    while j > 0:
        jp = j - b
        apply_block_householder_on_the_left(QR[jp:, jp:j], H[:j - jp, jp:j], M[jp:, :], False, conj_lhs=conj_lhs)
        j = jp
        b = bs


def qr_solve_lstsq(QR, H, rhs, conj_QR=False) -> None:
    """qr::no_pivoting::solve::solve_lstsq_in_place_with_conj (qr/no_pivoting/solve.rs:38-76): rhs <- op(Q)^H rhs (the
    transpose sequence with conj_QR composed with Yes), then the upper solve with op(R)[..size, ..] on rhs[..size, ..];
    op = conj if conj_QR. Q_basis and R are both the packed QR matrix, as in the reference's own test
    (solve.rs:232-246)."""
    m, n = QR.shape
    size = min(m, n)
    assert m >= n and rhs.shape[0] == m and H.shape[1] == size
    apply_q_transpose_sequence(QR, H, rhs, conj_lhs=not conj_QR)
    solve_triangular(QR[:size, :n], rhs[:size, :], lower=False, unit=False, conj=conj_QR)


def qr_solve(QR, H, rhs, conj_QR=False) -> None:
    """solve_in_place_with_conj (solve.rs:96-119): the square case of the above."""
    assert QR.shape[0] == QR.shape[1]
    qr_solve_lstsq(QR, H, rhs, conj_QR)


def qr_solve_transpose(QR, H, rhs, conj_QR=False) -> None:
    """solve_transpose_in_place_with_conj (solve.rs:140-176): lower solve with op(R)^T, then the forward sequence with
    conj_QR composed with Yes, i.e. rhs <- op(A)^-T rhs."""
    n = QR.shape[0]
    assert QR.shape[1] == n and rhs.shape[0] == n
    solve_triangular(QR.T, rhs, lower=True, unit=False, conj=conj_QR)
    apply_q_sequence(QR, H, rhs, conj_lhs=not conj_QR)


def bidiag(A, bl: int, br: int):
    """In-place bidiagonalization A = U B V^H, m >= n (svd/bidiag.rs:47-256). Returns (H_left [bl x n], H_right
    [br x (n-1)]). B is on A's diagonal / superdiagonal, left reflectors below the diagonal, right reflectors to the
    right of the superdiagonal."""
    m, n = A.shape
    size = min(m, n)
    Hl = np.zeros((bl, size), dtype=A.dtype, order="F")
    Hr = np.zeros((br, max(size - 1, 0)), dtype=A.dtype, order="F")
    r = load().oracle_bidiag(_DT[A.dtype], _om(A), _om(Hl), _om(Hr))
    assert r == 0
    return Hl, Hr


def tridiag(A, b: int):
    """In-place tridiagonalization A = Q T Q^H of a self-adjoint matrix, lower triangle only (evd/tridiag.rs:274-529).
    Returns H [b x (n-1)]. T is on A's diagonal / subdiagonal, reflectors below the subdiagonal."""
    n = A.shape[0]
    H = np.zeros((b, max(n - 1, 0)), dtype=A.dtype, order="F")
    r = load().oracle_tridiag(_DT[A.dtype], _om(A), _om(H))
    assert r == 0
    return H


def norm_l2(x) -> float:
    x = np.ascontiguousarray(x).ravel()
    return float(load().oracle_norm_l2(_DT[x.dtype], x.ctypes.data, x.size, 1))
