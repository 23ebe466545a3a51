"""The generic LDLT of csrc/ldlt_core.cuh (the f32 / c64 / c32 factorization on the GPU: an unblocked right-looking launch sequence
of flat maps) compiled for the host (tools/emul/ldlt_host.cpp) and run thread by thread, against the oracle's restatement of the
reference (cholesky/ldlt/factor.rs:299-498, 725-767): same L and D up to rounding, failure index and regularisation count exact,
strict upper triangle untouched, forward and reverse thread order bit-identical, strided storage; and the bodies of the solve /
reconstruct / inverse compositions (ldlt/solve.rs, reconstruct.rs, inverse.rs)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I64 = C.c_longlong
P = C.c_void_p
KIND = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}
DTYPES = [np.float32, np.float64, np.complex64, np.complex128]


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ldle") / "libldle.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, os.path.join(ROOT, "tools", "emul", "ldlt_host.cpp")])
    lib = C.CDLL(out)
    lib.ldlt_emul_factor.argtypes = [C.c_int, P, I64, I64, I64, P, P, C.c_double, C.c_double, P, C.c_int]
    lib.ldlt_emul_body.argtypes = [C.c_int, C.c_int, P, I64, I64, I64, I64, P, I64, P, C.c_int]
    return lib


def rdt(dtype):
    return np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64


def rand_sym(rng, n, dtype):
    G = rng.standard_normal((n, n))
    if np.issubdtype(dtype, np.complexfloating):
        G = G + 1j * rng.standard_normal((n, n))
    # indefinite but safely factorable without pivoting: strictly diagonally dominant with mixed signs
    A = (G + G.conj().T) / 2
    sgn = np.where(np.arange(n) % 3 == 0, -1.0, 1.0)
    A[np.diag_indices(n)] = sgn * (np.abs(A).sum(axis=1) + 1.0)
    return np.asfortranarray(A.astype(dtype)), sgn.astype(np.int8)


def factor(lib, A, signs=None, delta=0.0, eps=0.0, reverse=0):
    n = A.shape[0]
    es = A.itemsize
    D = np.zeros(max(n, 1), dtype=rdt(A.dtype))
    info = np.zeros(2, dtype=np.int64)
    sp = None if signs is None else np.ascontiguousarray(signs, dtype=np.int8).ctypes.data
    lib.ldlt_emul_factor(KIND[A.dtype], A.ctypes.data, A.strides[0] // es, A.strides[1] // es, n, D.ctypes.data, sp, delta, eps,
                         info.ctypes.data, reverse)
    return int(info[0]), int(info[1]), D[:n]


@pytest.mark.parametrize("dtype", DTYPES)
def test_factor_matches_the_oracle(emul, oracle, dtype):
    rng = np.random.default_rng(1600)
    u = np.finfo(rdt(dtype)).eps
    for n in [1, 2, 3, 17, 64, 65, 150]:
        A, _ = rand_sym(rng, n, dtype)
        want = A.copy(order="F")
        fail_o, cnt_o = oracle.ldlt(want)
        assert fail_o == -1 and cnt_o == 0
        got = A.copy(order="F")
        got[np.triu_indices(n, 1)] = 99                                  # the strict upper triangle is neither read nor written
        fail, cnt, D = factor(emul, got)
        rev = A.copy(order="F"); rev[np.triu_indices(n, 1)] = 99
        assert factor(emul, rev, reverse=1)[:2] == (fail, cnt) and np.array_equal(rev, got)
        assert (fail, cnt) == (-1, 0)
        assert np.all(got[np.triu_indices(n, 1)] == 99)
        tol = 64 * n * u
        assert np.abs(np.tril(got, -1) - np.tril(want, -1)).max(initial=0.0) <= tol * max(1.0, np.abs(np.tril(want, -1)).max(initial=0.0))
        assert np.abs(np.diag(got) - np.diag(want)).max() <= tol * np.abs(np.diag(want)).max()
        assert np.array_equal(np.diag(got).real.astype(D.dtype), D) and np.all(np.diag(got).imag == 0)
        # strided (row-major) storage: same values
        rm = np.ascontiguousarray(A)
        f2 = factor(emul, rm)
        assert f2[:2] == (fail, cnt) and np.array_equal(np.tril(rm), np.tril(got))


@pytest.mark.parametrize("dtype", DTYPES)
def test_zero_pivot_index_and_regularisation_count(emul, oracle, dtype):
    rng = np.random.default_rng(1601)
    n = 40
    A, sgn = rand_sym(rng, n, dtype)
    # an exactly singular leading block: the Schur complement pivot of column 7 is exactly zero
    Z = A.copy(order="F"); Z[7:, 7] = 0; Z[7, 7:] = 0; Z[7, :7] = 0; Z[7:, :7][0] = 0
    want = Z.copy(order="F"); fail_o, _ = oracle.ldlt(want)
    got = Z.copy(order="F"); fail, cnt, D = factor(emul, got)
    assert fail_o == 7 and fail == 7
    assert np.allclose(np.diag(got)[:8], np.diag(want)[:8], rtol=1e-4 if rdt(dtype) == np.float32 else 1e-12, atol=0)
    # dynamic regularisation with expected signs: tiny pivots of the wrong size are replaced, only the sign = +1 ones are counted
    R = A.copy(order="F")
    for j in (5, 11, 12, 30):
        R[j:, j] = 0; R[j, :j] = 0
        R[j, j] = 1e-30 * (1 if sgn[j] > 0 else -1)
    for signs in (sgn, None):
        want = R.copy(order="F"); fail_o, cnt_o = oracle.ldlt(want, delta=1e-3, eps=1e-8, signs=signs)
        got = R.copy(order="F"); fail, cnt, D = factor(emul, got, signs=signs, delta=1e-3, eps=1e-8)
        assert (fail, cnt) == (fail_o, cnt_o) and fail == -1
        assert np.allclose(np.tril(got), np.tril(want), rtol=1e-3 if rdt(dtype) == np.float32 else 1e-10, atol=1e-6 if rdt(dtype) == np.float32 else 1e-12)
    # a NaN pivot is a ZeroPivot too
    Nn = A.copy(order="F"); Nn[3, 3] = np.nan
    assert factor(emul, Nn)[0] == 3


@pytest.mark.parametrize("dtype", DTYPES)
def test_composition_bodies(emul, dtype):
    rng = np.random.default_rng(1602)
    n, k = 9, 4
    kind = KIND[np.dtype(dtype)]
    cx = np.issubdtype(dtype, np.complexfloating)
    r = rdt(dtype)
    A, _ = rand_sym(rng, n, dtype)
    es = A.itemsize
    for rev in (0, 1):
        # RecipDiag on the diagonal of A (stride rs + cs)
        dinv = np.zeros(n, dtype=r)
        emul.ldlt_emul_body(kind, 0, None, 0, 0, n, 0, A.ctypes.data, n + 1, dinv.ctypes.data, rev)
        assert np.array_equal(dinv, (1 / np.diag(A).real).astype(r))
        # ScaleRows on a row-major right-hand side
        X = np.ascontiguousarray((rng.standard_normal((n, k)) + (1j * rng.standard_normal((n, k)) if cx else 0)).astype(dtype))
        want = (X * dinv[:, None]).astype(dtype)
        emul.ldlt_emul_body(kind, 1, X.ctypes.data, X.strides[0] // es, X.strides[1] // es, n, k, dinv.ctypes.data, 1, None, rev)
        assert np.array_equal(X, want)
        # BuildLxD
        out = np.full((n, n), np.nan, dtype=dtype, order="F")
        emul.ldlt_emul_body(kind, 2, A.ctypes.data, 1, n, n, 0, A.ctypes.data, n + 1, out.ctypes.data, rev)
        d = np.diag(A).real.astype(r)
        want = (np.tril(A, -1) * d[None, :] + np.diag(d)).astype(dtype)
        assert np.array_equal(out, want)
        # SetIdentity
        out = np.full((n, n), np.nan, dtype=dtype, order="F")
        emul.ldlt_emul_body(kind, 3, None, 0, 0, n, 0, None, 0, out.ctypes.data, rev)
        assert np.array_equal(out, np.eye(n, dtype=dtype))
        # FillUpperAdjoint on a unit-lower M
        M = np.asfortranarray((np.tril(A, -1) + np.eye(n)).astype(dtype))
        M[np.triu_indices(n, 1)] = 55
        emul.ldlt_emul_body(kind, 4, M.ctypes.data, 1, n, n, 0, dinv.ctypes.data, 1, None, rev)
        Lm = np.tril(A, -1)
        want = Lm + np.diag(dinv).astype(dtype) + (Lm.conj().T * dinv[None, :]).astype(dtype)
        assert np.allclose(M, want, rtol=4 * np.finfo(r).eps, atol=0)


# ---- the compositions of csrc/ldlt_types.cu on the oracle's blocks, with the CUDA file's structure codes and conjugation flags ----
TRI_LOWER, TRI_UPPER, UNIT_LOWER, UNIT_UPPER = 1, 2, 5, 6


def _solve(orc, LD, rhs, conj):
    n = LD.shape[0]
    dinv = (1 / np.diag(LD).real).astype(rdt(LD.dtype))
    orc.solve_triangular(LD, rhs, lower=True, unit=True, conj=conj)
    rhs *= dinv[:, None]
    orc.solve_triangular(LD.T, rhs, lower=False, unit=True, conj=not conj)


def _reconstruct(orc, out, LD):
    n = LD.shape[0]
    d = np.diag(LD).real.astype(rdt(LD.dtype))
    LxD = np.asfortranarray((np.tril(LD, -1) * d[None, :] + np.diag(d)).astype(LD.dtype))
    orc.matmul_triangular(out, TRI_LOWER, False, LxD, TRI_LOWER, LD.T, UNIT_UPPER, 1.0, conj_lhs=False, conj_rhs=True)


def _inverse(orc, out, LD):
    n = LD.shape[0]
    dinv = (1 / np.diag(LD).real).astype(rdt(LD.dtype))
    M = np.asfortranarray(np.eye(n, dtype=LD.dtype))
    orc.solve_triangular(LD, M, lower=True, unit=True, conj=False)
    Ml = np.tril(M, -1)
    M = np.asfortranarray((Ml + np.diag(dinv) + Ml.conj().T * dinv[None, :]).astype(LD.dtype))   # FillUpperAdjoint
    orc.matmul_triangular(out, TRI_LOWER, False, M, TRI_UPPER, M, UNIT_LOWER, 1.0, conj_lhs=False, conj_rhs=False)


@pytest.mark.parametrize("dtype", DTYPES)
def test_ldlt_compositions_on_oracle_blocks(oracle, dtype):
    """ldlt/solve.rs tests (n = 50, k = 3, both conjugations), reconstruct.rs / inverse.rs tests (n = 50)."""
    rng = np.random.default_rng(1603)
    u = np.finfo(rdt(dtype)).eps
    wide = lambda x: x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
    for n in [1, 50, 97]:
        A, _ = rand_sym(rng, n, dtype)
        LD = A.copy(order="F")
        assert oracle.ldlt(LD)[0] == -1
        Aw = wide(A)
        for conj in (False, True):
            B = np.asfortranarray((rng.standard_normal((n, 3)) + (1j * rng.standard_normal((n, 3)) if np.iscomplexobj(A) else 0)).astype(dtype))
            X = B.copy(order="F"); _solve(oracle, LD, X, conj)
            Ae = Aw.conj() if conj else Aw
            assert np.abs(Ae @ wide(X) - wide(B)).max() <= 256 * n * u * np.linalg.cond(Aw) * np.abs(B).max()
        out = np.full((n, n), np.nan, dtype=dtype, order="F"); out[np.tril_indices(n)] = 7
        _reconstruct(oracle, out, LD)
        assert np.all(np.isnan(out[np.triu_indices(n, 1)]))
        assert np.abs(np.tril(wide(out)) - np.tril(Aw)).max() <= 256 * n * u * np.abs(A).max()
        inv = np.full((n, n), np.nan, dtype=dtype, order="F"); inv[np.tril_indices(n)] = 7
        _inverse(oracle, inv, LD)
        assert np.all(np.isnan(inv[np.triu_indices(n, 1)]))
        lo = np.tril(wide(inv)); full = lo + np.tril(lo, -1).conj().T
        assert np.abs(full @ Aw - np.eye(n)).max() <= 256 * n * u * np.linalg.cond(Aw)
