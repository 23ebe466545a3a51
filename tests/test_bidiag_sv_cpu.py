"""The singular-value bisection the GPU SVD driver runs per thread (csrc/bidiag_sv.cuh) is plain host/device code; this
test compiles the same header with g++ and checks it against LAPACK on bidiagonal matrices: random, graded (high relative
accuracy of the small values), with exact zeros, with clusters, n = 1, and the bidiagonal the oracle's bidiagonalization
produces from a dense matrix (end-to-end singular values against numpy)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bsv(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bsv") / "libbsv.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out,
                           os.path.join(ROOT, "tools", "emul", "bidiag_sv_host.cpp")])
    lib = C.CDLL(out)
    for name in ('bsv_f64', 'bsv_f32'):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]  # 64-bit pointers
        getattr(lib, name).restype = None

    def run(d, e):
        dt = d.dtype
        n = d.size
        d = np.ascontiguousarray(d); e = np.ascontiguousarray(e.astype(dt)) if n > 1 else np.zeros(1, dt)
        s = np.zeros(n, dt)
        f = lib.bsv_f64 if dt == np.float64 else lib.bsv_f32
        f(d.ctypes.data, e.ctypes.data, n, s.ctypes.data)
        return s
    return run


def _ref(d, e):
    B = np.diag(d.astype(np.float64)) + (np.diag(e.astype(np.float64), 1) if d.size > 1 else 0)
    return np.linalg.svd(B, compute_uv=False)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_random_and_special_bidiagonals(bsv, dtype):
    rng = np.random.default_rng(101)
    u = np.finfo(dtype).eps
    for n in [1, 2, 3, 10, 64, 257, 600]:
        d = rng.standard_normal(n).astype(dtype); e = rng.standard_normal(max(n - 1, 0)).astype(dtype)
        s = bsv(d, e); ref = _ref(d, e)
        assert np.all(np.diff(s) <= 0) and np.all(s >= 0)
        assert np.abs(s - ref).max() <= 8 * n * u * ref.max(), n
    # exact zeros and a cluster
    d = np.array([1.0, 0.0, 2.0, 2.0, 2.0, 0.0], dtype); e = np.array([0.0, 0.0, 0.0, 0.0, 0.0], dtype)
    assert np.allclose(bsv(d, e), [2, 2, 2, 1, 0, 0], atol=8 * u)
    d = np.ones(50, dtype); e = np.full(49, 1e-3, dtype)
    assert np.abs(bsv(d, e) - _ref(d, e)).max() <= 64 * u


def test_graded_matrix_high_relative_accuracy(bsv):
    # a graded bidiagonal: singular values span 1e0 .. 1e-60; the small ones must be relatively accurate, which no
    # normwise method delivers. Reference values: exact products for a diagonal-dominated graded matrix via mpmath-free
    # check: for e = 0 the singular values are |d| exactly.
    d = np.array([10.0 ** (-3 * i) for i in range(21)]); e = np.zeros(20)
    s = bsv(d, e)
    assert np.all(np.abs(s - d) <= 4 * np.finfo(np.float64).eps * d)
    # with coupling: compare with LAPACK on the scaled problem (each singular value to relative 1e-10)
    d = np.array([10.0 ** (-2 * i) for i in range(12)]); e = d[:-1] * 0.5
    s = bsv(d, e); ref = _ref(d, e)
    assert np.all(np.abs(s - ref) <= 1e-10 * ref + 1e-300)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_singular_values_of_a_dense_matrix_through_the_oracle_bidiagonalization(bsv, oracle, dtype):
    rng = np.random.default_rng(102)
    u = np.finfo(dtype).eps
    for (m, n) in [(30, 30), (120, 80), (300, 257)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        W = A.copy(order="F")
        oracle.bidiag(W, 1, 1)
        s = bsv(np.diagonal(W).copy(), np.diagonal(W, 1).copy())
        ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
        assert np.abs(s - ref).max() <= 32 * max(m, n) * u * ref.max(), (m, n)
