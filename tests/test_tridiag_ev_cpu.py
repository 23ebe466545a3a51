"""The eigenvalue bisection the GPU self-adjoint EVD driver runs per thread (csrc/tridiag_ev.cuh) is plain host/device
code; this test compiles the same header with g++ and checks it against LAPACK: random tridiagonals, zero couplings
(block-diagonal), clusters, negative and positive spectra, n = 1, and end to end through the oracle's
tridiagonalization of a dense symmetric matrix."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tev(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tev") / "libtev.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out,
                           os.path.join(ROOT, "tools", "emul", "tridiag_ev_host.cpp")])
    lib = C.CDLL(out)
    for name in ('tev_f64', 'tev_f32'):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]  # 64-bit pointers
        getattr(lib, name).restype = None

    def run(d, e):
        dt = d.dtype
        n = d.size
        d = np.ascontiguousarray(d); e = np.ascontiguousarray(e.astype(dt)) if n > 1 else np.zeros(1, dt)
        s = np.zeros(n, dt)
        (lib.tev_f64 if dt == np.float64 else lib.tev_f32)(d.ctypes.data, e.ctypes.data, n, s.ctypes.data)
        return s
    return run


def _ref(d, e):
    n = d.size
    T = np.diag(d.astype(np.float64))
    if n > 1:
        T = T + np.diag(e.astype(np.float64), 1) + np.diag(e.astype(np.float64), -1)
    return np.linalg.eigvalsh(T)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tridiagonal_eigenvalues(tev, dtype):
    rng = np.random.default_rng(121)
    u = np.finfo(dtype).eps
    for n in [1, 2, 3, 10, 64, 257, 600]:
        d = rng.standard_normal(n).astype(dtype); e = rng.standard_normal(max(n - 1, 0)).astype(dtype)
        s = tev(d, e); ref = _ref(d, e)
        assert np.all(np.diff(s) >= 0)
        assert np.abs(s - ref).max() <= 8 * n * u * np.abs(ref).max(), n
    d = np.array([3.0, -1.0, -1.0, 2.0, 2.0, 2.0], dtype); e = np.zeros(5, dtype)
    assert np.allclose(tev(d, e), [-1, -1, 2, 2, 2, 3], atol=16 * u)
    d = np.full(40, -5.0, dtype); e = np.full(39, 1e-3, dtype)
    assert np.abs(tev(d, e) - _ref(d, e)).max() <= 64 * u * 5
    d = (1e6 + rng.standard_normal(30)).astype(dtype); e = rng.standard_normal(29).astype(dtype)
    assert np.abs(tev(d, e) - _ref(d, e)).max() <= 64 * u * 1e6


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_eigenvalues_of_a_dense_symmetric_matrix_through_the_oracle_tridiagonalization(tev, oracle, dtype):
    rng = np.random.default_rng(122)
    u = np.finfo(dtype).eps
    for n in [2, 30, 150, 301]:
        G = rng.standard_normal((n, n)).astype(dtype)
        A = np.asfortranarray(G + G.T)
        W = A.copy(order="F")
        oracle.tridiag(W, 1)
        s = tev(np.diagonal(W).copy(), np.diagonal(W, -1).copy())
        ref = np.linalg.eigvalsh(A.astype(np.float64))
        assert np.abs(s - ref).max() <= 32 * n * u * np.abs(ref).max(), n
