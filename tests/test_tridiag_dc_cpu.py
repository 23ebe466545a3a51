"""The scalar core of the GPU divide-and-conquer eigensolver (csrc/tridiag_dc_core.cuh: implicit-QL leaves, dlaed2-style
deflation scan, bisection secular solver) run on the CPU by tools/emul/tridiag_dc_host.cpp with the kernels' glue as plain
loops, against LAPACK (scipy.linalg.eigh_tridiagonal) and the reference's own criteria: reconstruction Q diag(lam) Q^T ~ T and
orthogonality at eps * 128 * sqrt(8 n) (svd/mod.rs:780-783 style), on the reference's test bidiagonals in Golub-Kahan form too."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def tdc(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tdc") / "libtdc.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, os.path.join(ROOT, "tools", "emul", "tridiag_dc_host.cpp")])
    lib = C.CDLL(out)
    lib.tdc_eig.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.tdc_eig.restype = C.c_int

    def run(d, e):
        n = d.size
        d = np.ascontiguousarray(d, dtype=np.float64); e = np.ascontiguousarray(np.r_[e, 0.0], dtype=np.float64)
        lam = np.zeros(n); Q = np.zeros((n, n), order="F")
        rc = lib.tdc_eig(d.ctypes.data, e.ctypes.data, n, lam.ctypes.data, Q.ctypes.data)
        assert rc == 0
        return lam, Q
    return run


def check(d, e, lam, Q, what):
    n = d.size
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    scale = max(1.0, np.abs(T).max())
    tol = np.finfo(float).eps * 128 * np.sqrt(8 * n)
    assert np.all(np.diff(lam) >= 0), what
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= tol, (what, "orthogonality", np.abs(Q.T @ Q - np.eye(n)).max())
    assert np.abs(Q @ np.diag(lam) @ Q.T - T).max() <= tol * scale, (what, "reconstruction", np.abs(Q @ np.diag(lam) @ Q.T - T).max())
    import scipy.linalg as sla
    ref = sla.eigh_tridiagonal(d, e, eigvals_only=True) if n > 1 else d.copy()
    assert np.abs(lam - ref).max() <= tol * scale, (what, "values")


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 64, 65, 100, 257, 600])
def test_random(tdc, n):
    rng = np.random.default_rng(n)
    d = rng.standard_normal(n); e = rng.standard_normal(max(n - 1, 0))
    lam, Q = tdc(d, e)
    check(d, e, lam, Q, ("random", n))


def test_special_matrices(tdc):
    rng = np.random.default_rng(5)
    n = 201
    # Wilkinson W+: pairs of eigenvalues agreeing to many digits (heavy type-2 deflation)
    d = np.abs(np.arange(n) - n // 2).astype(float); e = np.ones(n - 1)
    check(d, e, *tdc(d, e), "wilkinson")
    # glued Wilkinson blocks, zero / tiny couplings (type-1 deflation, rho = 0)
    d = np.tile(np.abs(np.arange(21) - 10).astype(float), 10); e = np.ones(209); e[20::21] = 1e-9; e[41] = 0.0
    check(d, e, *tdc(d, e), "glued")
    # all zeros, identity, constant off-diagonal (1-2-1 Toeplitz), huge / tiny scales
    check(np.zeros(70), np.zeros(69), *tdc(np.zeros(70), np.zeros(69)), "zeros")
    check(np.ones(70), np.zeros(69), *tdc(np.ones(70), np.zeros(69)), "identity")
    d = 2 * np.ones(300); e = -np.ones(299)
    check(d, e, *tdc(d, e), "toeplitz")
    d = rng.standard_normal(150) * 1e150; e = rng.standard_normal(149) * 1e150
    lam, Q = tdc(d, e)
    check(d / 1e150, e / 1e150, lam / 1e150, Q, "huge")
    # graded
    d = 10.0 ** (-np.arange(120) / 6.0); e = d[:-1] * 0.3
    check(d, e, *tdc(d, e), "graded")


@pytest.mark.parametrize("name", ["svd64", "svd128", "svd512", "zink"])
def test_golub_kahan_form_of_the_reference_bidiagonals(tdc, name):
    """T_GK = P [0 B^T; B 0] P^T (zero diagonal, off-diagonals d_1, e_1, d_2, ...): eigenvalues +-sigma_i. This is the matrix the
    SVD driver hands to the eigensolver."""
    if name == "zink":
        fx = json.load(open(os.path.join(GOLD, "svd_zink.json")))
        dd = np.array(fx["diag"]); ss = np.array(fx["subdiag"])[:-1]
    else:
        fx = np.load(os.path.join(GOLD, f"svd_bidiag_{name}.npz"))
        dd = fx["diag"]; ss = fx["subdiag"][:-1]
    n = dd.size
    off = np.zeros(2 * n - 1); off[0::2] = dd; off[1::2] = ss
    lam, Q = tdc(np.zeros(2 * n), off)
    check(np.zeros(2 * n), off, lam, Q, name)
    sv = np.linalg.svd(np.diag(dd) + np.diag(ss, -1), compute_uv=False)
    assert np.abs(lam[n:][::-1] - sv).max() <= 64 * n * np.finfo(float).eps * sv[0]
    assert np.abs(lam[:n] + sv).max() <= 64 * n * np.finfo(float).eps * sv[0]
