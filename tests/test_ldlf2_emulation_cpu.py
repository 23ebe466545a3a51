"""The LDLT leaf kernel (csrc/ldlt_f64.cu: ldlf2_kernel) has not run on hardware yet; tools/emul/ldlf2_emul.cpp is a
statement-by-statement host transcription of it (threads run one after the other inside each barrier interval). This test
builds it with g++ and checks it against the oracle's leaf: bit-identical factors for n <= 64 (same recurrence, same fma
operands), reconstruction for 64 < n <= 128, zero-pivot index / diagonal initialisation, regularisation with signs, the
untouched strict upper triangle. It checks the kernel's algorithm, not CUDA-specific behaviour."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "libldlf2_emul.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-shared", "-o", out,
                           os.path.join(ROOT, "tools", "emul", "ldlf2_emul.cpp")])
    lib = C.CDLL(out)
    lib.emu_ldlf2.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_longlong, C.c_int, C.c_double, C.c_double,
                              C.c_void_p, C.c_void_p]
    lib.emu_ldlf2.restype = C.c_int

    def run(A, delta=0.0, eps=0.0, signs=None, j0=0):
        assert A.flags.f_contiguous and A.dtype == np.float64
        n = A.shape[0]
        info = np.array([-1, 0], dtype=np.int64)
        sp = None
        if signs is not None:
            signs = np.ascontiguousarray(signs, dtype=np.int8)
            sp = signs.ctypes.data
        lib.emu_ldlf2(A.ctypes.data, 1, A.strides[1] // 8, n, j0, int(delta > 0 and eps > 0), eps, delta, sp,
                      info.ctypes.data)
        return int(info[0]), int(info[1])
    return run


def _indefinite(rng, n):
    G = rng.standard_normal((n, n))
    s = np.where(rng.random(n) < 0.4, -1.0, 1.0)
    return np.asfortranarray((G + G.T) / np.sqrt(max(n, 1)) + np.diag(4.0 * s))


def test_leaf_bit_identical_to_the_oracle(emu, oracle):
    rng = np.random.default_rng(81)
    for n in list(range(1, 65)):
        A = _indefinite(rng, n)
        want = A.copy(order="F"); assert oracle.ldlt(want) == (-1, 0)
        got = A.copy(order="F"); got[np.triu_indices(n, 1)] = np.nan
        assert emu(got) == (-1, 0)
        assert np.all(np.isnan(got[np.triu_indices(n, 1)]))
        assert np.array_equal(np.tril(got), np.tril(want)), n


def test_leaf_up_to_128(emu, oracle):
    rng = np.random.default_rng(82)
    u = np.finfo(np.float64).eps
    for n in [65, 96, 100, 127, 128]:
        A = _indefinite(rng, n)
        got = A.copy(order="F"); assert emu(got) == (-1, 0)
        L = np.tril(got, -1) + np.eye(n); D = np.diagonal(got)
        assert np.abs(L @ np.diag(D) @ L.T - A).max() <= 64 * n * u * np.abs(A).max(), n
        want = A.copy(order="F"); oracle.ldlt(want)
        assert np.allclose(np.tril(got), np.tril(want), rtol=1e-10, atol=1e-12), n


def test_zero_pivot_regularisation_and_offsets(emu, oracle):
    rng = np.random.default_rng(83)
    n = 40
    A = _indefinite(rng, n)
    A[:4, :4] = np.diag([2.0, 4.0, 8.0, 0.0]); A[3, :3] = A[:3, 3] = [2.0, 4.0, 8.0]; A[3, 3] = 14.0
    got = A.copy(order="F")
    assert emu(got, j0=1000) == (1003, 0)                       # global column index = j0 + local
    assert np.array_equal(np.diagonal(got)[:4], [2.0, 4.0, 8.0, 0.0]) and np.array_equal(np.diagonal(got)[4:], np.diagonal(A)[4:])
    assert np.array_equal(got[:, 4:], A[:, 4:])                 # nothing beyond the failing column was stored
    Dg = np.asfortranarray(np.diag([1.0, -2.0, 1e-20, -1e-20, 3.0]))
    got = Dg.copy(order="F"); assert emu(got, 1e-3, 1e-10) == (-1, 0)
    assert np.array_equal(np.diagonal(got), [1.0, -2.0, 1e-3, -1e-3, 3.0])
    got = Dg.copy(order="F"); assert emu(got, 1e-3, 1e-10, signs=[1, 1, 1, -1, -1]) == (-1, 2)
    assert np.array_equal(np.diagonal(got), [1.0, 1e-3, 1e-3, -1e-3, -1e-3])
    # signs are indexed by the global column
    sg = np.zeros(10, np.int8); sg[5:] = [1, 1, 1, -1, -1]
    got = Dg.copy(order="F"); assert emu(got, 1e-3, 1e-10, signs=sg, j0=5) == (-1, 2)
    # agreement with the oracle on a regularised random case
    A = _indefinite(rng, 50)
    for j, v in ((10, 1e-14), (30, -3.0)):
        A[j, :] = 0.0; A[:, j] = 0.0; A[j, j] = v
    sg = np.where(np.diagonal(A) > 0, 1, -1).astype(np.int8); sg[30] = 1
    want = A.copy(order="F"); ro = oracle.ldlt(want, delta=1e-2, eps=1e-9, signs=sg)
    got = A.copy(order="F"); assert emu(got, 1e-2, 1e-9, signs=sg) == ro == (-1, 2)
    assert np.array_equal(np.tril(got), np.tril(want))


def test_recursive_driver_model(emu, oracle):
    """ldlt_rec (csrc/ldlt_f64.cu) restated step by step on the host: same split rule (a multiple of the 128-wide leaf
    closest to n / 2), leaf = the kernel transcription, unit-lower solve of the panel against A11 whose diagonal already
    holds D, W <- X and A21 <- X * recip(D) column by column, A22(lower) -= A21 W^T, recursion on A22 with the global column
    offset. Checks the driver's algebra (not its CUDA calls) against the oracle and the reconstruction."""
    rng = np.random.default_rng(84)
    u = np.finfo(np.float64).eps
    NB = 128

    def rec(A, j0, info):
        n = A.shape[0]
        if n <= NB:
            if info[0] >= 0:
                return
            blk = np.asfortranarray(A.copy())
            f, c = emu(blk, j0=j0)
            A[...] = blk
            if f >= 0:
                info[0] = f
            info[1] += c
            return
        n1 = ((n // 2 + NB - 1) // NB) * NB
        if n1 >= n:
            n1 = ((n - 1) // NB) * NB
        A11, A21, A22 = A[:n1, :n1], A[n1:, :n1], A[n1:, n1:]
        rec(A11, j0, info)
        oracle.solve_triangular(A11, A21.T, lower=True, unit=True)          # X^T = L11^-1 A21^T
        W = np.asfortranarray(A21.copy())
        with np.errstate(all="ignore"):                                        # after a zero pivot the GPU driver also runs on
            A21 *= (1.0 / np.diagonal(A11))[None, :]
        oracle.matmul_triangular(A22, 1, True, A21, 0, W.T, 0, -1.0)        # dst structure 1 = TriangularLower
        rec(A22, j0 + n1, info)

    for n in [129, 200, 256, 257, 400, 700, 1000]:
        A = _indefinite(rng, n)
        got = A.copy(order="F")
        info = [-1, 0]
        rec(got, 0, info)
        assert info == [-1, 0], n
        assert np.array_equal(np.triu(got, 1), np.triu(A, 1)), n
        L = np.tril(got, -1) + np.eye(n); D = np.diagonal(got)
        assert np.abs(L @ np.diag(D) @ L.T - A).max() <= 64 * n * u * np.abs(A).max(), n
        want = A.copy(order="F"); assert oracle.ldlt(want) == (-1, 0)
        assert np.allclose(np.tril(got), np.tril(want), rtol=1e-9, atol=1e-11), n
    # failure inside the second leaf: global index through j0
    n = 300
    A = _indefinite(rng, n)
    A[140, :] = 0.0; A[:, 140] = 0.0                                        # column 140 decoupled, zero pivot
    got = A.copy(order="F"); info = [-1, 0]
    rec(got, 0, info)
    assert info[0] == 140
    assert oracle.ldlt(A.copy(order="F"))[0] == 140
