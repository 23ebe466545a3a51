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
