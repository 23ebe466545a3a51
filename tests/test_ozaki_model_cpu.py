"""CPU test pinning the numerical model of the planned int8-sliced f64 GEMM (tools/next/ozaki_emulation.py; DESIGN.md §7
item 5): with 8 signed 7-bit slices per operand and the orders p + q >= 8 dropped, the result is as accurate as a plain f64
GEMM relative to (|A||B|)_ij, the int32 order accumulators cannot overflow up to k = 32768, and the slices reconstruct the
scaled operand exactly."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ozaki_emulation", os.path.join(HERE, "..", "tools", "next", "ozaki_emulation.py"))
oz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(oz)


def test_slices_are_int8_and_reconstruct_the_operand():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((37, 129)) * np.exp(rng.uniform(-20, 20, (37, 1)))
    X[5, :] = 0.0  # an all-zero row keeps exponent 0
    S = 9
    sl, e = oz.slices(X, 1, S)
    assert all(s.dtype == np.int8 and np.abs(s.astype(np.int64)).max() <= 64 for s in sl)
    rec = sum(np.ldexp(s.astype(np.float64), -6 - 7 * p) for p, s in enumerate(sl))
    rec = np.ldexp(rec, e)
    assert np.abs(rec - X).max() <= 2.0 ** (-6 - 7 * (S - 1) - 1) * np.abs(X).max(axis=1, keepdims=True).max() * 2
    assert np.all(rec[5] == 0.0)


def test_eight_slices_match_f64_gemm_accuracy():
    rng = np.random.default_rng(4)
    u = 2.0 ** -53
    for k in (17, 256, 2000):
        A = rng.standard_normal((40, k)) * np.exp(rng.uniform(-5, 5, (40, 1)))
        B = rng.standard_normal((k, 24)) * np.exp(rng.uniform(-5, 5, (1, 24)))
        ref = oz.exact(A, B)
        scale = np.abs(A).astype(np.longdouble) @ np.abs(B).astype(np.longdouble)
        err8 = float(np.max(np.abs(oz.ozaki_gemm(A, B, 8) - ref) / scale))
        err64 = float(np.max(np.abs(A @ B - ref) / scale))
        assert err8 <= 4 * u, (k, err8)
        assert err8 <= 2.0 * max(err64, u), (k, err8, err64)


def test_int32_accumulators_hold_up_to_k_32768():
    # worst case per order d: (d + 1) pairs x k terms x 64 x 64
    assert 8 * 32768 * 64 * 64 < 2 ** 31
