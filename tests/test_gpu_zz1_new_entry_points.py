"""Entry points exported after the round's last GPU session, kept apart (and sorted after the files that have run on
hardware) so that a surprise here cannot hide the validated tests under `pytest -x`: f32 triangular solves and the LU
transpose solve through the C ABI. Each body was dry-run on the CPU with the oracle standing in for the C ABI."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def test_f32_triangular_solve_vs_oracle(fb, oracle, cuda_dev):
    """f32 triangular solves through the C ABI (triangular_solve.rs:220-419; same recursion and leaf kernel as f64,
    updates on the f32 GEMM) against the oracle and the componentwise backward bound of substitution
    |T x - b| <= c n u |T| |x|."""
    la = fb.linalg
    rng = np.random.default_rng(43)
    u = np.finfo(np.float32).eps
    fns = {(True, False): la.solve_lower_triangular_in_place, (True, True): la.solve_unit_lower_triangular_in_place,
           (False, False): la.solve_upper_triangular_in_place, (False, True): la.solve_unit_upper_triangular_in_place}
    for n, k in [(1, 1), (2, 3), (5, 5), (33, 70), (64, 130), (129, 70), (600, 130)]:
        T = np.asfortranarray((rng.standard_normal((n, n)) / max(n, 1) + 2 * np.eye(n)).astype(np.float32))
        for lower, unit in itertools.product((True, False), (True, False)):
            for order in "FC":
                Bm = np.array(rng.standard_normal((n, k)), dtype=np.float32, order=order)
                want = Bm.copy(order="K")
                oracle.solve_triangular(T, want, lower, unit)
                got = Bm.copy(order="K")
                fns[(lower, unit)](T, got)
                tol = u * 16 * n * max(1.0, float(np.abs(want).max()))
                assert np.all(np.abs(got - want) <= tol), (n, k, lower, unit, order)
                Tt = (np.tril(T) if lower else np.triu(T)).astype(np.float64)
                if unit:
                    np.fill_diagonal(Tt, 1.0)
                x = got.astype(np.float64)
                resid = np.abs(Tt @ x - Bm.astype(np.float64))
                assert np.all(resid <= 8 * n * u * (np.abs(Tt) @ np.abs(x)) + 1e-30), (n, k, lower, unit, order)


@pytest.mark.parametrize("idx", [np.uint64, np.uint32])
def test_lu_solve_transpose(fb, oracle, cuda_dev, idx):
    """lu/partial_pivoting/solve.rs:55-86: A^T x = b from the factors of A (lower solve with U^T, unit-upper solve with
    L^T, inverse row permutation); expected values from numpy and from the same composition on the CPU."""
    la = fb.linalg
    rng = np.random.default_rng(33)
    for n, k in [(1, 1), (50, 3), (200, 7), (400, 130), (1000, 16)]:
        A = np.asfortranarray(rng.standard_normal((n, n)))
        B = np.asfortranarray(rng.standard_normal((n, k)))
        LU = A.copy(order="F"); p = np.zeros(n, idx); pi = np.zeros(n, idx)
        la.lu_in_place(LU, p, pi)
        X = B.copy(order="F"); la.lu_solve_transpose_in_place(LU, p, pi, X)
        cond = np.linalg.cond(A)
        assert np.all(np.abs(A.T @ X - B) <= EPS * 128 * 8 * n * cond * max(1.0, np.abs(B).max())), (n, k)
        assert np.allclose(X, np.linalg.solve(A.T, B), rtol=1e-7 * max(1, cond / 1e4), atol=1e-9), (n, k)
        # the same composition on the CPU from the same factors
        Xo = B.copy(order="F")
        oracle.solve_triangular(LU.T, Xo, lower=True, unit=False)
        oracle.solve_triangular(LU.T, Xo, lower=False, unit=True)
        Xo = Xo[pi.astype(np.int64)]
        assert np.allclose(X, Xo, rtol=1e-9 * max(1, cond / 1e2), atol=1e-11 * max(1, cond)), (n, k)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_apply_householder_on_the_right(fb, oracle, cuda_dev, dtype):
    """householder.rs:813-854: lhs <- lhs Q and lhs <- lhs Q^H (the left sequences on the transposed view), against Q formed
    explicitly with the left application and against the oracle."""
    la = fb.linalg
    rng = np.random.default_rng(44)
    u = np.finfo(dtype).eps
    for (m, n, k) in [(5, 3, 4), (64, 64, 7), (200, 60, 33), (300, 300, 129)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        QR = A.copy(order="F")
        bs = la.qr_recommended_block_size(m, n)
        H = np.zeros((bs, min(m, n)), dtype=dtype, order="F")
        la.qr_in_place(QR, H)
        Q = np.asfortranarray(np.eye(m, dtype=dtype))
        la.apply_block_householder_sequence_on_the_left_in_place(QR, H, Q)
        for order in "FC":
            M0 = np.array(rng.standard_normal((k, m)), dtype=dtype, order=order)
            tol = 64 * u * m * max(1.0, np.abs(M0).max())
            got = M0.copy(order="K"); la.apply_block_householder_sequence_on_the_right_in_place(QR, H, got)
            assert np.abs(got - M0 @ Q).max() <= tol, (m, n, k, order)
            want = M0.copy(order="K"); oracle.apply_q_transpose_sequence(QR, H, want.T)   # (Q^T M^T)^T = M Q
            assert np.abs(got - want).max() <= tol, (m, n, k, order)
            got = M0.copy(order="K"); la.apply_block_householder_sequence_transpose_on_the_right_in_place(QR, H, got)
            assert np.abs(got - M0 @ Q.T).max() <= tol, (m, n, k, order)
            want = M0.copy(order="K"); oracle.apply_q_sequence(QR, H, want.T)             # (Q M^T)^T = M Q^T
            assert np.abs(got - want).max() <= tol, (m, n, k, order)
