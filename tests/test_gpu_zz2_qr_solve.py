"""GPU tests of the solves on the QR factors, f64 and f32, through the C ABI.

reference tests restated for the real types: qr/no_pivoting/solve.rs:208-280 (test_lstsq: 100 x 50, block size 4, 3
right-hand sides, normal equations to eps * n) and 282-403 (test_solve: 50 x 50, solve and transpose solve). The
expected solution also comes from the oracle's restatement applied to the same factors (tests/test_oracle_qr_solve_cpu.py
pins that restatement against the reference's complex cases).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _factor(la, A, bs=None):
    m, n = A.shape
    QR = A.copy(order="F")
    bs = bs or la.qr_recommended_block_size(m, n)
    H = np.zeros((bs, min(m, n)), dtype=A.dtype, order="F")
    assert la.qr_in_place(QR, H).rank == min(m, n)
    return QR, H


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lstsq_reference_shapes(fb, oracle, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(61)
    u = np.finfo(dtype).eps
    for (m, n, k, bs) in [(100, 50, 3, 4), (1, 1, 1, None), (5, 3, 2, None), (64, 64, 5, None), (300, 40, 7, None),
                          (257, 129, 1, None), (2000, 300, 33, None)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        B = np.asfortranarray(rng.standard_normal((m, k)).astype(dtype))
        QR, H = _factor(la, A, bs)
        X = B.copy(order="F")
        la.qr_solve_lstsq_in_place(QR, H, QR, X)
        x = X[:n].astype(np.float64)
        A64, B64 = A.astype(np.float64), B.astype(np.float64)
        cond = np.linalg.cond(A64)
        # the reference's criterion: the normal equations
        lhs, rhs = A64.T @ (A64 @ x), A64.T @ B64
        assert np.all(np.abs(lhs - rhs) <= 8 * u * max(m, n) * cond * np.abs(A64).max() ** 2 * max(1.0, np.abs(x).max()) * np.sqrt(m)), (m, n, k)
        want = np.linalg.lstsq(A64, B64, rcond=None)[0]
        assert np.all(np.abs(x - want) <= 4 * u * cond * max(m, n) * max(1.0, np.abs(want).max())), (m, n, k)
        # same composition on the CPU from the same factors
        Xo = B.copy(order="F")
        oracle.qr_solve_lstsq(QR, H, Xo)
        assert np.all(np.abs(X - Xo) <= 4 * u * cond * max(m, n) * max(1.0, np.abs(Xo).max())), (m, n, k)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_square_solve_and_transpose(fb, oracle, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(62)
    u = np.finfo(dtype).eps
    for (n, k, bs) in [(50, 3, 4), (1, 2, None), (129, 5, None), (600, 64, None)]:
        A = np.asfortranarray((rng.standard_normal((n, n)) + np.sqrt(n) * np.eye(n)).astype(dtype))
        B = np.asfortranarray(rng.standard_normal((n, k)).astype(dtype))
        QR, H = _factor(la, A, bs)
        A64, B64 = A.astype(np.float64), B.astype(np.float64)
        cond = np.linalg.cond(A64)
        tol = 4 * u * cond * n
        X = B.copy(order="F"); la.qr_solve_in_place(QR, H, QR, X)
        want = np.linalg.solve(A64, B64)
        assert np.all(np.abs(X - want) <= tol * max(1.0, np.abs(want).max())), (n, k)
        Xo = B.copy(order="F"); oracle.qr_solve(QR, H, Xo)
        assert np.all(np.abs(X - Xo) <= tol * max(1.0, np.abs(Xo).max())), (n, k)
        Xt = B.copy(order="F"); la.qr_solve_transpose_in_place(QR, H, QR, Xt)
        want_t = np.linalg.solve(A64.T, B64)
        assert np.all(np.abs(Xt - want_t) <= tol * max(1.0, np.abs(want_t).max())), (n, k)
        Xto = B.copy(order="F"); oracle.qr_solve_transpose(QR, H, Xto)
        assert np.all(np.abs(Xt - Xto) <= tol * max(1.0, np.abs(Xto).max())), (n, k)


def test_separate_r_padded_rhs_and_device_buffers(fb, oracle, cuda_dev):
    """R passed as its own matrix (not aliasing Q_basis), a right-hand side with a padded leading dimension, and
    device-resident operands (used in place, no staging) give the same answer as the plain host call."""
    import torch
    la = fb.linalg
    rng = np.random.default_rng(63)
    m, n, k = 500, 120, 9
    A = np.asfortranarray(rng.standard_normal((m, n)))
    B = np.asfortranarray(rng.standard_normal((m, k)))
    QR, H = _factor(la, A)
    ref = B.copy(order="F"); la.qr_solve_lstsq_in_place(QR, H, QR, ref)
    # separate R (only its first n rows' upper triangle is read; poison the rest)
    R = np.asfortranarray(np.triu(QR[:n]) + np.tril(np.full((n, n), np.nan), -1))
    X = B.copy(order="F"); la.qr_solve_lstsq_in_place(QR, H, R, X)
    assert np.allclose(X, ref, rtol=1e-12, atol=1e-13)
    # padded rhs
    big = np.full((m + 13, k), np.nan, order="F"); Xp = big[:m]; Xp[...] = B
    la.qr_solve_lstsq_in_place(QR, H, QR, Xp)
    assert np.allclose(Xp, ref, rtol=1e-12, atol=1e-13) and np.all(np.isnan(big[m:]))
    # device-resident operands (column-major via transposed row-major tensors)
    dQR = torch.from_numpy(np.ascontiguousarray(QR.T)).to(cuda_dev).t()
    dH = torch.from_numpy(np.ascontiguousarray(H.T)).to(cuda_dev).t()
    dX = torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda_dev).t()
    la.qr_solve_lstsq_in_place(dQR, dH, dQR, dX)
    torch.cuda.synchronize()
    assert np.allclose(dX.cpu().numpy(), ref, rtol=1e-12, atol=1e-13)
    assert np.array_equal(dQR.cpu().numpy(), QR)  # factors are read-only
