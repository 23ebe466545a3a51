"""Oracle restatement of the solves on the QR factors (qr/no_pivoting/solve.rs:38-176) against the reference's own
tests: test_lstsq (solve.rs:208-280, 100 x 50 c64, block size 4, 3 right-hand sides, plain and conjugated factors,
normal equations to eps * n) and test_solve (solve.rs:282-403, 50 x 50: solve, conjugated solve, transpose solve,
adjoint solve)."""
import numpy as np
import pytest


def _rand_c64(rng, m, n):
    return np.asfortranarray(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n)))


def _approx(a, b, n):
    # utils::approx::ApproxEq::eps() * n: abs tol = rel tol = 128 eps n (utils/approx.rs)
    tol = 128 * np.finfo(np.float64).eps * n
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))) * np.sqrt(n))


def test_lstsq_reference_case(oracle):
    rng = np.random.default_rng(0)
    m, n, k = 100, 50, 3
    A = _rand_c64(rng, m, n); B = _rand_c64(rng, m, k)
    QR = A.copy(order="F")
    H, rank = oracle.qr(QR, block_size=4)
    assert rank == n
    X = B.copy(order="F")
    oracle.qr_solve_lstsq(QR, H, X)
    x = X[:n]
    assert _approx(A.conj().T @ A @ x, A.conj().T @ B, n)
    assert np.allclose(x, np.linalg.lstsq(A, B, rcond=None)[0], atol=1e-11)
    # QR.conjugate() in the reference is a lazy view: same storage, conj_QR = Yes -> solves conj(A) x = B
    X = B.copy(order="F")
    oracle.qr_solve_lstsq(QR, H, X, conj_QR=True)
    x = X[:n]
    assert _approx(A.T @ A.conj() @ x, A.T @ B, n)


def test_solve_reference_case(oracle):
    rng = np.random.default_rng(0)
    n, k = 50, 3
    A = _rand_c64(rng, n, n); B = _rand_c64(rng, n, k)
    QR = A.copy(order="F")
    H, rank = oracle.qr(QR, block_size=4)
    assert rank == n
    QRc, Hc = QR, H  # lazily conjugated views in the reference: same storage, conj_QR = Yes
    X = B.copy(order="F"); oracle.qr_solve(QR, H, X)
    assert _approx(A @ X, B, n)
    X = B.copy(order="F"); oracle.qr_solve(QRc, Hc, X, conj_QR=True)
    assert _approx(A.conj() @ X, B, n)
    X = B.copy(order="F"); oracle.qr_solve_transpose(QR, H, X)
    assert _approx(A.T @ X, B, n)
    X = B.copy(order="F"); oracle.qr_solve_transpose(QRc, Hc, X, conj_QR=True)
    assert _approx(A.conj().T @ X, B, n)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_real_lstsq_matches_numpy(oracle, dtype):
    rng = np.random.default_rng(7)
    u = np.finfo(dtype).eps
    for (m, n, k) in [(1, 1, 1), (5, 3, 2), (64, 64, 5), (300, 40, 7), (257, 129, 1)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        B = np.asfortranarray(rng.standard_normal((m, k)).astype(dtype))
        QR = A.copy(order="F")
        bs = oracle.qr_recommended_block_size(m, n)
        H, rank = oracle.qr(QR, block_size=bs)
        X = B.copy(order="F")
        oracle.qr_solve_lstsq(QR, H, X)
        want = np.linalg.lstsq(A.astype(np.float64), B.astype(np.float64), rcond=None)[0]
        cond = np.linalg.cond(A.astype(np.float64))
        assert np.all(np.abs(X[:n] - want) <= 64 * u * cond * max(m, n) * max(1.0, np.abs(want).max())), (m, n, k)
        if m == n:
            Xt = B.copy(order="F"); oracle.qr_solve_transpose(QR, H, Xt)
            wt = np.linalg.solve(A.astype(np.float64).T, B.astype(np.float64))
            assert np.all(np.abs(Xt - wt) <= 64 * u * cond * n * max(1.0, np.abs(wt).max())), (m, n, k)
