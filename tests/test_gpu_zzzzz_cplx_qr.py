"""Complex (c64 / c32) Householder QR without pivoting, block-Householder sequences and the QR solves through the C ABI
(csrc/cplx.cu) against the oracle's complex restatement (qr/no_pivoting/factor.rs:11-301, householder.rs:59-107, 132-272, 370-620,
724-854, qr/no_pivoting/solve.rs:38-176).

Contract (reference test qr/no_pivoting/factor.rs:327-538 at 1e-10 for c64): rank exact on well-separated
inputs, |A - Q R|, |Q^H Q - I| <= 128 u sqrt(8 max(m, n)) |A|, R's sign convention beta = -sign(head) * norm,
T = striu(V^H V) + diag(tau); rank-deficient inputs follow the reference's column-skipping path (same staircase, compacted
reflectors, zero / +inf fill of Q_coeff beyond the rank); the reference's own rank-deficient c64 matrix (factor.rs:540-4787) as
the committed fixture."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DTYPES = [np.complex128, np.complex64]


def ueps(dtype):
    return np.finfo(dtype).eps


def crandn(rng, shape, dtype):
    return np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype))


def form_q(la, QR, H):
    m = QR.shape[0]
    Q = np.asfortranarray(np.eye(m, dtype=QR.dtype))
    la.apply_block_householder_sequence_on_the_left_in_place(QR, H, Q)
    return Q


def approx_eq(a, b, abs_tol, rel_tol):
    d = np.abs(a - b)
    return np.all((d <= abs_tol) | (d <= rel_tol * np.maximum(np.abs(a), np.abs(b))))


@pytest.mark.parametrize("dtype", DTYPES)
def test_cplx_qr_vs_oracle(fb, oracle, dtype):
    la = fb.linalg
    rng = np.random.default_rng(161)
    u = ueps(dtype)
    for (m, n) in [(1, 1), (2, 2), (5, 3), (8, 8), (33, 32), (64, 64), (100, 37), (128, 128), (257, 200), (300, 64), (1000, 130),
                   (20, 50), (600, 600)]:
        A = crandn(rng, (m, n), dtype)
        size = min(m, n)
        for bs in sorted({la.qr_recommended_block_size(m, n), 1, min(15, size), min(64, size)}):
            QRo = A.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=bs)
            QR = A.copy(order="F"); H = np.zeros((bs, size), dtype=dtype, order="F")
            info = la.qr_in_place(QR, H)
            assert info.rank == rank_o == size, (m, n, bs, info.rank, rank_o)
            tol = 128 * u * np.sqrt(8 * max(m, n)) * max(1.0, float(np.abs(A).max()))
            Q = form_q(la, QR, H).astype(np.complex128)
            R = np.triu(QR).astype(np.complex128)
            assert np.all(np.abs(Q @ R - A) <= tol), (m, n, bs)
            assert np.all(np.abs(Q.conj().T @ Q - np.eye(m)) <= tol), (m, n, bs)
            loose = 2e3 * u * max(m, n)
            assert np.allclose(np.triu(QR)[:size], np.triu(QRo)[:size], rtol=loose, atol=loose * np.abs(A).max()), (m, n, bs)
            assert np.allclose(np.tril(QR, -1), np.tril(QRo, -1), rtol=loose, atol=loose), (m, n, bs)
            for j in range(0, size, bs):
                b = min(bs, size - j)
                assert np.allclose(np.triu(H[:b, j:j + b]), np.triu(Ho[:b, j:j + b]), rtol=loose, atol=loose), (m, n, bs, j)


def _check_rank_deficient(la, oracle, A, bs, rank_true):
    dtype = A.dtype
    m, n = A.shape
    size = min(m, n)
    u = ueps(dtype)
    QRo = A.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=bs)
    QR = A.copy(order="F"); H = np.full((bs, size), 7.0, dtype=dtype, order="F")
    info = la.qr_in_place(QR, H)
    rank = info.rank
    key = (str(dtype), m, n, bs, rank_true, rank, rank_o)
    assert rank >= min(rank_true, size), key
    sc = max(1.0, float(np.abs(A).max()))
    tol = 128 * u * np.sqrt(8 * max(m, n)) * sc * 4
    Q = form_q(la, QR, H).astype(np.complex128)
    dropped = 16.0 * m * u * np.linalg.norm(A.astype(np.complex128), axis=0)[None, :]  # see test_gpu_qr._check_rank_deficient
    assert np.all(np.abs(Q @ np.triu(QR).astype(np.complex128) - A) <= tol + dropped), key
    assert np.all(np.abs(Q.conj().T @ Q - np.eye(m)) <= tol), key
    # Q_coeff beyond the rank: zero columns with +inf on the block diagonals (factor.rs:287-299)
    for c in range(rank, size):
        col = H[:, c]
        assert np.isinf(col[c % bs].real) and col[c % bs].real > 0 and np.count_nonzero(col) == 1, key + (c,)
    if rank != rank_o:
        slack = size if size <= 8 else max(2, size // 50)  # both legal: reflectors beyond the true rank come from rounding noise
        assert abs(rank - rank_o) <= slack and min(rank, rank_o) >= min(rank_true, size), key
        return
    loose = 4e3 * u * max(m, n)
    assert np.allclose(np.triu(QR)[:size], np.triu(QRo)[:size], rtol=loose, atol=loose * sc), key
    dg = np.arange(bs)[:, None] == (np.arange(size) % bs)[None, :]
    assert np.array_equal(np.isinf(H.real) & dg, np.isinf(Ho.real) & dg), key
    live = np.array([np.isfinite(Ho[c % bs, c].real) and c < min(rank_true, rank_o) for c in range(rank)], dtype=bool)
    V = np.tril(QR, -1)[:, :rank][:, live]; Vo = np.tril(QRo, -1)[:, :rank][:, live]
    assert np.allclose(V, Vo, rtol=loose, atol=loose), key
    for j in range(0, rank, bs):
        b = min(bs, rank - j)
        lv = live[j:j + b]
        Tg = np.triu(H[:b, j:j + b])[np.ix_(lv, lv)]; To = np.triu(Ho[:b, j:j + b])[np.ix_(lv, lv)]
        assert np.allclose(Tg, To, rtol=loose, atol=loose), key + (j,)


@pytest.mark.parametrize("dtype", DTYPES)
def test_cplx_qr_rank_deficient_vs_oracle(fb, oracle, dtype):
    """The reference's `test_qr` grid (factor.rs:327-538), thinned: products of rank in {1, 3, 5, 100, full}; block sizes 1 and 15;
    square, tall and wide; plus deficiency in the middle of larger blocks, duplicated / zero columns and the zero matrix."""
    la = fb.linalg
    rng = np.random.default_rng(162)

    def product(m, n, r):
        if r >= min(m, n):
            return crandn(rng, (m, n), dtype)
        return np.asfortranarray((crandn(rng, (m, r), np.complex128) @ crandn(rng, (r, n), np.complex128)).astype(dtype))

    for rank_true in [1, 3, 5, 100, 10 ** 9]:
        for n in [2, 4, 16, 32, 127, 257]:
            r = min(n, rank_true)
            _check_rank_deficient(la, oracle, product(n, n, r), 1 if n <= 32 else min(15, n), r)
            _check_rank_deficient(la, oracle, product(n, n, r), min(15, n), r)
        for m in [2, 3, 16, 24, 128, 255, 512]:
            size = min(m, 20)
            r = min(size, rank_true)
            _check_rank_deficient(la, oracle, product(m, 20, r), min(15, size), r)
    for (m, n, r) in [(300, 200, 70), (600, 600, 257), (90, 200, 33), (1000, 130, 40)]:
        A = product(m, n, r)
        for bs in sorted({la.qr_recommended_block_size(m, n), 32, 64}):
            _check_rank_deficient(la, oracle, A, min(bs, min(m, n)), r)
    B = crandn(rng, (400, 48), dtype)
    A = np.asfortranarray(np.concatenate([B, B[:, :40], crandn(rng, (400, 12), dtype)], axis=1))
    _check_rank_deficient(la, oracle, A, 32, 60)
    Z = np.zeros((50, 30), dtype=dtype, order="F")
    H = np.full((8, 30), 3.0, dtype=dtype, order="F")
    info = la.qr_in_place(Z, H)
    assert info.rank == 0 and np.all(Z == 0)
    assert np.all(np.isinf(H[np.arange(30) % 8, np.arange(30)].real)) and np.count_nonzero(H) == 30
    A = crandn(rng, (64, 40), dtype); A[:, [0, 7, 8, 39]] = 0
    _check_rank_deficient(la, oracle, A, 16, 36)


def test_cplx_qr_reference_rank_deficient_fixture(fb, oracle):
    """The reference's `test_rank_deficient` matrix (factor.rs:540-4787; tests/golden/qr_rank_deficient_c64.npz), 100 x 40 c64,
    block size 20 as in the CPU pin of the oracle (test_reference_fixtures_cpu.py): Q R ~ A and Q^H Q ~ I with the reference's
    ApproxEq{1e-10, 1e-10}; the spectrum decays smoothly through 1e-10 ... 1e-11, so the rank is compared within a few columns."""
    la = fb.linalg
    A = np.asfortranarray(np.load(os.path.join(HERE, "golden", "qr_rank_deficient_c64.npz"))["A"])
    QRo = A.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=20)
    QR = A.copy(order="F"); H = np.zeros((20, 40), dtype=np.complex128, order="F")
    info = la.qr_in_place(QR, H)
    assert 25 <= info.rank < 40 and abs(info.rank - rank_o) <= 3, (info.rank, rank_o)
    Q = form_q(la, QR, H)
    assert approx_eq(Q @ np.triu(QR), A, 1e-10, 1e-10)
    assert approx_eq(Q.conj().T @ Q, np.eye(100), 1e-10, 1e-10)
    for c in range(info.rank, 40):
        col = H[:, c]
        assert np.isinf(col[c % 20].real) and np.count_nonzero(col) == 1


@pytest.mark.parametrize("dtype", DTYPES)
def test_cplx_householder_sequences_and_solves(fb, oracle, dtype):
    """rhs <- Q rhs, Q^H rhs, conj(Q) rhs, Q^T rhs and the right-hand variants (householder.rs:724-854) against the explicit Q, and
    the three QR solves with both conjugation settings (qr/no_pivoting/solve.rs:38-176) against numpy."""
    la = fb.linalg
    rng = np.random.default_rng(163)
    u = ueps(dtype)
    for (m, n, bs) in [(9, 9, 4), (70, 70, 16), (130, 40, 15), (200, 200, 32)]:
        A = crandn(rng, (m, n), dtype)
        QR = A.copy(order="F"); H = np.zeros((bs, min(m, n)), dtype=dtype, order="F")
        assert la.qr_in_place(QR, H).rank == min(m, n)
        Q = form_q(la, QR, H).astype(np.complex128)
        tol = 256 * u * np.sqrt(8 * m) * 8
        M = crandn(rng, (m, 5), dtype)
        for conj in (0, 1):
            Qc = Q.conj() if conj else Q
            X = M.copy(order="F"); la.apply_block_householder_sequence_on_the_left_in_place(QR, H, X, conj)
            assert np.all(np.abs(X - Qc @ M) <= tol * np.abs(M).max()), (m, n, bs, conj, "left")
            Xo = M.copy(order="F"); oracle.apply_q_sequence(QR, H, Xo, conj_lhs=bool(conj))
            assert np.all(np.abs(X - Xo) <= tol * np.abs(M).max()), (m, n, bs, conj, "left vs oracle")
            X = M.copy(order="F"); la.apply_block_householder_sequence_transpose_on_the_left_in_place(QR, H, X, conj)
            assert np.all(np.abs(X - Qc.T @ M) <= tol * np.abs(M).max()), (m, n, bs, conj, "transpose left")
            Mr = np.asfortranarray(M.T.copy())
            X = Mr.copy(order="F"); la.apply_block_householder_sequence_on_the_right_in_place(QR, H, X, conj)
            assert np.all(np.abs(X - Mr @ Qc) <= tol * np.abs(M).max()), (m, n, bs, conj, "right")
            X = Mr.copy(order="F"); la.apply_block_householder_sequence_transpose_on_the_right_in_place(QR, H, X, conj)
            assert np.all(np.abs(X - Mr @ Qc.T) <= tol * np.abs(M).max()), (m, n, bs, conj, "transpose right")
        A64 = A.astype(np.complex128)
        kappa = np.linalg.cond(A64)
        for conj in (0, 1):
            Ae = A64.conj() if conj else A64
            B = crandn(rng, (m, 3), dtype)
            X = B.copy(order="F"); la.qr_solve_lstsq_in_place(QR, H, QR, X, conj)
            want = np.linalg.lstsq(Ae, B.astype(np.complex128), rcond=None)[0]
            assert np.all(np.abs(X[:n] - want) <= 64 * u * m * kappa * np.abs(want).max()), (m, n, bs, conj, "lstsq")
            if m == n:
                X = B.copy(order="F"); la.qr_solve_in_place(QR, H, QR, X, conj)
                assert np.all(np.abs(Ae @ X - B) <= 64 * u * m * kappa * np.abs(B).max()), (m, n, bs, conj, "solve")
                X = B.copy(order="F"); la.qr_solve_transpose_in_place(QR, H, QR, X, conj)
                assert np.all(np.abs(Ae.T @ X - B) <= 64 * u * m * kappa * np.abs(B).max()), (m, n, bs, conj, "solve transpose")
