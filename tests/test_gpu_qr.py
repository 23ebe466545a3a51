"""GPU parity tests for Householder QR (no pivoting) and the block-Householder sequence application, f64 and f32,
through the C ABI, against the oracle.

reference tests restated: qr/mod.rs:116-191 (10x2 least squares known answer, 1e-6), qr/no_pivoting/factor.rs:327-538
(`Q R ~ A`, orthogonality, 1e-10 for 64-bit). Contract: rank exact; |A - Q R|, |Q^H Q - I| <=
128 u sqrt(8 max(m, n)) |A|; R's diagonal sign convention beta = -sign(head) * norm; T = striu(V^H V) + diag(tau).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def eps_of(dtype):
    return np.finfo(dtype).eps


def form_q(la, QR, H):
    m = QR.shape[0]
    Q = np.asfortranarray(np.eye(m, dtype=QR.dtype))
    la.apply_block_householder_sequence_on_the_left_in_place(QR, H, Q)
    return Q


def test_qr_lstsq_known_answer(fb):
    la = fb.linalg
    fx = json.load(open(os.path.join(HERE, "golden", "qr_lstsq_example.json")))
    a = np.asfortranarray(np.array(fx["a"])); b = np.asfortranarray(np.array(fx["b"]))
    want = np.array(fx["expected_solution"])
    qr = a.copy(order="F")
    bs = la.qr_recommended_block_size(*a.shape)
    H = np.zeros((bs, 2), order="F")
    info = la.qr_in_place(qr, H)
    assert info.rank == 2
    sol = b.copy(order="F")
    la.apply_block_householder_sequence_transpose_on_the_left_in_place(qr, H, sol)
    x = np.asfortranarray(sol[:2, :])
    la.solve_upper_triangular_in_place(np.asfortranarray(qr[:2, :2]), x)
    assert np.all(np.abs(x - want) <= fx["tolerance"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_qr_vs_oracle(fb, oracle, dtype):
    la = fb.linalg
    rng = np.random.default_rng(51)
    u = eps_of(dtype)
    for (m, n) in [(1, 1), (2, 2), (5, 3), (8, 8), (33, 32), (64, 64), (100, 37), (128, 128), (255, 255), (257, 200), (300, 64),
                   (1000, 130), (2000, 300), (20, 50), (600, 600)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        size = min(m, n)
        for bs in sorted({la.qr_recommended_block_size(m, n), min(15, size), min(64, size)}):
            QRo = A.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=bs)
            QR = A.copy(order="F"); H = np.zeros((bs, size), dtype=dtype, order="F")
            info = la.qr_in_place(QR, H)
            assert info.rank == rank_o == size, (m, n, bs)
            tol = 128 * u * np.sqrt(8 * max(m, n)) * max(1.0, float(np.abs(A).max()))
            Q = form_q(la, QR, H)
            R = np.triu(QR)
            assert np.all(np.abs(Q @ R - A) <= tol), (m, n, bs)
            assert np.all(np.abs(Q.T @ Q - np.eye(m)) <= tol), (m, n, bs)
            # same factors as the oracle (R incl. signs, V, and the T blocks) within a kappa-aware tolerance
            loose = 2e3 * u * max(m, n)
            assert np.allclose(np.triu(QR)[:size], np.triu(QRo)[:size], rtol=loose, atol=loose * np.abs(A).max()), (m, n, bs)
            assert np.allclose(np.tril(QR, -1), np.tril(QRo, -1), rtol=loose, atol=loose), (m, n, bs)
            for j in range(0, size, bs):
                b = min(bs, size - j)
                Tg = np.triu(H[:b, j:j + b]); To = np.triu(Ho[:b, j:j + b])
                fin = np.isfinite(np.diag(To))
                assert np.array_equal(np.isfinite(np.diag(Tg)), fin), (m, n, bs, j)
                To = To.copy(); Tg = Tg.copy()
                To[~np.isfinite(To)] = 0; Tg[~np.isfinite(Tg)] = 0
                assert np.allclose(Tg, To, rtol=loose, atol=loose), (m, n, bs, j)


def _check_rank_deficient(la, oracle, A, bs, rank_true, dtype):
    """The reference's criterion (test_qr, qr/no_pivoting/factor.rs:327-538): rank >= true rank and Q R ~ A; plus parity
    with the oracle: same rank, same R staircase, same compacted reflectors / T blocks for the reflectors that carry
    signal (reflectors beyond the true rank are built from rounding noise, identity reflectors have tau = +inf and an
    uninterpreted v), same zero / +inf fill of Q_coeff beyond the rank (factor.rs:287-299)."""
    m, n = A.shape
    size = min(m, n)
    u = eps_of(dtype)
    QRo = A.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=bs)
    QR = A.copy(order="F"); H = np.full((bs, size), 7.0, dtype=dtype, order="F")
    info = la.qr_in_place(QR, H)
    rank = info.rank
    key = (m, n, bs, rank_true, rank, rank_o)
    assert rank >= min(rank_true, size), key
    sc = max(1.0, float(np.abs(A).max()))
    tol = 128 * u * np.sqrt(8 * max(m, n)) * sc * 4
    Q = form_q(la, QR, H)
    # a column the rank test drops keeps a remainder of norm <= its threshold eps * 16 * (m - row) * ||column|| (factor.rs:52-59)
    # below the staircase, which Q R does not reproduce: that bound is part of the contract (visible in f32, where it exceeds
    # the backward-error tolerance of the kept columns)
    dropped = 16.0 * m * u * np.linalg.norm(A.astype(np.float64), axis=0)[None, :]
    assert np.all(np.abs(Q @ np.triu(QR) - A) <= tol + dropped), key
    assert np.all(np.abs(Q.T @ Q - np.eye(m)) <= tol), key
    if rank != rank_o:
        # both are legal outcomes: reflectors beyond the true rank come from rounding noise (a column whose remaining part is
        # noise passes or fails the threshold, or has an exactly zero tail and advances `row` with an identity reflector,
        # factor.rs:60-63, depending on the summation order). The reference's own test only asks for rank >= true rank.
        # (tiny problems: after the first reflector of a rank-1 4 x 20 f32 matrix every tail is either exactly zero or noise, so
        # anything between the true rank and `size` is legal)
        slack = size if size <= 8 else max(2, size // 50)
        assert abs(rank - rank_o) <= slack and min(rank, rank_o) >= min(rank_true, size), key
        return  # the reference's criterion above is all that applies then
    loose = 4e3 * u * max(m, n)
    assert np.allclose(np.triu(QR)[:size], np.triu(QRo)[:size], rtol=loose, atol=loose * sc), key
    assert np.array_equal(np.isinf(H[:, :]) & (np.arange(bs)[:, None] == (np.arange(size) % bs)[None, :]),
                          np.isinf(Ho) & (np.arange(bs)[:, None] == (np.arange(size) % bs)[None, :])), key
    assert np.all(H[:, rank:][~np.isinf(H[:, rank:])] == 0), key
    live = np.array([np.isfinite(Ho[c % bs, c]) and c < min(rank_true, rank_o) for c in range(rank)], dtype=bool)
    V = np.tril(QR, -1)[:, :rank][:, live]; Vo = np.tril(QRo, -1)[:, :rank][:, live]
    assert np.allclose(V, Vo, rtol=loose, atol=loose), key
    for j in range(0, rank, bs):
        b = min(bs, rank - j)
        lv = live[j:j + b]
        Tg = np.triu(H[:b, j:j + b])[np.ix_(lv, lv)]; To = np.triu(Ho[:b, j:j + b])[np.ix_(lv, lv)]
        assert np.allclose(Tg, To, rtol=loose, atol=loose), key + (j,)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_qr_rank_deficient_vs_oracle(fb, oracle, dtype):
    """The reference's `test_qr` grid (factor.rs:327-538): products A0 * A1 of rank in {1..5, 100, full}; square n in
    {2..257} with block size 1 (first loop) and 15 (second loop), tall / wide m x 20 with block size 15 (third loop)."""
    la = fb.linalg
    rng = np.random.default_rng(52)

    def product(m, n, r):
        if r >= min(m, n):
            return np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        return np.asfortranarray((rng.standard_normal((m, r)) @ rng.standard_normal((r, n))).astype(dtype))

    for rank_true in [1, 2, 3, 4, 5, 100, 10 ** 9]:
        for n in [2, 4, 8, 16, 24, 32, 127, 128, 257]:
            r = min(n, rank_true)
            _check_rank_deficient(la, oracle, product(n, n, r), 1, r, dtype)
        for n in [2, 3, 4, 8, 16, 24, 32, 128, 255, 256, 257, 512]:
            r = min(n, rank_true)
            _check_rank_deficient(la, oracle, product(n, n, r), min(15, n), r, dtype)
        for m in [2, 3, 4, 8, 16, 24, 32, 128, 255, 256, 257, 512]:
            size = min(m, 20)
            r = min(size, rank_true)
            _check_rank_deficient(la, oracle, product(m, 20, r), min(15, size), r, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_qr_rank_deficient_blocked_and_special(fb, oracle, dtype):
    """Deficiency met in the middle of the blocked fast path (hand-over to the general driver at a block boundary), wide
    and tall shapes, the recommended block sizes, zero columns / zero matrix, duplicated columns."""
    la = fb.linalg
    rng = np.random.default_rng(53)
    for (m, n, r) in [(300, 200, 70), (300, 200, 130), (600, 600, 257), (1000, 130, 40), (90, 200, 33), (2000, 300, 299),
                      (4096, 96, 5)]:
        A = np.asfortranarray((rng.standard_normal((m, r)) @ rng.standard_normal((r, n))).astype(dtype))
        for bs in sorted({la.qr_recommended_block_size(m, n), 32, 64}):
            _check_rank_deficient(la, oracle, A, min(bs, min(m, n)), r, dtype)
    # full-rank leading block, then exact copies of earlier columns (deficiency starts inside block 1 with bs = 32)
    B = rng.standard_normal((400, 48)).astype(dtype)
    A = np.asfortranarray(np.concatenate([B, B[:, :40], rng.standard_normal((400, 12)).astype(dtype)], axis=1))
    _check_rank_deficient(la, oracle, A, 32, 60, dtype)
    # zero columns and the zero matrix: rank 0 reflectors for them, Q R = A exactly
    Z = np.zeros((50, 30), dtype=dtype, order="F")
    H = np.full((8, 30), 3.0, dtype=dtype, order="F")
    info = la.qr_in_place(Z, H)
    assert info.rank == 0 and np.all(Z == 0)
    assert np.all(np.isinf(H[np.arange(30) % 8, np.arange(30)])) and np.count_nonzero(H) == 30
    A = np.asfortranarray(rng.standard_normal((64, 40)).astype(dtype)); A[:, [0, 7, 8, 39]] = 0
    _check_rank_deficient(la, oracle, A, 16, 36, dtype)


def test_qr_reference_rank_deficient_fixture(fb, oracle):
    """The reference's `test_rank_deficient` matrix (factor.rs:540-4787; tests/golden/qr_rank_deficient_c64.npz) as a real
    problem — the embedding [[Re, -Im], [Im, Re]], 200 x 80, block size 20 — through the GPU path: Q R ~ A with the
    reference's ApproxEq{1e-10, 1e-10}; the spectrum decays smoothly through 1e-10 ... 1e-11, so the rank is fuzzy and is
    compared with the oracle's within a few columns."""
    la = fb.linalg
    A = np.load(os.path.join(HERE, "golden", "qr_rank_deficient_c64.npz"))["A"]
    E = np.asfortranarray(np.block([[A.real, -A.imag], [A.imag, A.real]]))
    QRo = E.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=20)
    QR = E.copy(order="F"); H = np.zeros((20, 80), order="F")
    info = la.qr_in_place(QR, H)
    assert 50 <= info.rank < 80 and abs(info.rank - rank_o) <= 4, (info.rank, rank_o)
    Q = form_q(la, QR, H)
    d = np.abs(Q @ np.triu(QR) - E)
    assert np.all((d <= 1e-10) | (d <= 1e-10 * np.maximum(np.abs(E), np.abs(Q @ np.triu(QR)))))
    assert np.all(np.abs(Q.T @ Q - np.eye(200)) <= 1e-10)


def test_qr_tall_skinny_property_f32(fb, cuda_dev):
    """BASELINE.json configs[3] shape class (f32 tall-skinny), reduced to 16384 x 1024: Q^T applied to A gives R;
    |R| agrees with a float64 LAPACK QR of the same matrix."""
    import torch
    la = fb.linalg
    m, n = 16384, 1024
    torch.manual_seed(5)
    A0 = torch.randn((n, m), dtype=torch.float32, device=cuda_dev).T  # column-major m x n
    A = A0.clone(memory_format=torch.preserve_format)
    bs = la.qr_recommended_block_size(m, n)
    H = torch.zeros((n, bs), dtype=torch.float32, device=cuda_dev).T    # column-major bs x n
    info = la.qr_in_place(A, H)
    assert info.rank == n
    B = A0.clone(memory_format=torch.preserve_format)
    la.apply_block_householder_sequence_transpose_on_the_left_in_place(A, H, B)
    R = torch.triu(A[:n, :])
    u = float(np.finfo(np.float32).eps)
    scale = float(A0.abs().max()) * np.sqrt(8 * m)
    assert float((B[:n, :] - R).abs().max()) <= 128 * u * scale
    assert float(B[n:, :].abs().max()) <= 128 * u * scale
    Rl = np.linalg.qr(A0.cpu().numpy().astype(np.float64), mode="r")
    assert np.allclose(np.abs(R.cpu().numpy()), np.abs(Rl), rtol=2e-3, atol=2e-3 * np.abs(Rl).max())


def test_qr_outputs_interleaved_in_one_host_buffer(fb):
    """Two host outputs of one call living in the same address range (A = buf[0::2], Q_coeff = buf[1::2]; general strides are
    mirrored by address range inside the call): each must come back complete — only a view's own elements are written back."""
    la = fb.linalg
    rng = np.random.default_rng(77)
    m, n, bs = 60, 24, 8
    A0 = np.asfortranarray(rng.standard_normal((m, n)))
    want = A0.copy(order="F"); Hw = np.zeros((bs, n), order="F")
    la.qr_in_place(want, Hw)
    big = np.full(2 * m * n, 7.0)
    A = big[0::2][:m * n].reshape((m, n), order="F")
    H = big[1::2][:bs * n].reshape((bs, n), order="F")
    A[...] = A0
    H[...] = 0.0
    assert A.strides == (16, 16 * m) and not A.flags.f_contiguous
    la.qr_in_place(A, H)
    assert np.array_equal(A, want) and np.array_equal(H, Hw)
    assert np.all(big[1::2][bs * n:] == 7.0)  # the rest of the odd slots was nobody's
