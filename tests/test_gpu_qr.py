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


def test_qr_rank_deficient_is_reported(fb):
    la = fb.linalg
    rng = np.random.default_rng(52)
    A0 = rng.standard_normal((60, 3)); A1 = rng.standard_normal((3, 20))
    A = np.asfortranarray(A0 @ A1)
    H = np.zeros((8, 20), order="F")
    with pytest.raises(RuntimeError):
        la.qr_in_place(A.copy(order="F"), H)


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
