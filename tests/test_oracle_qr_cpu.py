"""CPU tests pinning the oracle's Householder QR (test infrastructure) to the reference's own tests:
  test_example (known answer, numpy-derived lstsq)  faer/src/linalg/qr/mod.rs:116-191            tol 1e-6
  test_qr (c64; rank-deficient A0*A1; bs in {1, 15, recommended}; square + tall)
                                                    faer/src/linalg/qr/no_pivoting/factor.rs:327-538  tol 1e-10
  norm_l2 scaling tests                             faer/src/linalg/reductions/norm_l2.rs:179-219      rel 1e-14
and cross-checked against LAPACK (numpy.linalg.qr): |R| agrees, R's diagonal sign convention is beta = -sign(head)*norm.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def randn(rng, shape, dtype):
    if np.dtype(dtype).kind == "c":
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


def form_q(oracle, QR, H):
    m = QR.shape[0]
    Q = np.asfortranarray(np.eye(m, dtype=QR.dtype))
    oracle.apply_q_sequence(QR, H, Q)
    return Q


def test_qr_lstsq_known_answer(oracle):
    fx = json.load(open(os.path.join(HERE, "golden", "qr_lstsq_example.json")))
    a = np.asfortranarray(np.array(fx["a"])); b = np.asfortranarray(np.array(fx["b"]))
    want = np.array(fx["expected_solution"])
    qr = a.copy(order="F")
    H, rank = oracle.qr(qr)
    assert rank == 2
    sol = b.copy(order="F")
    oracle.apply_q_transpose_sequence(qr, H, sol, conj_lhs=True)
    x = sol[:2, :].copy(order="F")
    oracle.solve_triangular(np.asfortranarray(qr[:2, :2]), x, lower=False, unit=False)
    assert np.all(np.abs(x - want) <= fx["tolerance"])


@pytest.mark.parametrize("dtype", [np.complex128, np.float64, np.float32])
def test_qr_reconstruction_and_rank(oracle, dtype):
    rng = np.random.default_rng(0)
    tol = 1e-10 if np.dtype(dtype).itemsize >= 8 and dtype != np.float32 else 2e-4
    for (m, n) in [(1, 1), (2, 2), (3, 3), (4, 4), (8, 8), (16, 16), (24, 24), (32, 32), (128, 128), (255, 255), (256, 256),
                   (257, 257), (8, 4), (128, 20), (255, 20), (256, 20), (257, 20), (300, 64), (20, 50)]:
        size = min(m, n)
        for rank_true in sorted({1, 2, 3, 5, 100, size} & set(range(1, size + 1)) | {size}):
            A0 = randn(rng, (m, rank_true), dtype); A1 = randn(rng, (rank_true, n), dtype)
            A = np.asfortranarray(A0 @ A1) if rank_true < size else np.asfortranarray(randn(rng, (m, n), dtype))
            for bs in sorted({1, min(15, size), oracle.qr_recommended_block_size(m, n)}):
                QR = A.copy(order="F")
                H, rank = oracle.qr(QR, block_size=bs)
                assert rank >= min(rank_true, size) or rank_true == size and rank == size, (m, n, rank_true, bs, rank)
                if rank_true == size:
                    assert rank == size
                Q = form_q(oracle, QR, H)
                R = np.triu(QR)
                scale = max(1.0, float(np.abs(A).max()))
                assert np.all(np.abs(Q @ R - A) <= tol * scale * max(1, size) ** 0.5), (m, n, rank_true, bs)
                assert np.all(np.abs(Q.conj().T @ Q - np.eye(m)) <= tol * max(1, m) ** 0.5), (m, n, rank_true, bs)


def test_qr_matches_lapack_and_sign_convention(oracle):
    rng = np.random.default_rng(1)
    for (m, n) in [(50, 50), (300, 40), (129, 129)]:
        A = np.asfortranarray(rng.standard_normal((m, n)))
        QR = A.copy(order="F")
        H, rank = oracle.qr(QR)
        assert rank == min(m, n)
        R = np.triu(QR)[:n, :]
        Rl = np.linalg.qr(A, mode="r")
        assert np.allclose(np.abs(R), np.abs(Rl), rtol=1e-10, atol=1e-10)
        # beta = -sign(head) * norm for the very first reflector (householder.rs:85-99)
        assert np.sign(R[0, 0]) == -np.sign(A[0, 0])
        # tau = (1 + |v_tail|^2) / 2 on the diagonal of each T block; T = striu(V^H V) + diag(tau)
        bs = H.shape[0]
        V = np.tril(QR, -1)[:, :n] + np.eye(m, n)
        for j in range(0, n, bs):
            b = min(bs, n - j)
            Vb = V[:, j:j + b]
            Tb = H[:b, j:j + b]
            G = Vb.T @ Vb
            fin = np.isfinite(np.diag(Tb))
            # an empty tail (last column of a square matrix) gives the identity reflector: tau = +inf (householder.rs:73-79)
            assert np.all(fin[:-1]) and (fin[-1] or (m == n and j + b == n))
            assert np.allclose(np.diag(Tb)[fin], 0.5 * np.diag(G)[fin], rtol=1e-12)
            assert np.allclose(np.triu(Tb, 1), np.triu(G, 1), rtol=1e-10, atol=1e-12)


def test_norm_l2_scaling(oracle):
    """reference: test_norm_l2 (reductions/norm_l2.rs:179-219): no overflow/underflow at 1e+-250, rel 1e-14."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal(1000)
    for s in [1.0, 1e250, 1e-250, 1e-300]:
        got = oracle.norm_l2(x * s)
        want = float(np.linalg.norm(x)) * s
        assert abs(got - want) <= 1e-13 * want
    assert oracle.norm_l2(np.zeros(7)) == 0.0
    z = (rng.standard_normal(100) + 1j * rng.standard_normal(100))
    assert abs(oracle.norm_l2(z) - np.linalg.norm(z)) <= 1e-13 * np.linalg.norm(z)
