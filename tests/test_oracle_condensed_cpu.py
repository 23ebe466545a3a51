"""CPU tests pinning the oracle's bidiagonalization / tridiagonalization (test infrastructure) to the reference's tests:
  test_bidiag_real / test_bidiag_cplx    faer/src/linalg/svd/bidiag.rs:383-502   ((8,4), (8,8); bl = 4, br = 3; ApproxEq eps)
  test_tridiag_real / test_tridiag_cplx  faer/src/linalg/evd/tridiag.rs:537-660  (n in 2,3,4,8,16; b = 3)
restated with our seeded inputs (plus larger shapes), and cross-checked against LAPACK-free invariants: the singular
values of B equal those of A, the eigenvalues of T equal those of A.
"""
import numpy as np
import pytest


def randn(rng, shape, dtype):
    if np.dtype(dtype).kind == "c":
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


def check_bidiag(oracle, A, UV, Hl, Hr, tol):
    """U^H A V == B, exactly as the reference test does it (bidiag.rs:405-440)."""
    m, n = A.shape
    size = min(m, n)
    W = A.copy(order="F")
    oracle.apply_q_transpose_sequence(UV[:, :size], Hl, W, conj_lhs=True)
    if size > 1:
        V = UV[:size - 1, 1:size]
        oracle.apply_q_transpose_sequence(V.T, Hr, W[:, 1:size].T, conj_lhs=True)
    B = UV.copy()
    i, j = np.indices(B.shape)
    B[(i > j) | (j > i + 1)] = 0
    scale = max(1.0, float(np.abs(A).max())) * max(m, n)
    assert np.abs(B - W).max() <= tol * scale, np.abs(B - W).max()
    return B


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32])
def test_bidiag_reconstruction(oracle, dtype):
    rng = np.random.default_rng(0)
    tol = 1e-5 if dtype == np.float32 else 1e-14
    for (m, n, bl, br) in [(8, 4, 4, 3), (8, 8, 4, 3), (1, 1, 1, 1), (2, 2, 1, 1), (5, 1, 2, 1), (33, 17, 8, 8), (64, 64, 16, 5),
                           (130, 97, 32, 32)]:
        A = np.asfortranarray(randn(rng, (m, n), dtype))
        UV = A.copy(order="F")
        Hl, Hr = oracle.bidiag(UV, bl, br)
        B = check_bidiag(oracle, A, UV, Hl, Hr, tol)
        sv = np.linalg.svd(A.astype(np.complex128 if np.dtype(dtype).kind == "c" else np.float64), compute_uv=False)
        sb = np.linalg.svd(B[:n, :n].astype(sv.dtype if np.dtype(dtype).kind != "c" else np.complex128), compute_uv=False)
        assert np.abs(sv - sb).max() <= tol * 50 * max(1.0, sv.max())
        if np.dtype(dtype).kind != "c":
            assert np.all(np.isfinite(B))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_tridiag_reconstruction(oracle, dtype):
    rng = np.random.default_rng(1)
    for n, b in [(2, 3), (3, 3), (4, 3), (8, 3), (16, 3), (1, 1), (45, 8), (100, 32)]:
        G = randn(rng, (n, n), dtype)
        A = np.asfortranarray(G + G.conj().T)
        V = A.copy(order="F")
        # poison the strict upper triangle: the routine reads/writes the lower triangle only (tridiag.rs:266-273)
        V[np.triu_indices(n, 1)] = np.nan
        H = oracle.tridiag(V, b)
        assert np.all(np.isnan(V[np.triu_indices(n, 1)]))
        W = A.copy(order="F")
        if n > 1:
            Vs = V[1:, :n - 1]
            oracle.apply_q_transpose_sequence(Vs, H, W[1:, :], conj_lhs=True)       # Q^H A
            oracle.apply_q_transpose_sequence(Vs, H, W.T[1:, :], conj_lhs=False)    # (Q^H A) Q, tridiag.rs:565-585
        T = np.zeros_like(A)
        for i in range(n):
            T[i, i] = V[i, i]
            if i + 1 < n:
                T[i + 1, i] = V[i + 1, i]
                T[i, i + 1] = np.conj(V[i + 1, i]) if iscomplex(dtype) else V[i + 1, i]
        assert np.abs(T - W).max() <= 1e-13 * max(1.0, np.abs(A).max()) * n, (n, np.abs(T - W).max())
        ev = np.linalg.eigvalsh(A)
        et = np.linalg.eigvalsh(T)
        assert np.abs(ev - et).max() <= 1e-12 * max(1.0, np.abs(ev).max())


def iscomplex(dtype):
    return np.dtype(dtype).kind == "c"
