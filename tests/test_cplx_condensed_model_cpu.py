"""The host model of the complex condensed-form drivers (tests/cplx_condensed_model.py, the statement-by-statement model of
csrc/cplx_condensed.cu) against the oracle's restatement of the reference's fused algorithms: the textbook two-sided
updates with the reference's reflectors give the same reflectors, taus and condensed entries as evd/tridiag.rs:274-529 and
svd/bidiag.rs:47-256 up to rounding, and the whole pipelines satisfy the reference's own test identities
(evd/mod.rs tests: A = U S U^H, U unitary; svd/mod.rs:780-783: ||A - U S V^H|| with eps * 128 * sqrt(8 n))."""
import numpy as np
import pytest

import cplx_condensed_model as cm

U = np.finfo(np.float64).eps


def tau_close(a, b, n):
    if np.isinf(a) or np.isinf(b):
        return np.isinf(a) and np.isinf(b)
    return abs(a - b) <= 4096 * n * U * abs(b)


def crandn(rng, shape):
    return np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


@pytest.mark.parametrize("n", [1, 2, 3, 8, 33, 100])
def test_tridiag_model_matches_the_oracle(oracle, n):
    rng = np.random.default_rng(500 + n)
    G = crandn(rng, (n, n))
    A = np.asfortranarray(G + G.conj().T)
    W, taus = cm.tridiag_unblocked(A)
    Ao = A.copy(order="F")
    H = oracle.tridiag(Ao, 1)
    scale = np.abs(A).max() * n
    for k in range(n):
        assert abs(W[k, k] - Ao[k, k]) <= 4096 * U * scale
        if k + 1 < n:
            assert abs(W[k + 1, k] - Ao[k + 1, k]) <= 4096 * U * scale                       # complex subdiagonal (beta)
            assert abs(abs(W[k + 1, k]) - abs(Ao[k + 1, k])) <= 64 * U * scale
            assert np.abs(W[k + 2:, k] - Ao[k + 2:, k]).max(initial=0.0) <= 4096 * n * U      # essentials (|v_i| <= 1)
            assert tau_close(taus[k], H[0, k].real, n)


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (5, 3), (8, 8), (40, 17), (64, 64), (100, 30)])
def test_bidiag_model_matches_the_oracle(oracle, shape):
    m, n = shape
    rng = np.random.default_rng(600 + m + n)
    A = crandn(rng, (m, n))
    W, tl, tr = cm.bidiag_unblocked(A)
    Ao = A.copy(order="F")
    Hl, Hr = oracle.bidiag(Ao, 1, 1)
    scale = np.abs(A).max() * max(m, n)
    for k in range(n):
        assert abs(W[k, k] - Ao[k, k]) <= 4096 * U * scale and abs(abs(W[k, k]) - abs(Ao[k, k])) <= 64 * U * scale
        assert np.abs(W[k + 1:, k] - Ao[k + 1:, k]).max(initial=0.0) <= 4096 * max(m, n) * U
        assert tau_close(tl[k], Hl[0, k].real, max(m, n))
        if k + 1 < n:
            assert abs(W[k, k + 1] - Ao[k, k + 1]) <= 4096 * U * scale and abs(abs(W[k, k + 1]) - abs(Ao[k, k + 1])) <= 64 * U * scale
            assert np.abs(W[k, k + 2:] - Ao[k, k + 2:]).max(initial=0.0) <= 4096 * max(m, n) * U
            assert tau_close(tr[k], Hr[0, k].real, max(m, n))


@pytest.mark.parametrize("n", [1, 2, 5, 32, 97])
def test_self_adjoint_evd_model_identities(n):
    rng = np.random.default_rng(700 + n)
    G = crandn(rng, (n, n))
    A = G + G.conj().T
    lam, Um = cm.self_adjoint_evd(A)
    assert np.all(np.diff(lam) >= 0)
    assert np.abs(Um.conj().T @ Um - np.eye(n)).max() <= 64 * n * U
    assert np.abs(Um @ np.diag(lam) @ Um.conj().T - A).max() <= 64 * n * U * np.abs(A).max()
    assert np.abs(lam - np.linalg.eigvalsh(A)).max() <= 64 * n * U * np.abs(A).max()
    # a matrix with zero sub-columns: reflectors with tau = inf, zero subdiagonal entries (phase 1)
    if n >= 5:
        D = np.diag(rng.standard_normal(n)).astype(np.complex128)
        D[2, 1] = D[1, 2] = 0.0
        lam, Um = cm.self_adjoint_evd(D)
        assert np.abs(Um @ np.diag(lam) @ Um.conj().T - D).max() <= 64 * n * U * np.abs(D).max()


@pytest.mark.parametrize("shape", [(1, 1), (3, 3), (10, 4), (4, 10), (33, 33), (60, 20), (20, 60)])
@pytest.mark.parametrize("full", [False, True])
def test_svd_model_identities(shape, full):
    m, n = shape
    size = min(m, n)
    rng = np.random.default_rng(800 + m * 3 + n)
    A = crandn(rng, (m, n))
    S, Um, Vm = cm.svd(A, full=full)
    assert Um.shape == (m, m if full else size) and Vm.shape == (n, n if full else size)
    assert np.all(np.diff(S) <= 0) and np.all(S >= 0)
    tol = U * 128 * np.sqrt(8 * max(m, n)) * max(1.0, np.abs(A).max())
    assert np.abs(Um[:, :size] @ np.diag(S) @ Vm[:, :size].conj().T - A).max() <= tol
    assert np.abs(Um.conj().T @ Um - np.eye(Um.shape[1])).max() <= tol
    assert np.abs(Vm.conj().T @ Vm - np.eye(Vm.shape[1])).max() <= tol
    assert np.abs(S - np.linalg.svd(A, compute_uv=False)).max() <= tol


@pytest.mark.parametrize("conj", [False, True])
def test_back_transform_convention_matches_the_block_householder_sequence(oracle, conj):
    """The drivers hand the back-transforms to `apply_block_householder_sequence_on_the_left_in_place_with_conj`
    (householder.rs:724-765) with block size 1 — basis = the reduced matrix, factor = the row of taus, Conj::No for U and for the
    eigenvectors, Conj::Yes on the transposed rows for V (svd/mod.rs:403-429). The model's `apply_sequence` is that call: checked
    here against the oracle's restatement of the sequence (which the GPU's complex sequence is tested against on hardware)."""
    rng = np.random.default_rng(1400 + int(conj))
    m, n, k = 23, 9, 5
    A = crandn(rng, (m, n))
    W, tl, tr = cm.bidiag_unblocked(A)
    M = crandn(rng, (m, k))
    want = M.copy(order="F")
    H = np.asfortranarray(tl.astype(np.complex128)[None, :])           # 1 x n factor: the taus are the 1 x 1 T blocks
    oracle.apply_q_sequence(np.asfortranarray(W), H, want, conj_lhs=conj)
    got = M.copy()
    cm.apply_sequence(W, tl, got, conj=conj)
    assert np.abs(got - want).max() <= 64 * m * U * np.abs(M).max()
    # the right reflectors through the transposed corner, rows 1.. (the V back-transform)
    Wt = np.asfortranarray(W[:n, :].T)
    Mv = crandn(rng, (n, n))
    want = Mv.copy(order="F")
    Hr = np.asfortranarray(tr.astype(np.complex128)[None, :])
    oracle.apply_q_sequence(np.asfortranarray(Wt[1:, :n - 1]), Hr, want[1:, :], conj_lhs=conj)
    got = Mv.copy()
    cm.apply_sequence(Wt[1:, :n - 1], tr, got[1:, :], conj=conj)
    assert np.abs(got - want).max() <= 64 * n * U * np.abs(Mv).max()
