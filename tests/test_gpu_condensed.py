"""GPU parity tests for the reductions to condensed form against the oracle.

Reference tests restated: svd/bidiag.rs:383-502 (U^H A V == B within ApproxEq eps, shapes (8,4), (8,8), bl = 4, br = 3).
Tolerances (written here). What is pinned tightly is what the reference itself tests: the
reconstruction identity U^H A V == B (resp. Q^H A Q == T) to 64 * n * eps * max|A|, plus the invariants (singular values /
eigenvalues) to the same bound. The ENTRIES of a condensed form are not forward-stable functions of A (two backward-stable
Householder reductions with different summation orders drift apart like a Lanczos recurrence: measured 1e-13 at n = 64,
2e-10 at n = 300, 3e-10 at n = 700 in f64), so the elementwise comparison with the oracle uses 1024 * n * eps * max|A| for
n <= 64 and a drift bound n^2.5 * eps * 16 * max|A| beyond (f64; it still catches sign / convention / indexing errors,
which are O(1)); f32 compares entries only for n <= 64.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reconstruct(oracle, A, UV, Hl, Hr):
    m, n = A.shape
    W = A.copy(order="F")
    oracle.apply_q_transpose_sequence(UV[:, :n], Hl, W, conj_lhs=True)
    if n > 1:
        oracle.apply_q_transpose_sequence(UV[:n - 1, 1:n].T, Hr, W[:, 1:n].T, conj_lhs=True)
    B = UV.copy()
    i, j = np.indices(B.shape)
    B[(i > j) | (j > i + 1)] = 0
    return B, W


def _entry_tol(n, dtype, A):
    """Elementwise bound against the oracle (see the module docstring); None = entries not compared."""
    eps = np.finfo(dtype).eps
    amax = max(1.0, float(np.abs(A).max()))
    if n <= 64:
        return 1024 * n * eps * amax
    if dtype == np.float32:
        return None
    return 16 * n ** 2.5 * eps * amax


def _close_with_inf(got, want, tol, what):
    """Elementwise |got - want| <= tol * max(1, max finite |want|); infinities (tau of an identity reflector) must coincide."""
    assert np.array_equal(np.isinf(got), np.isinf(want)), what
    fin = np.isfinite(want)
    assert np.all(np.isfinite(got[fin])), what
    if fin.any():
        assert np.abs(got[fin] - want[fin]).max() <= tol * max(1.0, np.abs(want[fin]).max()), what


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bidiag_vs_oracle(fb, oracle, dtype):
    la = fb.linalg
    rng = np.random.default_rng(7)
    eps = np.finfo(dtype).eps
    for (m, n, bl, br) in [(8, 4, 4, 3), (8, 8, 4, 3), (1, 1, 1, 1), (2, 2, 1, 1), (5, 1, 2, 1), (33, 17, 8, 8), (64, 64, 16, 5),
                           (130, 97, 32, 32), (300, 300, 32, 16), (1000, 37, 8, 8), (513, 512, 64, 64)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        want = A.copy(order="F")
        Hl_w, Hr_w = oracle.bidiag(want, bl, br)
        got = A.copy(order="F")
        Hl = np.zeros((bl, n), dtype=dtype, order="F")
        Hr = np.zeros((br, max(n - 1, 0)), dtype=dtype, order="F")
        la.bidiag_in_place(got, Hl, Hr)
        tol = 64 * max(m, n) * eps * max(1.0, np.abs(A).max())
        etol = _entry_tol(max(m, n), dtype, A)
        assert np.all(np.isfinite(got)), (m, n)
        if etol is not None:
            assert np.abs(got - want).max() <= etol, (m, n, np.abs(got - want).max(), etol)
            # T blocks: diag = tau, strict upper = V^H V; entries below the diagonal of each block are untouched (zero)
            _close_with_inf(Hl, Hl_w, etol, (m, n, "Hl"))
            if n > 1:
                _close_with_inf(Hr, Hr_w, etol, (m, n, "Hr"))  # tau = +inf for a reflector with an empty tail
        else:
            assert np.array_equal(np.isinf(Hr), np.isinf(Hr_w)) and np.array_equal(np.isinf(Hl), np.isinf(Hl_w))
        B, W = _reconstruct(oracle, A, got, Hl, Hr)
        assert np.abs(B - W).max() <= tol, (m, n, np.abs(B - W).max())
        sv_a = np.linalg.svd(A.astype(np.float64), compute_uv=False)
        sv_b = np.linalg.svd(B[:n, :n].astype(np.float64), compute_uv=False)
        assert np.abs(sv_a - sv_b).max() <= tol, (m, n, np.abs(sv_a - sv_b).max())


def test_bidiag_device_resident_singular_values(fb, cuda_dev):
    """Size-independent property at a size the oracle would take minutes for: the singular values of B equal those of A
    (orthogonal equivalence), checked through LAPACK on the host; and the call works in place on device memory."""
    import torch
    la = fb.linalg
    m, n, bl, br = 3000, 2048, 64, 64
    g = torch.Generator(device="cpu"); g.manual_seed(3)
    A = torch.randn((n, m), dtype=torch.float64, generator=g).to(cuda_dev).T  # column-major m x n
    A0 = A.cpu().numpy().copy()
    Hl = torch.zeros((n, bl), dtype=torch.float64, device=cuda_dev).T
    Hr = torch.zeros((n - 1, br), dtype=torch.float64, device=cuda_dev).T
    la.bidiag_in_place(A, Hl, Hr)
    R = A.cpu().numpy()
    d = np.diagonal(R).copy(); e = np.diagonal(R, 1).copy()
    B = np.diag(d) + np.diag(e, 1)
    sv_a = np.linalg.svd(A0, compute_uv=False)
    sv_b = np.linalg.svd(B, compute_uv=False)
    assert np.abs(sv_a - sv_b).max() <= 1e-11 * sv_a.max()
    # reflector scaling: every tau = (1 + |v|^2)/2 >= 1/2, and v is bounded by construction (|v_i| <= 1 for Householder with
    # beta = -sign(head) * norm)
    Hn = Hl.cpu().numpy()
    taus = Hn[np.arange(n) % bl, np.arange(n)]  # the diagonal of every T block
    assert np.all(taus >= 0.5)
    assert np.abs(np.tril(R, -1)).max() <= 1.0 + 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tridiag_vs_oracle(fb, oracle, dtype):
    """evd/tridiag.rs:537-596 restated (A + A^H, b = 3, sizes 2..16) plus larger sizes; the strict upper triangle is poisoned
    with NaN: the routine must neither read nor write it."""
    la = fb.linalg
    rng = np.random.default_rng(8)
    eps = np.finfo(dtype).eps
    for n, b in [(2, 3), (3, 3), (4, 3), (8, 3), (16, 3), (1, 1), (45, 8), (100, 32), (257, 16), (700, 64)]:
        Gm = rng.standard_normal((n, n)).astype(dtype)
        A = np.asfortranarray(Gm + Gm.T)
        want = A.copy(order="F")
        H_w = oracle.tridiag(want, b)
        got = A.copy(order="F")
        got[np.triu_indices(n, 1)] = np.nan
        H = np.zeros((b, max(n - 1, 0)), dtype=dtype, order="F")
        la.tridiag_in_place(got, H)
        assert np.all(np.isnan(got[np.triu_indices(n, 1)])), n
        tol = 64 * n * eps * max(1.0, np.abs(A).max())
        etol = _entry_tol(n, dtype, A)
        lo = np.tril_indices(n)
        assert np.all(np.isfinite(got[lo])), n
        if etol is not None:
            assert np.abs(got[lo] - want[lo]).max() <= etol, (n, np.abs(got[lo] - want[lo]).max(), etol)
            if n > 1:
                _close_with_inf(H, H_w, etol, (n, "H"))
        # reconstruction Q^H A Q == T exactly as the reference test does it (tridiag.rs:565-596), with the oracle's
        # reflector application as the checker
        W = A.copy(order="F")
        if n > 1:
            Vs = np.tril(got)[1:, :n - 1]
            oracle.apply_q_transpose_sequence(Vs, H, W[1:, :], conj_lhs=True)
            oracle.apply_q_transpose_sequence(Vs, H, W.T[1:, :], conj_lhs=False)
        Tm = np.zeros_like(A)
        for i in range(n):
            Tm[i, i] = got[i, i]
            if i + 1 < n:
                Tm[i + 1, i] = Tm[i, i + 1] = got[i + 1, i]
        assert np.abs(Tm - W).max() <= tol, (n, np.abs(Tm - W).max(), tol)
        # eigenvalues of T == eigenvalues of A
        d = np.diagonal(got).astype(np.float64); e = np.diagonal(got, -1).astype(np.float64)
        T = np.diag(d) + np.diag(e, -1) + np.diag(e, 1)
        ev_a = np.linalg.eigvalsh(A.astype(np.float64)); ev_t = np.linalg.eigvalsh(T)
        assert np.abs(ev_a - ev_t).max() <= 50 * tol, (n, np.abs(ev_a - ev_t).max())


def test_tridiag_device_resident_eigenvalues(fb, cuda_dev):
    import torch
    la = fb.linalg
    n, b = 2500, 64
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    Gm = torch.randn((n, n), dtype=torch.float64, generator=g)
    A0 = (Gm + Gm.T).numpy()
    A = torch.from_numpy(A0).to(cuda_dev).T.contiguous().T  # column-major, symmetric
    H = torch.zeros((n - 1, b), dtype=torch.float64, device=cuda_dev).T
    la.tridiag_in_place(A, H)
    R = A.cpu().numpy()
    d = np.diagonal(R); e = np.diagonal(R, -1)
    T = np.diag(d) + np.diag(e, -1) + np.diag(e, 1)
    ev_a = np.linalg.eigvalsh(A0); ev_t = np.linalg.eigvalsh(T)
    assert np.abs(ev_a - ev_t).max() <= 1e-11 * np.abs(ev_a).max()
