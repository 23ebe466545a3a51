"""`faer_b200_bidiag_in_place_{c64,c32}` / `faer_b200_tridiag_in_place_{c64,c32}` (extensions mirroring svd::bidiag::bidiag_in_place
and evd::tridiag::tridiag_in_place for complex T; csrc/cplx_condensed.cu) against the oracle's restatement of the reference
(test_bidiag_cplx svd/bidiag.rs:383-502, test_tridiag_cplx evd/tridiag.rs:597-660): condensed entries, reflectors and T blocks close
to the oracle's (the phases of late entries inherit the rounding of every earlier step: moduli are compared tightly, entries with a
looser bound), the reference's reconstruction identities through the block-Householder sequences exactly as its tests do them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CDTYPES = [np.complex128, np.complex64]


def crandn(rng, shape, dtype):
    return np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype))


def rdt(dtype):
    return np.float64 if dtype == np.complex128 else np.float32


@pytest.mark.parametrize("dtype", CDTYPES)
def test_cplx_bidiag_vs_oracle(fb, oracle, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(495)
    u = float(np.finfo(rdt(dtype)).eps)
    for (m, n, bl, br) in [(8, 4, 4, 3), (8, 8, 4, 3), (1, 1, 1, 1), (2, 2, 1, 1), (5, 1, 2, 1), (33, 17, 8, 8), (64, 64, 16, 5), (130, 97, 32, 32)]:
        A = crandn(rng, (m, n), dtype)
        want = A.astype(np.complex128).copy(order="F")
        Hl_w, Hr_w = oracle.bidiag(want, bl, br)
        got = A.copy(order="F" if m % 2 else "C")
        Hl = np.full((bl, n), np.nan, dtype=dtype, order="F"); Hr = np.full((br, max(n - 1, 0)), np.nan, dtype=dtype, order="F")
        la.bidiag_in_place(got, Hl, Hr)
        assert np.all(np.isfinite(got))
        scale = max(1.0, float(np.abs(A).max())) * max(m, n)
        d_g, d_w = np.diagonal(got), np.diagonal(want)
        assert np.abs(np.abs(d_g) - np.abs(d_w)).max() <= 256 * u * scale, (m, n)            # |B(k, k)|: phase-free
        assert np.abs(got - want).max() <= 8192 * u * scale, (m, n)
        # reconstruction as in the reference's test: U^H A V == B through the sequences with the returned T blocks
        W = A.copy(order="F")
        oracle.apply_q_transpose_sequence(np.asfortranarray(got[:, :n]), np.asfortranarray(Hl), W, conj_lhs=True)
        if n > 1:
            oracle.apply_q_transpose_sequence(np.asfortranarray(got[:n - 1, 1:n].T), np.asfortranarray(Hr), W[:, 1:n].T, conj_lhs=True)
        B = got.copy()
        i, j = np.indices(B.shape)
        B[(i > j) | (j > i + 1)] = 0
        assert np.abs(B - W).max() <= 256 * max(m, n) * u * max(1.0, np.abs(A).max()), (m, n)
        sv_a = np.linalg.svd(A.astype(np.complex128), compute_uv=False)
        sv_b = np.linalg.svd(B[:n, :n].astype(np.complex128), compute_uv=False)
        assert np.abs(sv_a - sv_b).max() <= 256 * max(m, n) * u * max(1.0, np.abs(A).max()), (m, n)


@pytest.mark.parametrize("dtype", CDTYPES)
def test_cplx_tridiag_vs_oracle(fb, oracle, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(496)
    u = float(np.finfo(rdt(dtype)).eps)
    for n, b in [(2, 3), (3, 3), (4, 3), (8, 3), (16, 3), (1, 1), (45, 8), (100, 32), (257, 16)]:
        G = crandn(rng, (n, n), np.complex128)
        A = np.asfortranarray((G + G.conj().T).astype(dtype))
        want = A.astype(np.complex128).copy(order="F")
        oracle.tridiag(want, b)
        got = A.copy(order="F")
        got[np.triu_indices(n, 1)] = np.nan                          # the strict upper triangle is neither read nor written
        H = np.full((b, max(n - 1, 0)), np.nan, dtype=dtype, order="F")
        la.tridiag_in_place(got, H)
        assert np.all(np.isnan(got[np.triu_indices(n, 1)])), n
        lo = np.tril_indices(n)
        assert np.all(np.isfinite(got[lo])) and np.all(np.diagonal(got).imag == 0), n
        scale = max(1.0, float(np.abs(A).max())) * n
        assert np.abs(np.diagonal(got) - np.diagonal(want)).max() <= 8192 * u * scale, n
        assert np.abs(np.abs(np.diagonal(got, -1)) - np.abs(np.diagonal(want, -1))).max(initial=0.0) <= 256 * u * scale, n
        # Q^H A Q == T exactly as the reference test does it (tridiag.rs:565-596)
        W = A.copy(order="F")
        if n > 1:
            Vs = np.asfortranarray(np.tril(got)[1:, :n - 1])
            oracle.apply_q_transpose_sequence(Vs, np.asfortranarray(H), W[1:, :], conj_lhs=True)
            oracle.apply_q_transpose_sequence(Vs, np.asfortranarray(H), W.T[1:, :], conj_lhs=False)
        Tm = np.zeros_like(A)
        for i in range(n):
            Tm[i, i] = got[i, i]
            if i + 1 < n:
                Tm[i + 1, i] = got[i + 1, i]; Tm[i, i + 1] = np.conj(got[i + 1, i])
        assert np.abs(Tm - W).max() <= 256 * n * u * max(1.0, np.abs(A).max()), n
