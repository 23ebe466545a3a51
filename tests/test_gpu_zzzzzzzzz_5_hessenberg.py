"""`faer_b200_hessenberg_in_place_{f64,f32,c64,c32}` (extension mirroring evd::hessenberg::hessenberg_in_place,
evd/hessenberg.rs:549-567; csrc/cplx_condensed.cu on the launch sequence of cplx_condensed_core.cuh): the reference's own tests
(test_hessenberg_real / _cplx: n in {1, 2, 3, 4, 8, 16}, block size 3) restated and extended — Q^H A Q through the block-Householder
sequences with the returned T blocks equals the Hessenberg part; the eigenvalues of H are those of A."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32, np.complex128, np.complex64]


def rdt(dtype):
    return np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64


@pytest.mark.parametrize("dtype", DTYPES)
def test_hessenberg(fb, oracle, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(491)
    u = float(np.finfo(rdt(dtype)).eps)
    for n in [1, 2, 3, 4, 8, 16, 45, 130]:
        for bs in (3, 32):
            G = rng.standard_normal((n, n))
            if np.issubdtype(dtype, np.complexfloating):
                G = G + 1j * rng.standard_normal((n, n))
            A = np.asfortranarray(G.astype(dtype))
            W = A.copy(order="F" if bs == 3 else "C")                       # both layouts
            Hf = np.full((bs, max(n - 1, 0)), np.nan, dtype=dtype, order="F")
            la.hessenberg_in_place(W, Hf)
            assert np.all(np.isfinite(W))
            H = np.triu(W, -1)
            if n > 1:
                V = np.asfortranarray(W[1:, :n - 1])
                B = A.copy(order="F")
                oracle.apply_q_transpose_sequence(V, Hf, B[1:, :], conj_lhs=True)          # Q^H A   (hessenberg.rs tests)
                oracle.apply_q_transpose_sequence(V, Hf, B.T[1:, :], conj_lhs=False)       # (Q^H A) Q
                assert np.abs(B - H).max() <= 256 * n * u * np.abs(A).max(), (n, bs)
            else:
                assert np.array_equal(W, A)
            if rdt(dtype) == np.float64 and n <= 16:                        # (eigenvalues of a non-normal matrix: small n only)
                ev_a = np.linalg.eigvals(A.astype(np.complex128))
                ev_h = list(np.linalg.eigvals(H.astype(np.complex128)))
                for ev in ev_a:                                             # greedy matching: conjugate pairs have no stable sort order
                    k = int(np.argmin([abs(ev - x) for x in ev_h]))
                    assert abs(ev - ev_h.pop(k)) <= 1e-8 * max(1.0, np.abs(ev_a).max()), (n, bs)
