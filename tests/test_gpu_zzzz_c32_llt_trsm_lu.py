"""c32 (complex64) triangular solves, Cholesky LLT and partial-pivoting LU through the C ABI (csrc/cplx.cu instantiated for float)
against the oracle's c32 restatement of the same recursions (triangular_solve.rs:220-604, cholesky/llt/factor.rs:68-97,
lu/partial_pivoting/factor.rs:19-295 for complex T). Same checks as the c64 file with the unit roundoff of f32; LU permutations
are compared with the oracle's exactly where no pivot near-tie can occur (small n) and validated by the factorization otherwise
(f32 rounding differs between the GPU's and the oracle's summation orders, so a near-tie may legitimately resolve differently)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
U = 2.0 ** -24


def crandn(rng, shape):
    return np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64))


def test_c32_triangular_solves_vs_oracle(fb, oracle):
    la = fb.linalg
    rng = np.random.default_rng(151)
    for n, k in [(1, 1), (5, 3), (32, 7), (33, 40), (100, 64), (257, 130)]:
        T = (crandn(rng, (n, n)) + 4 * np.sqrt(n) * np.eye(n)).astype(np.complex64)
        B = crandn(rng, (n, k))
        for lower in (True, False):
            for unit in (False, True):
                for conj in (0, 1):
                    Tt = np.asfortranarray(np.tril(T) if lower else np.triu(T))
                    if unit:
                        Tt = np.asfortranarray((Tt / (2.0 * np.sqrt(n))).astype(np.complex64))
                    want = B.copy(order="F"); oracle.solve_triangular(Tt, want, lower, unit, bool(conj))
                    got = B.copy(order="F")
                    f = {(True, False): la.solve_lower_triangular_in_place, (False, False): la.solve_upper_triangular_in_place,
                         (True, True): la.solve_unit_lower_triangular_in_place, (False, True): la.solve_unit_upper_triangular_in_place}[(lower, unit)]
                    f(Tt, got, conj)
                    assert got.dtype == np.complex64
                    Te = ((np.tril(Tt, -1) if lower else np.triu(Tt, 1)) + (np.eye(n) if unit else np.diag(np.diag(Tt)))).astype(np.complex128)
                    Te = Te.conj() if conj else Te
                    g = got.astype(np.complex128)
                    res = np.abs(Te @ g - B)
                    bound = 16 * n * U * (np.abs(Te) @ np.abs(g) + np.abs(B)) + 1e-30
                    assert np.all(res <= bound), (n, k, lower, unit, conj, float((res / bound).max()))
                    assert np.allclose(got, want, rtol=2e-3, atol=2e-4 * np.abs(want).max()), (n, k, lower, unit, conj)


def test_c32_llt_vs_oracle(fb, oracle):
    la = fb.linalg
    rng = np.random.default_rng(152)
    for n in [1, 2, 31, 32, 33, 64, 100, 257, 600]:
        G = crandn(rng, (n, n)).astype(np.complex128)
        A = np.asfortranarray((G @ G.conj().T + n * np.eye(n)).astype(np.complex64))
        want = A.copy(order="F"); fail, cnt = oracle.llt(want)
        assert fail == -1
        got = A.copy(order="F")
        got[np.triu_indices(n, 1)] = 123.0 + 7j  # the strict upper triangle is neither read nor written
        info = la.cholesky_in_place(got)
        assert info.dynamic_regularization_count == 0
        assert np.all(got[np.triu_indices(n, 1)] == np.complex64(123.0 + 7j)), n
        L = np.tril(got).astype(np.complex128)
        A64 = A.astype(np.complex128)
        assert np.max(np.abs(L @ L.conj().T - A64)) <= 64 * n * U * np.max(np.abs(A64)), n
        assert np.allclose(got[np.tril_indices(n)], want[np.tril_indices(n)], rtol=1e-3, atol=1e-4 * np.abs(want).max()), n
        B = crandn(rng, (n, 3))
        for conj in (0, 1):
            X = B.copy(order="F"); la.llt_solve_in_place(got, X, conj)
            Ae = A64.conj() if conj else A64
            assert np.max(np.abs(Ae @ X.astype(np.complex128) - B)) <= 256 * n * U * np.linalg.cond(A64) * np.max(np.abs(B)), (n, conj)


def test_c32_llt_error_index_and_regularization(fb, oracle):
    la = fb.linalg
    rng = np.random.default_rng(153)
    n = 90
    G = crandn(rng, (n, n)).astype(np.complex128)
    A = np.asfortranarray((G @ G.conj().T + n * np.eye(n)).astype(np.complex64))
    bad = A.copy(order="F"); bad[57, 57] = -1.0
    want = bad.copy(order="F"); fail, _ = oracle.llt(want)
    assert fail >= 0
    with pytest.raises(la.LltError) as e:
        la.cholesky_in_place(bad.copy(order="F"))
    assert f"index: {fail}" in str(e.value)
    R = A.copy(order="F"); R[10, 10] = 1e-30; R[11:, 10] = 0; R[10, :10] = 0
    want = R.copy(order="F"); fail, cnt = oracle.llt(want, delta=1e-3, eps=1e-5)
    got = R.copy(order="F"); info = la.cholesky_in_place(got, regularization=(1e-3, 1e-5))
    assert fail == -1 and info.dynamic_regularization_count == cnt and cnt >= 1


def test_c32_lu_vs_oracle(fb, oracle):
    la = fb.linalg
    rng = np.random.default_rng(154)
    for (m, n) in [(1, 1), (7, 7), (16, 16), (17, 17), (40, 40), (130, 70), (70, 130), (300, 300), (800, 800)]:
        for idx in (np.uint32, np.uint64):
            A = crandn(rng, (m, n))
            want = A.copy(order="F"); po, pio, nt = oracle.lu(want)
            got = A.copy(order="F")
            p = np.zeros(m, dtype=idx); pi = np.zeros(m, dtype=idx)
            info = la.lu_in_place(got, p, pi)
            pp = p.astype(np.int64)
            assert sorted(pp.tolist()) == list(range(m)) and np.array_equal(pi.astype(np.int64)[pp], np.arange(m)), (m, n)
            if max(m, n) <= 40:
                assert np.array_equal(pp, po) and info.transposition_count == nt, (m, n)
                assert np.allclose(got, want, rtol=1e-3, atol=1e-4 * np.abs(want).max()), (m, n)
            # P A = L U and |l_ij| bounded by the abs1 pivot rule (|l| <= sqrt(2) up to rounding)
            s = min(m, n)
            g = got.astype(np.complex128)
            L = np.tril(g[:, :s], -1) + np.eye(m, s)
            Uf = np.triu(g[:s, :])
            PA = A.astype(np.complex128)[pp, :]
            assert np.max(np.abs(L @ Uf - PA)) <= 32 * s * U * np.max(np.abs(Uf)), (m, n)
            assert np.max(np.abs(L)) <= np.sqrt(2.0) * (1 + 1e-5), (m, n)
            if m == n:
                B = crandn(rng, (n, 3))
                for conj in (0, 1):
                    X = B.copy(order="F"); la.lu_solve_in_place(got, p, pi, X, conj)
                    Ae = A.astype(np.complex128); Ae = Ae.conj() if conj else Ae
                    assert np.max(np.abs(Ae @ X.astype(np.complex128) - B)) <= 256 * n * U * np.linalg.cond(Ae) * np.max(np.abs(B)), (n, conj)
                    X = B.copy(order="F"); la.lu_solve_transpose_in_place(got, p, pi, X, conj)  # solve.rs:55-86
                    assert np.max(np.abs(Ae.T @ X.astype(np.complex128) - B)) <= 256 * n * U * np.linalg.cond(Ae) * np.max(np.abs(B)), (n, conj, "T")
