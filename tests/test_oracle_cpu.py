"""CPU tests that pin the oracle (test infrastructure) before anything is compared against it.

They restate the reference's own unit tests with our seeded inputs (rand 0.9's StdRng stream cannot be
reproduced without Rust) and the reference's tolerances:
  test_matmul      faer/src/linalg/matmul/mod.rs:1758-2061   (shapes, strides, conj, alpha/beta; abs 1e-3 c32)
  test_triangular  faer/src/linalg/matmul/mod.rs:2106-2266   (7^3 structures; 1e-10; untouched dst preserved)
  test_cholesky    faer/src/linalg/cholesky/ldlt/factor.rs:776-868 (n=0..64 leaf, {2,4,8,31,127,240}; 1e-12)
  test_plu         faer/src/linalg/lu/partial_pivoting/factor.rs:304-404 (1e-13)
  solve tests      faer/src/linalg/triangular_solve.rs / cholesky/llt/solve.rs:55-... (eps*128*8n)
and cross-check against LAPACK (scipy) where the algorithms coincide (LU pivot rule, Cholesky factor).
"""
import itertools

import numpy as np
import pytest
import scipy.linalg as sla

S_RECT, S_LOW, S_UP, S_SLOW, S_SUP, S_ULOW, S_UUP = range(7)


def approx_eq(a, b, abs_tol, rel_tol):
    """faer's ApproxEq (faer/src/utils/approx.rs:48-57): |a-b| <= abs_tol or <= rel_tol*max(|a|,|b|), elementwise."""
    d = np.abs(a - b)
    ok = (d <= abs_tol) | (d <= rel_tol * np.maximum(np.abs(a), np.abs(b)))
    return bool(np.all(ok))


def randn(rng, shape, dtype):
    if np.dtype(dtype).kind == "c":
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
    return rng.standard_normal(shape).astype(dtype)


def mask(a, s):
    a = a.copy()
    if s == S_RECT:
        return a
    a = np.tril(a) if s in (S_LOW, S_SLOW, S_ULOW) else np.triu(a)
    if s in (S_SLOW, S_SUP):
        np.fill_diagonal(a, 0)
    if s in (S_ULOW, S_UUP):
        np.fill_diagonal(a, 1)
    return a


def dst_select(n, s):
    if s == S_RECT:
        return np.ones((n, n), bool)
    sel = np.tril(np.ones((n, n), bool)) if s in (S_LOW, S_SLOW, S_ULOW) else np.triu(np.ones((n, n), bool))
    if s >= S_SLOW:
        np.fill_diagonal(sel, False)
    return sel


# shapes of the reference's test_matmul (matmul/mod.rs:1783-1802)
MATMUL_SHAPES = [(2, 2, 2), (4, 4, 4), (8, 8, 8), (16, 16, 16), (127, 127, 127), (128, 128, 128), (129, 129, 129),
                 (15, 15, 15), (17, 17, 17), (1, 1, 1), (1, 16, 16), (16, 1, 16), (16, 16, 1), (0, 4, 4), (4, 0, 4),
                 (4, 4, 0), (63, 9, 100), (100, 63, 9)]


@pytest.mark.parametrize("dtype", [np.float64, np.complex64, np.float32, np.complex128])
def test_matmul_shapes_strides_conj(oracle, dtype):
    rng = np.random.default_rng(0)
    tol = 1e-3 if np.dtype(dtype).itemsize in (4, 8) and np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else 1e-10
    for (m, n, k) in MATMUL_SHAPES:
        for layout in itertools.product("CF", repeat=3):
            for rev in [(1, 1, 1), (-1, 1, -1), (1, -1, 1)]:
                for conj_a, conj_b in [(False, False), (True, False), (True, True)]:
                    for add, alpha in [(False, 1.0), (True, -1.0), (True, 0.5)]:
                        A = np.array(randn(rng, (m, k), dtype), order=layout[0])[::rev[0], ::rev[1]]
                        B = np.array(randn(rng, (k, n), dtype), order=layout[1])[::rev[1], ::rev[2]]
                        Cm = np.array(randn(rng, (m, n), dtype), order=layout[2])[::rev[0], ::rev[2]]
                        want = alpha * ((A.conj() if conj_a else A) @ (B.conj() if conj_b else B))
                        if add:
                            want = Cm + want
                        got = Cm.copy(order="K") if False else Cm
                        oracle.matmul(got, add, A, B, alpha, conj_a, conj_b)
                        assert np.all(np.abs(got - want) <= tol), (m, n, k, layout, rev, conj_a, conj_b, add)


def test_matmul_replace_ignores_nan_dst(oracle):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((9, 5)); B = rng.standard_normal((5, 7))
    Cm = np.full((9, 7), np.nan)
    oracle.matmul(Cm, False, A, B, 2.0)
    assert np.allclose(Cm, 2.0 * A @ B, rtol=0, atol=1e-13)
    # K == 0: Replace zero-fills, Add is a no-op (matmul/mod.rs:1193-1198)
    C0 = np.full((3, 4), np.nan)
    oracle.matmul(C0, False, np.zeros((3, 0)), np.zeros((0, 4)), 1.0)
    assert np.all(C0 == 0)
    C1 = np.ones((3, 4))
    oracle.matmul(C1, True, np.zeros((3, 0)), np.zeros((0, 4)), 1.0)
    assert np.all(C1 == 1)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_triangular_all_structures(oracle, dtype):
    """reference: test_triangular / run_test_problem (matmul/mod.rs:2106-2266), tol 1e-10."""
    rng = np.random.default_rng(2)
    for ds, ls, rs in itertools.product(range(7), repeat=3):
        n = int(rng.integers(1, 100))
        m = n
        k = n if (ls != S_RECT or rs != S_RECT) else int(rng.integers(1, 100))
        if ds == S_RECT and ls != S_RECT and rs == S_RECT:
            pass
        nn = n if (ds != S_RECT or rs != S_RECT) else int(rng.integers(1, 100))
        if ds != S_RECT:
            m = nn = n
        if ls != S_RECT:
            m = k = n
            if ds != S_RECT:
                nn = n
        if rs != S_RECT:
            k = n
            nn = n
            if ls == S_RECT and ds == S_RECT:
                m = int(rng.integers(1, 100))
        A = np.asfortranarray(randn(rng, (m, k), dtype))
        B = np.asfortranarray(randn(rng, (k, nn), dtype))
        C0 = np.asfortranarray(randn(rng, (m, nn), dtype))
        for add in (False, True):
            Cm = C0.copy(order="F")
            alpha = 2.5
            full = alpha * (mask(A, ls) @ mask(B, rs))
            if add:
                full = C0 + full
            sel = dst_select(m, ds) if ds != S_RECT else np.ones((m, nn), bool)
            want = np.where(sel, full, C0)
            oracle.matmul_triangular(Cm, ds, add, A, ls, B, rs, alpha)
            assert approx_eq(Cm, want, 1e-10, 1e-10), (ds, ls, rs, m, nn, k, add)
            assert np.array_equal(Cm[~sel], C0[~sel])


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_triangular_solve(oracle, dtype):
    rng = np.random.default_rng(3)
    eps = np.finfo(np.float64).eps
    for n, k in [(0, 3), (1, 1), (2, 3), (3, 2), (4, 5), (5, 5), (17, 3), (50, 20), (129, 70), (200, 65)]:
        T = np.asfortranarray(randn(rng, (n, n), dtype) / max(n, 1) + 2 * np.eye(n))  # well conditioned, also with unit diag
        for lower, unit, conj in itertools.product((True, False), (True, False), (True, False)):
            Tm = np.tril(T) if lower else np.triu(T)
            if unit:
                np.fill_diagonal(Tm, 1)
            if conj:
                Tm = Tm.conj()
            Bm = np.asfortranarray(randn(rng, (n, k), dtype))
            X = Bm.copy(order="F")
            oracle.solve_triangular(T, X, lower, unit, conj)
            tol = eps * 128 * 8 * max(n, 1) * max(1.0, float(np.max(np.abs(Bm))) if Bm.size else 1.0) * 10
            assert np.all(np.abs(Tm @ X - Bm) <= tol), (n, k, lower, unit, conj)
            if n and not conj:
                ref = sla.solve_triangular(Tm, Bm, lower=lower)
                assert np.allclose(X, ref, rtol=1e-9, atol=1e-9)


def spd(rng, n, dtype):
    G = randn(rng, (n, n), dtype)
    return np.asfortranarray(G @ G.conj().T + n * np.eye(n))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_llt_leaf_sizes(oracle, dtype):
    """reference: test_simd_cholesky, n = 0..=64 (ldlt/factor.rs:776-817), tol 1e-12 (relative)."""
    rng = np.random.default_rng(4)
    for n in range(0, 65):
        A = spd(rng, n, dtype)
        L = A.copy(order="F")
        fail, cnt = oracle.llt(L)
        assert fail == -1 and cnt == 0
        Lt = np.tril(L)
        assert approx_eq(Lt @ Lt.conj().T, A, 1e-12, 1e-12), n
        assert np.array_equal(np.triu(L, 1), np.triu(A, 1))  # strict upper triangle untouched


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_llt_blocked(oracle, dtype):
    """reference: test_cholesky n in {2,4,8,31,127,240} with thresholds 32/32 (ldlt/factor.rs:819-868)."""
    rng = np.random.default_rng(5)
    for n in [2, 4, 8, 31, 127, 240]:
        for rt, bs in [(32, 32), (64, 128), (2, 4)]:
            A = spd(rng, n, dtype)
            L = A.copy(order="F")
            fail, _ = oracle.llt(L, recursion_threshold=rt, block_size=bs)
            assert fail == -1
            Lt = np.tril(L)
            assert approx_eq(Lt @ Lt.conj().T, A, 1e-12, 1e-12), (n, rt, bs)
            if np.dtype(dtype).kind == "f":
                assert np.allclose(Lt, np.linalg.cholesky(A), rtol=1e-11, atol=1e-11)


def test_llt_error_index_and_regularization(oracle):
    rng = np.random.default_rng(6)
    for n, bad in [(10, 0), (10, 9), (200, 150), (300, 64), (300, 255)]:
        A = spd(rng, n, np.float64)
        A[bad, bad] = -1.0
        L = A.copy(order="F")
        fail, _ = oracle.llt(L)
        assert fail == bad
    # NaN pivot -> !(d > 0)
    A = spd(rng, 20, np.float64); A[7, 7] = np.nan
    assert oracle.llt(A.copy(order="F"))[0] == 7
    # dynamic regularisation: pivots <= eps are replaced by delta and counted (ldlt/factor.rs:122-144)
    A = np.asfortranarray(np.diag([4.0, 1e-20, 9.0, -3.0]))
    L = A.copy(order="F")
    fail, cnt = oracle.llt(L, delta=1.0, eps=1e-10)
    assert fail == -1 and cnt == 2
    # reference quirk (ldlt/factor.rs:161-175): the stored diagonal is the UNregularised a_jj times recip(sqrt(delta))
    assert np.allclose(np.diag(L), [2.0, 1e-20, 3.0, -3.0])


def unpack_lu(LU):
    m, n = LU.shape
    size = min(m, n)
    L = np.tril(LU[:, :size], -1) + np.eye(m, size)
    U = np.triu(LU[:size, :])
    return L, U


def test_plu(oracle):
    """reference: test_plu (lu/partial_pivoting/factor.rs:304-404): P^-1 L U ~ A at 1e-13, recursion_threshold 2."""
    rng = np.random.default_rng(7)
    for n in [1, 2, 3, 128, 255, 256, 257]:
        A = np.asfortranarray(rng.standard_normal((n, n)))
        for rt in (2, 16):
            LU = A.copy(order="F")
            perm, perm_inv, nt = oracle.lu(LU, recursion_threshold=rt)
            L, U = unpack_lu(LU)
            assert approx_eq((L @ U), A[perm, :], 1e-13 * max(1, n / 16), 1e-13 * max(1, n / 16)), n
            assert np.array_equal(perm_inv[perm], np.arange(n))
            # LAPACK uses the same first-max |x| rule for reals -> identical pivots on Gaussian input
            _, piv = sla.lu_factor(A)
            p = np.arange(n)
            for i, pi in enumerate(piv):
                p[i], p[pi] = p[pi], p[i]
            assert np.array_equal(p, perm), n
            assert nt == int(np.sum(piv != np.arange(n)))
    for m in [8, 128, 255, 256, 257]:
        A = np.asfortranarray(rng.standard_normal((m, 8)))
        LU = A.copy(order="F")
        perm, perm_inv, nt = oracle.lu(LU)
        L, U = unpack_lu(LU)
        assert approx_eq(L @ U, A[perm, :], 1e-13, 1e-13)


def test_lu_wide_and_complex_and_ties(oracle):
    rng = np.random.default_rng(8)
    # m < n: right block gets the unit-lower solve (factor.rs:278-285)
    A = np.asfortranarray(rng.standard_normal((40, 100)))
    LU = A.copy(order="F")
    perm, _, _ = oracle.lu(LU)
    L, U = unpack_lu(LU)
    assert np.allclose(L @ U, A[perm, :], atol=1e-12)
    # complex: abs1 = |re| + |im| pivoting (faer-traits lib.rs:2643-2646) == LAPACK izamax
    Z = np.asfortranarray(rng.standard_normal((120, 120)) + 1j * rng.standard_normal((120, 120)))
    LU = Z.copy(order="F")
    perm, _, _ = oracle.lu(LU)
    L, U = unpack_lu(LU)
    assert np.allclose(L @ U, Z[perm, :], atol=1e-12)
    _, piv = sla.lu_factor(Z)
    p = np.arange(120)
    for i, pi in enumerate(piv):
        p[i], p[pi] = p[pi], p[i]
    assert np.array_equal(p, perm)
    # ties resolve to the LOWEST row index (strict `>`, factor.rs:37-43)
    T = np.asfortranarray(np.array([[1.0, 2.0], [-1.0, 5.0], [1.0, 7.0]]))
    perm, _, nt = oracle.lu(T.copy(order="F"))
    assert perm[0] == 0
