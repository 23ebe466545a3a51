"""Oracle restatement of LDLT (cholesky/ldlt/factor.rs, solve.rs) against the reference's own tests:
test_simd_cholesky (factor.rs:771-817: n = 0..64, c64, L D L^H ~ A to 1e-12), test_cholesky (818-866: n in
{2, 4, 8, 31, 127, 240}, recursion threshold = block size = 32), and the solve test (solve.rs: A X ~ B); plus the
documented semantics: D on the diagonal / unit-lower L below, upper triangle untouched, ZeroPivot { index } with the
diagonal initialised up to and including the failing column (factor.rs:757-765), dynamic regularisation with and without
expected signs (122-144), and the relation to LLT on SPD input."""
import numpy as np
import pytest


def _hpd(rng, n, dtype=np.complex128):
    G = rng.standard_normal((n, n))
    if np.issubdtype(dtype, np.complexfloating):
        G = G + 1j * rng.standard_normal((n, n))
    G = G.astype(dtype)
    return np.asfortranarray(G @ G.conj().T)


def _factors(LD):
    n = LD.shape[0]
    return np.tril(LD, -1) + np.eye(n, dtype=LD.dtype), np.real(np.diagonal(LD)).copy()


def _close(a, b, tol=1e-12):
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))) * max(1, a.shape[0]))


def test_simd_cholesky_reference_case(oracle):
    rng = np.random.default_rng(0)
    for n in range(0, 65):
        A = _hpd(rng, n)
        LD = A.copy(order="F")
        assert oracle.ldlt(LD) == (-1, 0)
        L, D = _factors(LD)
        assert _close(L @ np.diag(D) @ L.conj().T, A), n
        assert np.all(D > 0)


def test_cholesky_recursion_reference_case(oracle):
    rng = np.random.default_rng(0)
    for n in [2, 4, 8, 31, 127, 240]:
        A = _hpd(rng, n)
        LD = A.copy(order="F")
        assert oracle.ldlt(LD, recursion_threshold=32, block_size=32) == (-1, 0)
        L, D = _factors(LD)
        assert _close(L @ np.diag(D) @ L.conj().T, A), n


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128, np.complex64])
def test_ldlt_semantics_all_dtypes(oracle, dtype):
    rng = np.random.default_rng(3)
    u = np.finfo(dtype).eps
    for n in [1, 7, 64, 65, 200, 333]:
        # indefinite self-adjoint input with a dominant diagonal of mixed signs (no pivoting in LDLT)
        G = rng.standard_normal((n, n))
        if np.issubdtype(dtype, np.complexfloating):
            G = G + 1j * rng.standard_normal((n, n))
        s = np.where(rng.random(n) < 0.4, -1.0, 1.0)
        A = np.asfortranarray(((G + G.conj().T) / np.sqrt(n) + np.diag(s * 4.0)).astype(dtype))
        LD = A.copy(order="F")
        LD[np.triu_indices(n, 1)] = np.nan  # the strict upper triangle is neither read nor written
        assert oracle.ldlt(LD) == (-1, 0)
        assert np.all(np.isnan(LD[np.triu_indices(n, 1)]))
        L, D = _factors(np.tril(LD))
        assert np.all(np.imag(np.diagonal(np.tril(LD))) == 0)
        ev = np.linalg.eigvalsh(A.astype(np.complex128))
        assert (D < 0).sum() == (ev < 0).sum() and np.all(D != 0)  # Sylvester's law of inertia
        R = L.astype(np.complex128) @ np.diag(D.astype(np.float64)) @ L.conj().T.astype(np.complex128)
        assert np.abs(R - A).max() <= 64 * n * u * np.abs(A).max(), (n, np.abs(R - A).max())
        # solve (ldlt/solve.rs:11-49)
        B = rng.standard_normal((n, 3)).astype(dtype)
        X = np.asfortranarray(B.copy())
        oracle.ldlt_solve(np.asfortranarray(np.tril(LD)), X)
        assert np.abs(A.astype(np.complex128) @ X - B).max() <= 256 * n * u * max(1.0, np.abs(X).max()) * np.abs(A).max()


def test_ldlt_matches_llt_on_spd(oracle):
    rng = np.random.default_rng(4)
    for n in [5, 64, 150]:
        A = _hpd(rng, n, np.float64)
        LL = A.copy(order="F"); assert oracle.llt(LL)[0] == -1
        LD = A.copy(order="F"); assert oracle.ldlt(LD)[0] == -1
        L, D = _factors(LD)
        Lc = np.tril(LL)
        assert np.allclose(D, np.diagonal(Lc) ** 2, rtol=1e-10)
        assert np.allclose(L, Lc / np.diagonal(Lc)[None, :], rtol=1e-10, atol=1e-12)


def test_zero_pivot_and_diagonal_initialisation(oracle):
    rng = np.random.default_rng(5)
    n = 40
    A = _hpd(rng, n, np.float64)
    # make the Schur complement vanish exactly at column 3: rows/cols 3 duplicates of a combination is hard to do exactly,
    # so use a matrix whose leading 3x3 block is diagonal and a_33 chosen from the exact recurrence
    A[:4, :4] = np.diag([2.0, 4.0, 8.0, 0.0]); A[3, :3] = A[:3, 3] = [2.0, 4.0, 8.0]
    A[3, 3] = 2.0 + 4.0 + 8.0   # d_3 = a_33 - sum l_3k^2 d_k = 14 - (1*2 + 1*4 + 1*8) = 0 exactly
    A = np.asfortranarray(A)
    LD = A.copy(order="F")
    fail, count = oracle.ldlt(LD)
    assert (fail, count) == (3, 0)
    assert np.array_equal(np.diagonal(LD)[:4], [2.0, 4.0, 8.0, 0.0])   # initialised up to and including the failing column
    assert np.array_equal(np.diagonal(LD)[4:], np.diagonal(A)[4:])     # untouched beyond it


def test_dynamic_regularisation(oracle):
    """factor.rs:122-144: with signs, a pivot of the wrong sign or below eps in magnitude is replaced by sign * delta (only
    the +1 case is counted); without signs only |d| <= eps is replaced, keeping the sign of d."""
    A = np.asfortranarray(np.diag([1.0, -2.0, 1e-20, -1e-20, 3.0]))
    LD = A.copy(order="F")
    fail, count = oracle.ldlt(LD, delta=1e-3, eps=1e-10)
    assert fail == -1 and count == 0
    assert np.array_equal(np.diagonal(LD), [1.0, -2.0, 1e-3, -1e-3, 3.0])
    LD = A.copy(order="F")
    fail, count = oracle.ldlt(LD, delta=1e-3, eps=1e-10, signs=[1, 1, 1, -1, -1])
    assert fail == -1 and count == 2                       # columns 1 (wrong sign) and 2 (tiny) with sign +1
    assert np.array_equal(np.diagonal(LD), [1.0, 1e-3, 1e-3, -1e-3, -1e-3])
    # regularisation off unless both delta and eps are positive (factor.rs:744-745)
    LD = A.copy(order="F")
    assert oracle.ldlt(LD, delta=1e-3, eps=0.0, signs=[1, 1, 1, -1, -1]) == (-1, 0)
    assert np.array_equal(np.diagonal(LD), np.diagonal(A))
