"""LDLT through the C ABI beyond the f64 factorization (csrc/ldlt_types.cu): factor / solve for f32 / c64 / c32 (the flat-map launch
sequence of ldlt_core.cuh, run thread by thread on the CPU in test_ldlt_types_emul_cpu.py) and `ldlt_reconstruct` / `ldlt_inverse`
for every dtype, against the oracle's restatement of cholesky/ldlt/factor.rs and the reference's own tests (ldlt/solve.rs,
reconstruct.rs, inverse.rs: n = 50, complex draws).

Contract: ZeroPivot index and regularisation count exact; L and D reconstruct A within 64 n u |A| and are close to the oracle's;
the strict upper triangle untouched; the solve within the backward bound for both conjugation settings; reconstruct / inverse write
the lower triangle only."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS_DTYPES = [np.float32, np.complex128, np.complex64]          # factor / solve added here (f64: test_gpu_zz6_ldlt.py)
ALL_DTYPES = [np.float64, np.float32, np.complex128, np.complex64]


def rdt(dtype):
    return np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64


def wide(x):
    return x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)


def indefinite(rng, n, dtype):
    """self-adjoint, indefinite, safely factorable without pivoting (the diagonal dominates, mixed signs)"""
    G = rng.standard_normal((n, n))
    if np.issubdtype(dtype, np.complexfloating):
        G = G + 1j * rng.standard_normal((n, n))
    s = np.where(rng.random(n) < 0.4, -1.0, 1.0)
    A = (G + G.conj().T) / np.sqrt(max(n, 1)) + np.diag(4.0 * s)
    A[np.diag_indices(n)] = A[np.diag_indices(n)].real
    return np.asfortranarray(A.astype(dtype))


def crand(rng, shape, dtype):
    a = rng.standard_normal(shape)
    if np.issubdtype(dtype, np.complexfloating):
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a.astype(dtype))


@pytest.mark.parametrize("dtype", FS_DTYPES)
def test_ldlt_types_vs_oracle(fb, oracle, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(471)
    u = float(np.finfo(rdt(dtype)).eps)
    for n in [1, 2, 7, 33, 64, 65, 129, 300]:
        A = indefinite(rng, n, dtype)
        want = A.copy(order="F"); assert oracle.ldlt(want) == (-1, 0)
        got = A.copy(order="F"); got[np.triu_indices(n, 1)] = np.nan
        info = la.ldlt_in_place(got)
        assert info.dynamic_regularization_count == 0
        assert np.all(np.isnan(got[np.triu_indices(n, 1)])), n       # the strict upper triangle is untouched
        LD = np.tril(got)
        assert np.all(np.diagonal(LD).imag == 0)
        L = wide(np.tril(LD, -1)) + np.eye(n); D = wide(np.diagonal(LD)).real
        assert np.abs(L @ np.diag(D) @ L.conj().T - wide(A)).max() <= 64 * n * u * np.abs(A).max(), n
        assert np.allclose(LD, np.tril(want), rtol=2e3 * u, atol=2e3 * u), n
        # solve, both conjugation settings (ldlt/solve.rs tests)
        B = crand(rng, (n, 5), dtype)
        for conj in (0, 1):
            X = B.copy(order="F"); la.ldlt_solve_in_place(np.asfortranarray(LD), X, conj)
            Ae = wide(A).conj() if conj else wide(A)
            assert np.abs(Ae @ wide(X) - wide(B)).max() <= 256 * n * u * max(1.0, np.abs(X).max()) * np.abs(A).max(), (n, conj)
        # D given as a separate vector of the matrix dtype
        X2 = B.copy(order="F"); la.ldlt_solve_in_place(np.asfortranarray(LD), X2, 0, D=np.ascontiguousarray(np.diagonal(LD)))
        X1 = B.copy(order="F"); la.ldlt_solve_in_place(np.asfortranarray(LD), X1, 0)
        assert np.array_equal(X1, X2), n


@pytest.mark.parametrize("dtype", FS_DTYPES)
def test_ldlt_types_zero_pivot_and_regularisation(fb, oracle, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(472)
    n = 120
    A = indefinite(rng, n, dtype)
    A[:4, :4] = np.diag([2.0, 4.0, 8.0, 0.0]); A[3, :3] = A[:3, 3] = [2.0, 4.0, 8.0]; A[3, 3] = 14.0
    want = A.copy(order="F"); fail, _ = oracle.ldlt(want); assert fail == 3
    got = A.copy(order="F")
    with pytest.raises(la.LdltError) as e:
        la.ldlt_in_place(got)
    assert e.value.index == 3
    assert np.array_equal(np.diagonal(got)[:4], np.array([2.0, 4.0, 8.0, 0.0], dtype=dtype))
    # dynamic regularisation, with and without expected signs (ldlt/factor.rs:122-144)
    r = rdt(dtype)
    Dg = np.asfortranarray(np.diag(np.array([1.0, -2.0, 1e-20, -1e-20, 3.0])).astype(dtype))
    got = Dg.copy(order="F"); info = la.ldlt_in_place(got, regularization=(1e-3, 1e-10))
    assert info.dynamic_regularization_count == 0
    assert np.array_equal(np.diagonal(got).real, np.array([1.0, -2.0, r(1e-3), -r(1e-3), 3.0], dtype=r))
    got = Dg.copy(order="F"); info = la.ldlt_in_place(got, regularization=(1e-3, 1e-10), signs=[1, 1, 1, -1, -1])
    assert info.dynamic_regularization_count == 2
    assert np.array_equal(np.diagonal(got).real, np.array([1.0, r(1e-3), r(1e-3), -r(1e-3), -r(1e-3)], dtype=r))
    # two isolated pivots go through the regulariser inside a larger matrix: count and factors agree with the oracle
    n = 150
    A = indefinite(rng, n, dtype)
    for j, v in ((50, 1e-14), (101, -3.0)):
        A[j, :] = 0.0; A[:, j] = 0.0; A[j, j] = v
    sg = np.where(np.diagonal(A).real > 0, 1, -1).astype(np.int8); sg[101] = 1
    want = A.copy(order="F"); fo, co = oracle.ldlt(want, delta=1e-2, eps=1e-9, signs=sg)
    assert (fo, co) == (-1, 2)
    got = A.copy(order="F"); info = la.ldlt_in_place(got, regularization=(1e-2, 1e-9), signs=sg)
    assert info.dynamic_regularization_count == co
    assert np.diagonal(got)[50] == r(1e-2) and np.diagonal(got)[101] == r(1e-2)
    u = float(np.finfo(r).eps)
    assert np.allclose(np.tril(got), np.tril(want), rtol=2e3 * u, atol=2e3 * u)


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_ldlt_reconstruct_and_inverse(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(473)
    u = float(np.finfo(rdt(dtype)).eps)
    for n in [1, 50, 130, 257]:
        A = indefinite(rng, n, dtype)
        LD = A.copy(order="F"); la.ldlt_in_place(LD)
        fill = dtype(7.0)
        out = np.full((n, n), np.nan, dtype=dtype, order="F"); out[np.tril_indices(n)] = fill
        la.ldlt_reconstruct(out, LD)          # LD's strict upper part still holds A's entries: it must not be read
        assert np.all(np.isnan(out[np.triu_indices(n, 1)])), n               # only the lower triangle is written
        assert np.abs(np.tril(wide(out)) - np.tril(wide(A))).max() <= 128 * n * u * np.abs(A).max(), n
        inv = np.full((n, n), np.nan, dtype=dtype, order="F"); inv[np.tril_indices(n)] = fill
        la.ldlt_inverse(inv, LD)
        assert np.all(np.isnan(inv[np.triu_indices(n, 1)])), n
        lo = np.tril(wide(inv)); full = lo + np.tril(lo, -1).conj().T
        assert np.abs(full @ wide(A) - np.eye(n)).max() <= 128 * n * u * np.linalg.cond(wide(A)), n
        # L and D given separately (`Ldlt::L()`, `Ldlt::D()`)
        out2 = np.full((n, n), np.nan, dtype=dtype, order="F"); out2[np.tril_indices(n)] = fill
        Lsep = np.asfortranarray(np.tril(LD, -1) + np.eye(n, dtype=dtype))
        la.ldlt_reconstruct(out2, Lsep, D=np.ascontiguousarray(np.diagonal(LD)))
        assert np.array_equal(np.tril(out2), np.tril(out)), n


@pytest.mark.parametrize("dtype", [np.complex128, np.float32])
def test_ldlt_solver_class_other_dtypes(fb, cuda_dev, dtype):
    """`Ldlt::new` and the Solve family (solvers.rs:818-872, 93-282) on complex / f32 matrices: L / D contracts, both sides, the four
    solves (A x, conj(A) x, A^T x, A^H x)."""
    sv = fb.solvers
    rng = np.random.default_rng(474)
    n = 50
    u = float(np.finfo(rdt(dtype)).eps)
    A = indefinite(rng, n, dtype)
    dec = sv.Ldlt.new(A, sv.Side.Lower)
    L, D = wide(np.asarray(dec.L())), wide(np.asarray(dec.D()))
    assert np.all(np.diag(L) == 1) and np.all(np.triu(L, 1) == 0) and D.shape == (n,) and np.all(D.imag == 0)
    Aw = wide(A)
    assert np.abs(L @ np.diag(D) @ L.conj().T - Aw).max() <= 64 * n * u * np.abs(A).max()
    assert (D.real < 0).sum() == (np.linalg.eigvalsh(Aw) < 0).sum()
    assert np.abs(wide(np.asarray(dec.reconstruct())) - Aw).max() <= 128 * n * u * np.abs(A).max()
    B = crand(rng, (n, 4), dtype)
    tol = 256 * n * u * np.linalg.cond(Aw) * np.abs(B).max()
    assert np.abs(Aw @ wide(np.asarray(dec.solve(B))) - wide(B)).max() <= tol
    assert np.abs(Aw.conj() @ wide(np.asarray(dec.solve_conjugate(B))) - wide(B)).max() <= tol
    assert np.abs(Aw.T @ wide(np.asarray(dec.solve_transpose(B))) - wide(B)).max() <= tol
    assert np.abs(Aw.conj().T @ wide(np.asarray(dec.solve_adjoint(B))) - wide(B)).max() <= tol
    poisoned = A.copy(order="F"); poisoned[np.tril_indices(n, -1)] = np.nan
    up = sv.Ldlt.new(poisoned, sv.Side.Upper)
    assert np.allclose(np.asarray(up.L()), np.asarray(dec.L()), rtol=1e3 * u, atol=1e3 * u)
