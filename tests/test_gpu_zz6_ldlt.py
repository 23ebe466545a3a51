"""GPU tests of LDLT through the C ABI against the oracle's restatement, which
tests/test_oracle_ldlt_cpu.py pins to the reference's own tests (kernel and driver: csrc/ldlt_f64.cu).

Contract: ZeroPivot index and regularisation count exact; D and L within 64 n u |A| of the reconstruction and close to
the oracle's (same recurrence, different blocking); leaf blocks (n <= 64) bit-identical to the oracle's leaf; the strict
upper triangle untouched; the solve within the backward bound."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
U = np.finfo(np.float64).eps


def _indefinite(rng, n):
    G = rng.standard_normal((n, n))
    s = np.where(rng.random(n) < 0.4, -1.0, 1.0)
    return np.asfortranarray((G + G.T) / np.sqrt(max(n, 1)) + np.diag(4.0 * s))


def test_ldlt_vs_oracle(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(71)
    for n in [1, 2, 7, 33, 64, 65, 128, 129, 200, 257, 600, 1000]:
        A = _indefinite(rng, n)
        want = A.copy(order="F"); assert oracle.ldlt(want) == (-1, 0)
        got = A.copy(order="F"); got[np.triu_indices(n, 1)] = np.nan
        info = la.ldlt_in_place(got)
        assert info.dynamic_regularization_count == 0
        assert np.all(np.isnan(got[np.triu_indices(n, 1)])), n       # the strict upper triangle is untouched
        LD = np.tril(got)
        L = np.tril(LD, -1) + np.eye(n); D = np.diagonal(LD).copy()
        assert np.abs(L @ np.diag(D) @ L.T - A).max() <= 64 * n * U * np.abs(A).max(), n
        if n <= 64:
            assert np.array_equal(LD, np.tril(want)), n                # same recurrence, same fma order as the leaf
        else:
            assert np.allclose(LD, np.tril(want), rtol=1e-9, atol=1e-11), n
        # solve
        B = np.asfortranarray(rng.standard_normal((n, 5)))
        X = B.copy(order="F"); la.ldlt_solve_in_place(np.asfortranarray(LD), X)
        Xo = B.copy(order="F"); oracle.ldlt_solve(np.asfortranarray(np.tril(want)), Xo)
        assert np.abs(A @ X - B).max() <= 256 * n * U * max(1.0, np.abs(X).max()) * np.abs(A).max(), n
        assert np.allclose(X, Xo, rtol=1e-8, atol=1e-10), n


def test_ldlt_zero_pivot_and_regularisation(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(72)
    n = 200
    A = _indefinite(rng, n)
    A[:4, :4] = np.diag([2.0, 4.0, 8.0, 0.0]); A[3, :3] = A[:3, 3] = [2.0, 4.0, 8.0]; A[3, 3] = 14.0
    want = A.copy(order="F"); fail, _ = oracle.ldlt(want); assert fail == 3
    got = A.copy(order="F")
    with pytest.raises(la.LdltError) as e:
        la.ldlt_in_place(got)
    assert e.value.index == 3
    assert np.array_equal(np.diagonal(got)[:4], [2.0, 4.0, 8.0, 0.0])
    # dynamic regularisation, with and without expected signs (ldlt/factor.rs:122-144)
    Dg = np.asfortranarray(np.diag([1.0, -2.0, 1e-20, -1e-20, 3.0]))
    got = Dg.copy(order="F"); info = la.ldlt_in_place(got, regularization=(1e-3, 1e-10))
    assert info.dynamic_regularization_count == 0 and np.array_equal(np.diagonal(got), [1.0, -2.0, 1e-3, -1e-3, 3.0])
    got = Dg.copy(order="F"); info = la.ldlt_in_place(got, regularization=(1e-3, 1e-10), signs=[1, 1, 1, -1, -1])
    assert info.dynamic_regularization_count == 2 and np.array_equal(np.diagonal(got), [1.0, 1e-3, 1e-3, -1e-3, -1e-3])
    # a larger case through the blocked driver: two isolated pivots go through the regulariser (a tiny one, a wrong-signed
    # one), count and factors agree with the oracle
    n = 300
    A = _indefinite(rng, n)
    for j, v in ((50, 1e-14), (180, -3.0)):
        A[j, :] = 0.0; A[:, j] = 0.0; A[j, j] = v
    sg = np.where(np.diagonal(A) > 0, 1, -1).astype(np.int8); sg[180] = 1
    want = A.copy(order="F"); fo, co = oracle.ldlt(want, delta=1e-2, eps=1e-9, signs=sg)
    assert (fo, co) == (-1, 2)
    got = A.copy(order="F"); info = la.ldlt_in_place(got, regularization=(1e-2, 1e-9), signs=sg)
    assert info.dynamic_regularization_count == co
    assert np.diagonal(got)[50] == 1e-2 and np.diagonal(got)[180] == 1e-2
    assert np.allclose(np.tril(got), np.tril(want), rtol=1e-9, atol=1e-11)


def test_ldlt_device_resident_large(fb, cuda_dev):
    """n = 8192 on device memory: L D L^T x = A x on probes; the upper triangle is untouched."""
    import torch
    la = fb.linalg
    n = 8192
    torch.manual_seed(73)
    G = torch.randn((n, n), dtype=torch.float64, device=cuda_dev)
    s = torch.where(torch.rand(n, device=cuda_dev, dtype=torch.float64) < 0.4, -1.0, 1.0)
    A0 = ((G + G.T) / np.sqrt(n) + torch.diag(4.0 * s)).T  # symmetric, column-major view
    del G
    A = A0.clone(memory_format=torch.preserve_format)
    info = la.ldlt_in_place(A)
    assert info.dynamic_regularization_count == 0
    assert torch.equal(torch.triu(A, 1), torch.triu(A0, 1))
    L = torch.tril(A, -1) + torch.eye(n, dtype=torch.float64, device=cuda_dev)
    D = torch.diagonal(A).clone()
    assert int((D < 0).sum()) == int((s < 0).sum())                 # inertia (the diagonal dominates)
    x = torch.randn((n, 6), dtype=torch.float64, device=cuda_dev)
    r = A0 @ x - L @ (D[:, None] * (L.T @ x))
    assert float(r.abs().max()) <= 128 * U * n * float(A0.abs().max()) * float(x.abs().max())


def test_ldlt_solver_class(fb, cuda_dev):
    """faer_b200.solvers.Ldlt through the C ABI: the shared cases of tests/solvers_cases.py (also run on the CPU with the
    oracle-backed stand-in)."""
    from solvers_cases import run_ldlt
    run_ldlt(fb.solvers)
