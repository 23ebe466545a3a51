"""f32 LLT through the C ABI (csrc/llt.cu: the templated leaf kernel and recursive driver instantiated for float) against the
oracle in f32.
Contract as for f64 (tests/test_gpu_parity.py): error index and regularisation count exact, leaf blocks (n <= 64)
bit-identical to the oracle's leaf, L L^T = A within 64 n u |A|, strict upper triangle untouched, solve within the
backward bound."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
U = float(np.finfo(np.float32).eps)


def _spd(rng, n):
    G = rng.standard_normal((n, n)).astype(np.float32)
    return np.asfortranarray((G @ G.T + n * np.eye(n, dtype=np.float32)).astype(np.float32))


def test_llt_f32_vs_oracle(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(91)
    for n in [1, 2, 17, 64, 65, 128, 129, 300, 700, 1500]:
        A = _spd(rng, n)
        want = A.copy(order="F"); assert oracle.llt(want) == (-1, 0)
        got = A.copy(order="F"); got[np.triu_indices(n, 1)] = np.nan
        info = la.cholesky_in_place(got)
        assert info.dynamic_regularization_count == 0
        assert np.all(np.isnan(got[np.triu_indices(n, 1)])), n
        L = np.tril(got).astype(np.float64)
        assert np.abs(L @ L.T - A).max() <= 64 * n * U * np.abs(A).max(), n
        if n <= 64:
            assert np.array_equal(np.tril(got), np.tril(want)), n
        else:
            assert np.allclose(np.tril(got), np.tril(want), rtol=2e-3, atol=2e-3 * np.abs(want).max()), n
        B = np.asfortranarray(rng.standard_normal((n, 4)).astype(np.float32))
        X = B.copy(order="F"); la.llt_solve_in_place(np.asfortranarray(np.tril(got)), X)
        r = A.astype(np.float64) @ X - B
        assert np.abs(r).max() <= 256 * n * U * np.abs(A).max() * max(1.0, np.abs(X).max()), n


def test_llt_f32_error_index_and_regularisation(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(92)
    for n, bad in [(50, 7), (300, 131), (300, 299)]:
        A = _spd(rng, n); A[bad, bad] = -1.0
        want = A.copy(order="F"); fo, _ = oracle.llt(want)
        assert fo == bad
        with pytest.raises(la.LltError) as e:
            la.cholesky_in_place(A.copy(order="F"))
        assert e.value.index == bad
    # dynamic regularisation: count equals the oracle's (llt/factor.rs:85-87, ldlt/factor.rs:122-144)
    A = np.asfortranarray(np.diag(np.array([4.0, 1e-30, 9.0, -2.0, 1.0], dtype=np.float32)))
    want = A.copy(order="F"); ro = oracle.llt(want, delta=1e-2, eps=1e-6)
    got = A.copy(order="F"); info = la.cholesky_in_place(got, regularization=(1e-2, 1e-6))
    assert ro == (-1, 2) and info.dynamic_regularization_count == 2
    assert np.array_equal(np.diagonal(got), np.diagonal(want))
