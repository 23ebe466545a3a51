"""Eigenvalues of a self-adjoint matrix through the C ABI (`self_adjoint_evd` with U = None) against LAPACK:
tridiagonalization on the GPU + one bisection thread per value (csrc/evd.cu; the bisection routine is checked on the CPU by
tests/test_tridiag_ev_cpu.py, tridiag.cu by tests/test_gpu_condensed.py). Tolerance 32 n u |lambda|_max (the reference's
EVD tests use eps * n on unit-scale matrices)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_self_adjoint_eigenvalues_vs_lapack(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(131)
    u = np.finfo(dtype).eps
    for n in [1, 2, 3, 16, 64, 100, 257, 700]:
        G = rng.standard_normal((n, n)).astype(dtype)
        A = (G + G.T).astype(dtype)
        ref = np.linalg.eigvalsh(A.astype(np.float64))
        for order in "FC":
            M = np.array(A, order=order)
            M[np.triu_indices(n, 1)] = np.nan                      # only the lower triangle may be read
            keep = M.copy()
            s = la.self_adjoint_eigenvalues(M)
            assert np.array_equal(np.nan_to_num(M), np.nan_to_num(keep))
            assert s.shape == (n,) and np.all(np.diff(s) >= 0)
            assert np.abs(s - ref).max() <= 32 * n * u * max(1.0, np.abs(ref).max()), (n, order)
        up = np.array(A, order="F"); up[np.tril_indices(n, -1)] = np.nan
        s = fb.solvers.self_adjoint_eigenvalues(up, fb.solvers.Side.Upper)
        assert np.abs(s - ref).max() <= 32 * n * u * max(1.0, np.abs(ref).max()), n


def test_self_adjoint_eigenvalues_n8192_device(fb, cuda_dev):
    import torch
    la = fb.linalg
    n = 8192
    torch.manual_seed(132)
    G = torch.randn((n, n), dtype=torch.float64, device=cuda_dev)
    A = G + G.T
    del G
    s = la.self_adjoint_eigenvalues(A)
    assert s.is_cuda and tuple(s.shape) == (n,) and bool((s[1:] >= s[:-1]).all())
    tr = float(torch.diagonal(A).sum()); fro2 = float((A * A).sum())
    assert abs(float(s.sum()) - tr) <= 1e-10 * float(torch.diagonal(A).abs().sum())
    assert abs(float((s * s).sum()) - fro2) <= 1e-11 * fro2
    ref = torch.linalg.eigvalsh(A)
    assert float((s - ref).abs().max()) <= 32 * n * np.finfo(np.float64).eps * float(ref.abs().max())
