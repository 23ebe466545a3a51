"""Singular values through the C ABI (`svd` with U = V = None; BASELINE.json configs[4]) against LAPACK: bidiagonalization
on the GPU + one bisection thread per value (csrc/svd.cu; the bisection routine itself is checked on the CPU by
tests/test_bidiag_sv_cpu.py, bidiag.cu by tests/test_gpu_condensed.py). Tolerance: the reference's own SVD tests use
eps * n on unit-scale matrices (svd/mod.rs tests); here 32 max(m, n) u sigma_max."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_singular_values_vs_lapack(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(111)
    u = np.finfo(dtype).eps
    for (m, n) in [(1, 1), (2, 2), (5, 3), (3, 5), (64, 64), (100, 37), (37, 100), (300, 300), (1000, 130), (700, 701)]:
        for order in "FC":
            A = np.array(rng.standard_normal((m, n)), dtype=dtype, order=order)
            keep = A.copy()
            s = la.singular_values(A)
            assert np.array_equal(A, keep)                              # the input is not modified
            ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
            assert s.shape == (min(m, n),) and np.all(np.diff(s) <= 0) and np.all(s >= 0)
            assert np.abs(s - ref).max() <= 32 * max(m, n) * u * ref.max(), (m, n, order)
    # rank-deficient input
    B = rng.standard_normal((200, 5)).astype(dtype) @ rng.standard_normal((5, 120)).astype(dtype)
    s = la.singular_values(np.asfortranarray(B)); ref = np.linalg.svd(B.astype(np.float64), compute_uv=False)
    assert np.abs(s - ref).max() <= 64 * 200 * u * ref.max()
    assert fb.solvers.singular_values(np.asfortranarray(B)).shape == (120,)
    # tall and rank-deficient: the QR-first path (nrows / ncols > 11/6) reports the deficiency and the driver falls back
    B = rng.standard_normal((300, 5)).astype(dtype) @ rng.standard_normal((5, 100)).astype(dtype)
    s = la.singular_values(np.asfortranarray(B)); ref = np.linalg.svd(B.astype(np.float64), compute_uv=False)
    assert np.abs(s - ref).max() <= 64 * 300 * u * ref.max()


def test_singular_values_n8192_device(fb, cuda_dev):
    """configs[4] size on device memory; checked against the invariants sum sigma^2 = |A|_F^2 and sum sigma^4 = |A^T A|_F^2
    and against a library SVD of the same matrix."""
    import torch
    la = fb.linalg
    n = 8192
    torch.manual_seed(112)
    A = torch.randn((n, n), dtype=torch.float64, device=cuda_dev).T
    s = la.singular_values(A)
    assert s.is_cuda and tuple(s.shape) == (n,)
    assert bool((s[:-1] >= s[1:]).all()) and float(s.min()) >= 0.0
    p2 = float((A * A).sum()); G = A.T @ A; p4 = float((G * G).sum())
    assert abs(float((s ** 2).sum()) - p2) <= 1e-11 * p2
    assert abs(float((s ** 4).sum()) - p4) <= 1e-10 * p4
    ref = torch.linalg.svdvals(A)
    assert float((s - ref).abs().max()) <= 32 * n * np.finfo(np.float64).eps * float(ref.max())
