"""BASELINE.json configs at their full sizes, through size-independent properties (the oracle would take many minutes
there): configs[2] LU n = 32768 on one GPU, configs[3] f32 QR 65536 x 4096, configs[4] bidiagonalization n = 8192 and c64
GEMM n = 8192, plus tridiagonalization n = 8192. configs[1] (LLT n = 16384) is tests/test_gpu_parity.py. Operands live
on the device; torch is the checker (probe products), never the path under test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
U = float(np.finfo(np.float64).eps)


def test_lu_n32768_property(fb, cuda_dev):
    """configs[2] on one GPU: P A x = L (U x) on probes; perm / perm_inv consistent; multipliers bounded by 1."""
    import torch
    la = fb.linalg
    n = 32768
    torch.manual_seed(11)
    A0 = torch.randn((n, n), dtype=torch.float64, device=cuda_dev).T
    A = A0.clone(memory_format=torch.preserve_format)
    p = torch.zeros(n, dtype=torch.int64, device=cuda_dev); pi = torch.zeros(n, dtype=torch.int64, device=cuda_dev)
    la.lu_in_place(A, p, pi)
    ar = torch.arange(n, device=cuda_dev)
    assert torch.equal(torch.sort(p).values, ar) and torch.equal(pi[p], ar)
    x = torch.randn((n, 4), dtype=torch.float64, device=cuda_dev)
    Ux = torch.triu(A) @ x
    umax = float(torch.triu(A).abs().max())
    Lx = torch.tril(A, -1) @ Ux + Ux
    assert float(torch.tril(A, -1).abs().max()) <= 1.0
    del A
    r = A0[p, :] @ x - Lx
    growth = max(1.0, umax / float(A0.abs().max()))
    assert float(r.abs().max()) <= 128 * 8 * U * n * float(A0.abs().max()) * float(x.abs().max()) * growth


def test_qr_f32_65536x4096_property(fb, cuda_dev):
    """configs[3]: Q^T A = [R; 0], and R keeps the column norms of A."""
    import torch
    la = fb.linalg
    m, n = 65536, 4096
    torch.manual_seed(12)
    A0 = torch.randn((n, m), dtype=torch.float32, device=cuda_dev).T  # column-major m x n
    A = A0.clone(memory_format=torch.preserve_format)
    bs = la.qr_recommended_block_size(m, n)
    H = torch.zeros((n, bs), dtype=torch.float32, device=cuda_dev).T
    assert la.qr_in_place(A, H).rank == n
    B = A0.clone(memory_format=torch.preserve_format)
    la.apply_block_householder_sequence_transpose_on_the_left_in_place(A, H, B)
    R = torch.triu(A[:n, :])
    u = float(np.finfo(np.float32).eps)
    scale = float(A0.abs().max()) * np.sqrt(8 * m)
    assert float((B[:n, :] - R).abs().max()) <= 128 * u * scale
    assert float(B[n:, :].abs().max()) <= 128 * u * scale
    cn_a = torch.linalg.vector_norm(A0.double(), dim=0); cn_r = torch.linalg.vector_norm(R.double(), dim=0)
    assert float(((cn_a - cn_r).abs() / cn_a).max()) <= 64 * u * np.sqrt(n)


def _probe_close(got, want, rel):
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= rel * scale, (float((got - want).abs().max()), scale)


def test_bidiag_n8192_probes(fb, cuda_dev):
    """configs[4]: U^T A V = B checked on probes: with y = V x, U^T (A y) must equal B x; and the Frobenius norm is kept."""
    import torch
    la = fb.linalg
    n, bl, br, k = 8192, 64, 64, 6
    torch.manual_seed(13)
    A0 = torch.randn((n, n), dtype=torch.float64, device=cuda_dev).T
    A = A0.clone(memory_format=torch.preserve_format)
    Hl = torch.zeros((n, bl), dtype=torch.float64, device=cuda_dev).T
    Hr = torch.zeros((n - 1, br), dtype=torch.float64, device=cuda_dev).T
    la.bidiag_in_place(A, Hl, Hr)
    d = torch.diagonal(A).clone(); e = torch.diagonal(A, 1).clone()
    fro2 = float((A0 * A0).sum())
    assert abs(float((d * d).sum() + (e * e).sum()) - fro2) <= 1e-11 * fro2
    x = torch.randn((k, n), dtype=torch.float64, device=cuda_dev).T                 # column-major n x k
    y = x.clone(memory_format=torch.preserve_format)
    Vr = A[:n - 1, 1:n].contiguous().t()                                              # column-major copy of the right basis
    la.apply_block_householder_sequence_on_the_left_in_place(Vr, Hr, y[1:, :])      # y = V x, V = 1 (+) Q_r
    z = (A0 @ y).t().contiguous().t()                                                 # column-major n x k
    la.apply_block_householder_sequence_transpose_on_the_left_in_place(A, Hl, z)    # U^T (A V x)
    Bx = d[:, None] * x
    Bx[:-1] += e[:, None] * x[1:]
    _probe_close(z, Bx, 1e-10)


def test_tridiag_n8192_probes(fb, cuda_dev):
    """Q^T A Q = T on probes (Q = 1 (+) Q'); the strict upper triangle is neither read nor written."""
    import torch
    la = fb.linalg
    n, b, k = 8192, 64, 6
    torch.manual_seed(14)
    G = torch.randn((n, n), dtype=torch.float64, device=cuda_dev)
    A0 = (G + G.T)
    del G
    A = A0.clone()
    iu = torch.triu_indices(n, n, 1, device=cuda_dev)
    A[iu[0], iu[1]] = float("nan")
    A = A.t().contiguous().t()                                                        # column-major, lower triangle valid
    H = torch.zeros((n - 1, b), dtype=torch.float64, device=cuda_dev).T
    la.tridiag_in_place(A, H)
    assert bool(torch.isnan(A[iu[0], iu[1]]).all())
    d = torch.diagonal(A).clone(); e = torch.diagonal(A, -1).clone()
    assert bool(torch.isfinite(d).all()) and bool(torch.isfinite(e).all())
    basis = torch.tril(A[1:, :n - 1]).t().contiguous().t()                            # reflectors, NaN-free, column-major
    x = torch.randn((k, n), dtype=torch.float64, device=cuda_dev).T
    y = x.clone(memory_format=torch.preserve_format)
    la.apply_block_householder_sequence_on_the_left_in_place(basis, H, y[1:, :])     # y = Q x
    z = (A0 @ y).t().contiguous().t()
    la.apply_block_householder_sequence_transpose_on_the_left_in_place(basis, H, z[1:, :])  # Q^T (A Q x)
    Tx = d[:, None] * x
    Tx[:-1] += e[:, None] * x[1:]
    Tx[1:] += e[:, None] * x[:-1]
    _probe_close(z, Tx, 1e-10)
    assert abs(float(d.sum()) - float(torch.diagonal(A0).sum())) <= 1e-10 * float(torch.diagonal(A0).abs().sum())  # trace


def test_c64_gemm_n8192_forward_bound(fb, cuda_dev):
    """configs[4], second item: c64 GEMM n = 8192 against a complex128 library product, within the forward bound of
    SURVEY appendix B on |A||B|."""
    import torch
    la = fb.linalg
    n = 8192
    torch.manual_seed(15)
    A = torch.randn((n, n), dtype=torch.complex128, device=cuda_dev).T
    B = torch.randn((n, n), dtype=torch.complex128, device=cuda_dev).T
    C = torch.full((n, n), float("nan"), dtype=torch.complex128, device=cuda_dev).T
    la.matmul(C, la.Accum.Replace, A, B, 1.0)
    ref = A @ B
    bound = 8 * n * U * (A.abs() @ B.abs())
    assert bool(((C - ref).abs() <= bound).all())
