"""`faer_b200_bidiag_in_place` / `faer_b200_tridiag_in_place` on host views that are not column-major (round-1 advisor item: they
aborted on the kernels' row-stride-1 assertion): row-major and strided inputs go through a compact column-major copy and give
the results of the column-major call (to a few ulps: the kernels see the same values, possibly under another leading dimension),
elements outside the view untouched."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def same(a, b, dtype, scale):
    fin = np.isfinite(b)
    return np.array_equal(np.isfinite(a), fin) and np.abs(a[fin] - b[fin]).max(initial=0.0) <= 256 * np.finfo(dtype).eps * scale


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bidiag_row_major_and_strided_host_views(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(410)
    for (m, n, bl, br) in [(8, 4, 4, 3), (33, 17, 8, 8), (130, 97, 32, 32)]:
        A = rng.standard_normal((m, n)).astype(dtype)
        ref = np.asfortranarray(A); Hl0 = np.zeros((bl, n), dtype=dtype, order="F"); Hr0 = np.zeros((br, n - 1), dtype=dtype, order="F")
        la.bidiag_in_place(ref, Hl0, Hr0)
        rm = np.array(A, order="C", copy=True); Hl = np.zeros_like(Hl0); Hr = np.zeros_like(Hr0)
        la.bidiag_in_place(rm, Hl, Hr)
        sc = max(m, n) * max(1.0, float(np.abs(A).max()))
        assert same(rm, ref, dtype, sc) and same(Hl, Hl0, dtype, sc) and same(Hr, Hr0, dtype, sc), (m, n)
        big = np.full((2 * m, 3 * n), np.nan, dtype=dtype); big[::2, ::3] = A
        Hl = np.zeros_like(Hl0); Hr = np.zeros_like(Hr0)
        la.bidiag_in_place(big[::2, ::3], Hl, Hr)
        assert same(big[::2, ::3], ref, dtype, sc) and same(Hl, Hl0, dtype, sc) and same(Hr, Hr0, dtype, sc), (m, n)
        assert np.all(np.isnan(big[1::2, :])) and np.all(np.isnan(big[:, 1::3])) and np.all(np.isnan(big[:, 2::3]))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tridiag_row_major_and_strided_host_views(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(411)
    for (n, b) in [(6, 2), (40, 8), (129, 16)]:
        G = rng.standard_normal((n, n))
        A = ((G + G.T) / 2).astype(dtype)
        A[np.triu_indices(n, 1)] = 123.0                       # the strict upper triangle is neither read nor written
        ref = np.asfortranarray(A); H0 = np.zeros((b, n - 1), dtype=dtype, order="F")
        la.tridiag_in_place(ref, H0)
        assert np.all(ref[np.triu_indices(n, 1)] == 123.0)
        rm = np.array(A, order="C", copy=True); H = np.zeros_like(H0)
        la.tridiag_in_place(rm, H)
        sc = n * max(1.0, float(np.abs(A).max()))
        assert same(rm, ref, dtype, sc) and same(H, H0, dtype, sc), n
        big = np.full((3 * n, 2 * n), np.nan, dtype=dtype); big[::3, ::2] = A
        H = np.zeros_like(H0)
        la.tridiag_in_place(big[::3, ::2], H)
        assert same(big[::3, ::2], ref, dtype, sc) and same(H, H0, dtype, sc), n
        assert np.all(np.isnan(big[1::3, :])) and np.all(np.isnan(big[2::3, :])) and np.all(np.isnan(big[:, 1::2]))
