"""`inverse_[unit_]triangular_{lower,upper}_in_place` through the C ABI for f64 / f32 / c64 / c32 (csrc/reconstruct_types.cu: the
triangular solve applied to the identity, the triangle copied out): linalg/triangular_inverse.rs — the product with the source
triangle is the identity, and nothing outside the triangle (nor the diagonal, for the unit variants) is written."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32, np.complex128, np.complex64]


def rdt(dtype):
    return np.float32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64


def wide(x):
    return x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)


@pytest.mark.parametrize("dtype", DTYPES)
def test_invert_triangular(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(481)
    u = float(np.finfo(rdt(dtype)).eps)
    for n in [1, 2, 33, 100, 257]:
        G = rng.standard_normal((n, n))
        if np.issubdtype(dtype, np.complexfloating):
            G = G + 1j * rng.standard_normal((n, n))
        G = G / np.sqrt(n) + 2.0 * np.eye(n)                     # well-conditioned triangles
        src = np.asfortranarray(G.astype(dtype))
        for lower in (True, False):
            for unit in (False, True):
                tri = np.tril if lower else np.triu
                T = tri(wide(src), -1 if lower else 1) + (np.eye(n) if unit else np.diag(np.diag(wide(src))))
                dst = np.full((n, n), np.nan, dtype=dtype, order="F" if n % 2 else "C")   # both layouts
                sel = tri(np.ones((n, n), bool))
                if unit:
                    np.fill_diagonal(sel, False)
                dst[sel] = 7
                la.invert_triangular(dst, src, lower, unit)
                assert np.all(np.isnan(dst[~sel])), (n, lower, unit)                      # nothing else is written
                inv = np.where(sel, wide(dst), 0) + (np.eye(n) if unit else 0)
                assert np.abs(inv @ T - np.eye(n)).max() <= 64 * n * u * np.linalg.cond(T), (n, lower, unit)
