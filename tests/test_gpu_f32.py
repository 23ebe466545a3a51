"""GPU parity tests for the f32 path (3xTF32 error-compensated tensor-core GEMM) against the oracle (f32 schoolbook).

Tolerance: the reference's own matmul test uses abs 1e-3 on 32-bit inputs (matmul/mod.rs:2015-2016); we use the much
tighter forward bound  |dC| <= 4 k u32 sum|a||b|  (u32 = 2^-24) — i.e. the compensated product must be as accurate as an
fp32 FMA chain, NOT merely tf32-accurate (a plain tf32 product would violate this bound by ~2^10).
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
U32 = 2.0 ** -24
SHAPES = [(2, 2, 2), (8, 8, 8), (16, 16, 16), (127, 127, 127), (128, 128, 128), (129, 129, 129), (15, 15, 15), (17, 17, 17),
          (1, 1, 1), (1, 16, 16), (16, 1, 16), (16, 16, 1), (0, 4, 4), (4, 4, 0), (63, 9, 100), (100, 63, 9), (300, 520, 260)]


def test_f32_matmul_vs_oracle(fb, oracle, cuda_dev):
    import torch
    la = fb.linalg
    rng = np.random.default_rng(41)
    for (m, n, k) in SHAPES:
        for layout in itertools.product("CF", repeat=3):
            for add, alpha in [(False, 1.0), (True, -1.0), (True, 0.5)]:
                A = np.array(rng.standard_normal((m, k)), dtype=np.float32, order=layout[0])
                B = np.array(rng.standard_normal((k, n)), dtype=np.float32, order=layout[1])
                C0 = np.array(rng.standard_normal((m, n)), dtype=np.float32, order=layout[2])
                want = C0.copy(order="K")
                if not add:
                    want[...] = np.nan
                oracle.matmul(want, add, A, B, alpha)
                exact = (alpha * (A.astype(np.float64) @ B.astype(np.float64)) + (C0 if add else 0)).astype(np.float64)
                bound = 4 * max(k, 1) * U32 * abs(alpha) * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)) \
                    + 4 * U32 * np.abs(exact) + (4 * U32 * np.abs(C0).astype(np.float64) if add else 0.0) + 1e-30
                got = C0.copy(order="K")
                if not add:
                    got[...] = np.nan
                la.matmul(got, la.Accum.Add if add else la.Accum.Replace, A, B, alpha)
                assert np.all(np.abs(got.astype(np.float64) - exact) <= bound), (m, n, k, layout, add)
                assert np.all(np.abs(want.astype(np.float64) - exact) <= bound)  # the oracle obeys the same bound
    # device resident, n = 2048: accuracy must be fp32-class, far better than plain tf32
    n = 2048
    A = rng.standard_normal((n, n)).astype(np.float32); B = rng.standard_normal((n, n)).astype(np.float32)
    dA = torch.from_numpy(A).to(cuda_dev); dB = torch.from_numpy(B).to(cuda_dev)
    dC = torch.empty((n, n), dtype=torch.float32, device=cuda_dev)
    la.matmul(dC, la.Accum.Replace, dA, dB, 1.0)
    exact = A.astype(np.float64) @ B.astype(np.float64)
    err = np.abs(dC.cpu().numpy().astype(np.float64) - exact)
    assert np.all(err <= 4 * n * U32 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)))
    # fp32 accumulation over k = 2048 gives ~sqrt(k) u32 sum|a||b| ~ 4e-3 absolute here (measured 4.0e-3, i.e. 1.6e-5 of
    # max|C|); un-compensated tf32 inputs would give ~2e-2 (9e-5 of max|C|)
    assert err.max() / np.abs(exact).max() < 4e-5


def test_f32_triangular_structures(fb, oracle):
    la = fb.linalg
    rng = np.random.default_rng(42)
    for ds, ls, rs in [(0, 5, 0), (1, 0, 0), (0, 0, 6), (6, 6, 5), (2, 1, 2), (3, 4, 0), (5, 0, 0), (0, 2, 1)]:
        n = 150
        A = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32))
        B = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32))
        C0 = np.asfortranarray(rng.standard_normal((n, n)).astype(np.float32))
        want = C0.copy(order="F"); oracle.matmul_triangular(want, ds, True, A, ls, B, rs, 0.5)
        got = C0.copy(order="F"); la.matmul_triangular(got, ds, la.Accum.Add, A, ls, B, rs, 0.5)
        assert np.all(np.abs(got - want) <= 1e-4 * np.maximum(1.0, np.abs(want))), (ds, ls, rs)
        if ds:
            sel = np.tril(np.ones((n, n), bool)) if ds in (1, 3, 5) else np.triu(np.ones((n, n), bool))
            if ds >= 3:
                np.fill_diagonal(sel, False)
            assert np.array_equal(got[~sel], C0[~sel])


def test_f32_tcgen05_path_layouts(fb, cuda_dev):
    """Products large enough for the tcgen05 kernel (gemm_f32_tc.cuh), every operand layout / stride sign the packing pass
    must absorb, Replace and Add, against the fp32 forward bound with an f64 reference."""
    import torch
    la = fb.linalg
    rng = np.random.default_rng(43)
    m, n, k = 520, 390, 700  # 2 m n k = 2.8e8: above the dispatch threshold; ragged against the 128 / 128 / 32 tiles
    A = rng.standard_normal((m, k)).astype(np.float32); B = rng.standard_normal((k, n)).astype(np.float32)
    C0 = rng.standard_normal((m, n)).astype(np.float32)
    exact_ab = A.astype(np.float64) @ B.astype(np.float64)
    absab = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    dA = torch.from_numpy(A).to(cuda_dev); dB = torch.from_numpy(B).to(cuda_dev)
    views_a = {"row-major": dA, "col-major": dA.T.contiguous().T, "reversed rows": torch.flip(dA, [0]).contiguous().flip(0)}
    views_b = {"row-major": dB, "col-major": dB.T.contiguous().T, "reversed cols": torch.flip(dB, [1]).contiguous().flip(1)}
    for na, va in views_a.items():
        for nb_, vb in views_b.items():
            for add, alpha in [(False, 1.0), (True, -0.5)]:
                for cmaj in ("row", "col"):
                    dC = torch.from_numpy(C0).to(cuda_dev)
                    if cmaj == "col":
                        dC = dC.T.contiguous().T
                    if not add:
                        dC.fill_(float("nan"))
                    la.matmul(dC, la.Accum.Add if add else la.Accum.Replace, va, vb, alpha)
                    exact = alpha * exact_ab + (C0.astype(np.float64) if add else 0.0)
                    bound = 4 * k * U32 * abs(alpha) * absab + 4 * U32 * np.abs(exact) + (4 * U32 * np.abs(C0) if add else 0.0)
                    err = np.abs(dC.cpu().numpy().astype(np.float64) - exact)
                    assert np.all(err <= bound), (na, nb_, add, cmaj, float((err / bound).max()))


@pytest.mark.parametrize("shape", [(40, 30, 50), (520, 390, 700)])  # mma.sync path / tcgen05 path
def test_f32_matmul_nonfinite_inputs_stay_confined(fb, shape):
    """Contract for non-finite inputs (numerical contract): the compensated product a_hi*b_lo of an infinite
    a_hi with a signed correction term is -inf or NaN, so a row of A holding +inf yields non-finite (inf OR NaN) entries in
    exactly that row of C, never a finite wrong value, and every other row is unaffected."""
    la = fb.linalg
    m, n, k = shape
    rng = np.random.default_rng(44)
    A = rng.uniform(0.5, 1.0, (m, k)).astype(np.float32); B = rng.uniform(0.5, 1.0, (k, n)).astype(np.float32)
    A[3, 7] = np.inf
    C = np.full((m, n), np.nan, dtype=np.float32)
    la.matmul(C, la.Accum.Replace, A, B, 1.0)
    assert not np.any(np.isfinite(C[3, :]))
    rest = np.delete(C, 3, axis=0)
    assert np.all(np.isfinite(rest))
    want = np.delete(A, 3, axis=0).astype(np.float64) @ B.astype(np.float64)
    assert np.abs(rest - want).max() <= 1e-4 * np.abs(want).max()


def test_c32_matmul_vs_oracle(fb, oracle):
    """c32 matmul / triangular matmul (4M formulation on the f32 kernels; the large case takes the tcgen05 kernel with
    stride-2 planes through its packing pass) vs the oracle's c32 schoolbook product, inside the fp32 forward bound."""
    la = fb.linalg
    rng = np.random.default_rng(45)

    def crandn(shape, order):
        return np.array(rng.standard_normal(shape) + 1j * rng.standard_normal(shape), dtype=np.complex64, order=order)

    for (m, n, k) in [(1, 1, 1), (17, 17, 17), (127, 129, 65), (16, 1, 16), (4, 4, 0), (100, 63, 9), (520, 390, 700)]:
        for layout in [("F", "F", "F"), ("C", "F", "C")]:
            for add, alpha in [(False, 1.0), (True, 0.5 - 2.0j), (False, 1.5j)]:
                A = crandn((m, k), layout[0]); B = crandn((k, n), layout[1]); C0 = crandn((m, n), layout[2])
                got = C0.copy(order="K")
                if not add:
                    got[...] = np.nan
                la.matmul(got, la.Accum.Add if add else la.Accum.Replace, A, B, alpha)
                exact = alpha * (A.astype(np.complex128) @ B.astype(np.complex128)) + (C0.astype(np.complex128) if add else 0)
                bound = 16 * abs(alpha) * max(k, 1) * U32 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)) \
                    + 8 * U32 * np.abs(exact) + (8 * U32 * np.abs(C0) if add else 0.0) + 1e-30
                assert np.all(np.abs(got.astype(np.complex128) - exact) <= bound), (m, n, k, layout, add, alpha)
                if m * n * k <= 200 * 200 * 200:
                    want = C0.copy(order="K")
                    if not add:
                        want[...] = np.nan
                    oracle.matmul(want, add, A, B, alpha)
                    assert np.all(np.abs(want.astype(np.complex128) - exact) <= bound)
    for ds, ls, rs in [(0, 5, 0), (1, 0, 0), (6, 6, 5), (3, 4, 0)]:
        n = 70
        A = crandn((n, n), "F"); B = crandn((n, n), "F"); C0 = crandn((n, n), "F")
        want = C0.copy(order="F"); oracle.matmul_triangular(want, ds, True, A, ls, B, rs, 0.5 + 1j)
        got = C0.copy(order="F"); la.matmul_triangular(got, ds, la.Accum.Add, A, ls, B, rs, 0.5 + 1j)
        assert np.all(np.abs(got - want) <= 2e-4 * np.maximum(1.0, np.abs(want))), (ds, ls, rs)

