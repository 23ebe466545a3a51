"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI (ctypes -> libfaer_b200.so),
against the CPU oracle on the same seeded inputs.

Tolerances (derived from the reference's own tests):
  GEMM      |dC_ij| <= 2 k u sum_k |a_ik||b_kj|   (forward bound valid for any summation order);
            reference analogue: abs 1e-10 at k < 100 (matmul/mod.rs:2146-2148)
  LLT       status tag / NonPositivePivot.index / regularisation count EXACT; |A - L L^T| <= 8 n 128 u |A|_max
  LU        perm_fwd, perm_inv, transposition count BIT-EXACT; |P A - L U| <= 8 n 128 u |A|_max
  TRSM      residual eps*128*8n (cholesky/llt/solve.rs tolerance)
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

U = np.finfo(np.float64).eps / 2
S_RECT, S_LOW, S_UP, S_SLOW, S_SUP, S_ULOW, S_UUP = range(7)
MATMUL_SHAPES = [(2, 2, 2), (4, 4, 4), (8, 8, 8), (16, 16, 16), (127, 127, 127), (128, 128, 128), (129, 129, 129),
                 (15, 15, 15), (17, 17, 17), (1, 1, 1), (1, 16, 16), (16, 1, 16), (16, 16, 1), (0, 4, 4), (4, 0, 4),
                 (4, 4, 0), (63, 9, 100), (100, 63, 9), (256, 256, 256), (300, 520, 260)]


def to_dev(a, dev):
    """torch CUDA tensor with the same logical layout (strides) as the numpy array `a` (non-negative strides)."""
    import torch
    if a.size == 0:
        return torch.empty(a.shape, dtype=torch.float64, device=dev)
    if a.flags.c_contiguous:
        return torch.from_numpy(a).to(dev)
    if a.flags.f_contiguous:
        return torch.from_numpy(np.ascontiguousarray(a.T)).to(dev).T
    base = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return base


def gemm_bound(A, B, k):
    return 2.0 * max(k, 1) * (2 * U) * (np.abs(A) @ np.abs(B)) + 1e-300


def test_matmul_vs_oracle_host_and_device(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(0)
    for (m, n, k) in MATMUL_SHAPES:
        for layout in itertools.product("CF", repeat=3):
            for add, alpha in [(False, 1.0), (True, -1.0), (True, 0.5)]:
                A = np.array(rng.standard_normal((m, k)), order=layout[0])
                B = np.array(rng.standard_normal((k, n)), order=layout[1])
                C0 = np.array(rng.standard_normal((m, n)), order=layout[2])
                want = C0.copy(order="K")
                if not add:
                    want[...] = np.nan  # Replace must not read dst
                oracle.matmul(want, add, A, B, alpha)
                bound = abs(alpha) * gemm_bound(A, B, k) + (2 * U) * np.abs(want) * 2
                # host path (staged)
                got = C0.copy(order="K")
                if not add:
                    got[...] = np.nan
                la.matmul(got, la.Accum.Add if add else la.Accum.Replace, A, B, alpha)
                assert np.all(np.abs(got - want) <= bound), ("host", m, n, k, layout, add)
                # device path (in place)
                dA, dB, dC = to_dev(A, cuda_dev), to_dev(B, cuda_dev), to_dev(C0.copy(order="K"), cuda_dev)
                if not add and dC.numel():
                    dC.fill_(float("nan"))
                la.matmul(dC, la.Accum.Add if add else la.Accum.Replace, dA, dB, alpha)
                assert np.all(np.abs(dC.cpu().numpy() - want) <= bound), ("dev", m, n, k, layout, add)


def test_matmul_strided_and_reversed_views(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(1)
    A = rng.standard_normal((200, 300)); B = rng.standard_normal((300, 150)); Cm = rng.standard_normal((400, 300))
    for (Av, Bv, sl) in [
        (A[::2, ::-1][:, :100], B[:100, ::3], (slice(1, None, 4), slice(None, None, 6))),
        (A[::-1, :77][:90], B[76::-1, 10:60], (slice(399, 39, -4), slice(0, 50))),
        (A[3:103:1, 5:105:1], B[100:200, 50:0:-1], (slice(0, 100), slice(299, 249, -1))),
    ]:
        Cv = Cm[sl]
        Cv = Cv[:Av.shape[0], :Bv.shape[1]]
        want = Cv.copy()
        oracle.matmul(want, True, np.ascontiguousarray(Av), np.ascontiguousarray(Bv), 2.0)
        la.matmul(Cv, la.Accum.Add, Av, Bv, 2.0)
        assert np.all(np.abs(Cv - want) <= 2.0 * gemm_bound(Av, Bv, Av.shape[1]) + 1e-14)


def test_triangular_all_structures_vs_oracle(fb, oracle):
    """reference test_triangular (matmul/mod.rs:2106-2266): all 7^3 structure combos, dims < 100 plus larger
    multi-tile sizes, tol 1e-10, and the unselected part of dst preserved bit-for-bit."""
    la = fb.linalg
    rng = np.random.default_rng(2)
    combos = list(itertools.product(range(7), repeat=3))
    for it, (ds, ls, rs) in enumerate(combos):
        n = int(rng.integers(1, 100)) if it % 5 else int(rng.integers(130, 300))
        m = nn = k = n
        if ls == S_RECT and rs == S_RECT:
            k = int(rng.integers(1, 100))
        if ds == S_RECT and ls == S_RECT:
            m = int(rng.integers(1, 100))
        if ds == S_RECT and rs == S_RECT:
            nn = int(rng.integers(1, 100))
        A = np.asfortranarray(rng.standard_normal((m, k)))
        B = np.asfortranarray(rng.standard_normal((k, nn)))
        C0 = np.asfortranarray(rng.standard_normal((m, nn)))
        for add in (False, True):
            want = C0.copy(order="F")
            oracle.matmul_triangular(want, ds, add, A, ls, B, rs, 2.5)
            got = C0.copy(order="F")
            la.matmul_triangular(got, ds, la.Accum.Add if add else la.Accum.Replace, A, ls, B, rs, 2.5)
            d = np.abs(got - want)
            ok = (d <= 1e-10) | (d <= 1e-10 * np.maximum(np.abs(got), np.abs(want)))
            assert np.all(ok), (ds, ls, rs, m, nn, k, add, float(d.max()))
            if ds != S_RECT:
                sel = np.tril(np.ones((m, nn), bool)) if ds in (S_LOW, S_SLOW, S_ULOW) else np.triu(np.ones((m, nn), bool))
                if ds >= S_SLOW:
                    np.fill_diagonal(sel, False)
                assert np.array_equal(got[~sel], C0[~sel])


def test_triangular_solve_vs_oracle(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(3)
    eps = np.finfo(np.float64).eps
    fns = {(True, False): la.solve_lower_triangular_in_place, (True, True): la.solve_unit_lower_triangular_in_place,
           (False, False): la.solve_upper_triangular_in_place, (False, True): la.solve_unit_upper_triangular_in_place}
    for n, k in [(0, 3), (1, 1), (2, 3), (4, 5), (5, 5), (17, 3), (33, 70), (50, 20), (128, 300), (129, 70), (257, 64),
                 (600, 130)]:
        T = np.asfortranarray(rng.standard_normal((n, n)) / max(n, 1) + 2 * np.eye(n))
        for lower, unit in itertools.product((True, False), (True, False)):
            for order in "FC":
                Bm = np.array(rng.standard_normal((n, k)), order=order)
                want = Bm.copy(order="K")
                oracle.solve_triangular(T, want, lower, unit)
                got = Bm.copy(order="K")
                fns[(lower, unit)](T, got)
                tol = eps * 128 * 8 * max(n, 1) * max(1.0, float(np.abs(want).max()) if want.size else 1.0)
                assert np.all(np.abs(got - want) <= tol), (n, k, lower, unit, order)
                if n:
                    dT, dB = to_dev(T, cuda_dev), to_dev(Bm.copy(order="K"), cuda_dev)
                    fns[(lower, unit)](dT, dB)
                    assert np.all(np.abs(dB.cpu().numpy() - want) <= tol), ("dev", n, k, lower, unit, order)


def spd(rng, n):
    G = rng.standard_normal((n, n))
    return np.asfortranarray(G @ G.T + n * np.eye(n))


def test_llt_vs_oracle(fb, oracle, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(4)
    for n in [0, 1, 2, 3, 4, 8, 31, 63, 64, 65, 127, 128, 129, 240, 255, 256, 257, 500, 1000]:
        A = spd(rng, n)
        want = A.copy(order="F")
        fail, cnt = oracle.llt(want)
        assert fail == -1
        got = A.copy(order="F")
        info = la.cholesky_in_place(got)
        assert info.dynamic_regularization_count == cnt == 0
        assert np.array_equal(np.triu(got, 1), np.triu(A, 1))  # strict upper triangle untouched
        L = np.tril(got)
        amax = np.abs(A).max() if n else 1.0
        assert np.all(np.abs(L @ L.T - A) <= 8 * max(n, 1) * 128 * U * amax), n
        # factors agree with the oracle to a (generous, kappa-aware) elementwise tolerance
        assert np.allclose(L, np.tril(want), rtol=1e-10, atol=1e-10 * np.sqrt(amax)), n
        if n:
            dA = to_dev(A.copy(order="F"), cuda_dev)
            la.cholesky_in_place(dA)
            assert np.array_equal(dA.cpu().numpy(), got), n  # host-staged and in-place device paths are identical
    # row-major input (faer falls back to the scalar leaf; same math)
    A = spd(rng, 200)
    got = np.ascontiguousarray(A)
    la.cholesky_in_place(got)
    L = np.tril(got)
    assert np.all(np.abs(L @ L.T - A) <= 8 * 200 * 128 * U * np.abs(A).max())


def test_llt_leaf_bitwise_matches_oracle(fb, oracle):
    """n <= 64 is a single leaf in the reference; our leaf performs the same FMA chain => identical bits."""
    la = fb.linalg
    rng = np.random.default_rng(5)
    for n in [1, 2, 5, 16, 33, 64]:
        A = spd(rng, n)
        want = A.copy(order="F"); oracle.llt(want)
        got = A.copy(order="F"); la.cholesky_in_place(got)
        assert np.array_equal(got, want), n


def test_llt_error_index_and_regularization(fb, oracle):
    la = fb.linalg
    rng = np.random.default_rng(6)
    for n, bad in [(10, 0), (10, 9), (200, 150), (300, 64), (300, 255), (700, 513)]:
        A = spd(rng, n)
        A[bad, bad] = -1.0
        want_fail, _ = oracle.llt(A.copy(order="F"))
        with pytest.raises(la.LltError) as ei:
            la.cholesky_in_place(A.copy(order="F"))
        assert ei.value.index == want_fail == bad
    A = spd(rng, 20); A[7, 7] = np.nan
    with pytest.raises(la.LltError) as ei:
        la.cholesky_in_place(A.copy(order="F"))
    assert ei.value.index == 7
    A = np.asfortranarray(np.diag([4.0, 1e-20, 9.0, -3.0]))
    want = A.copy(order="F"); fail, cnt = oracle.llt(want, delta=1.0, eps=1e-10)
    got = A.copy(order="F"); info = la.cholesky_in_place(got, regularization=(1.0, 1e-10))
    assert fail == -1 and info.dynamic_regularization_count == cnt == 2
    assert np.array_equal(got, want)


def unpack_lu(LU):
    m, n = LU.shape
    size = min(m, n)
    return np.tril(LU[:, :size], -1) + np.eye(m, size), np.triu(LU[:size, :])


@pytest.mark.parametrize("idx", [np.uint64, np.uint32])
def test_plu_vs_oracle(fb, oracle, cuda_dev, idx):
    """reference test_plu (lu/partial_pivoting/factor.rs:304-404) shapes + larger ones; permutations bit-exact."""
    la = fb.linalg
    rng = np.random.default_rng(7)
    shapes = [(n, n) for n in [1, 2, 3, 16, 17, 63, 64, 65, 128, 255, 256, 257, 700, 1500]]
    shapes += [(m, 8) for m in [8, 128, 255, 256, 257]] + [(3000, 100), (40, 100), (129, 500)]
    for (m, n) in shapes:
        A = np.asfortranarray(rng.standard_normal((m, n)))
        want = A.copy(order="F")
        perm_o, pinv_o, nt_o = oracle.lu(want)
        got = A.copy(order="F")
        perm = np.zeros(m, dtype=idx); pinv = np.zeros(m, dtype=idx)
        info = la.lu_in_place(got, perm, pinv)
        assert np.array_equal(perm.astype(np.int64), perm_o), (m, n)
        assert np.array_equal(pinv.astype(np.int64), pinv_o), (m, n)
        assert info.transposition_count == nt_o, (m, n)
        L, Um = unpack_lu(got)
        amax = np.abs(A).max()
        growth = max(1.0, np.abs(Um).max() / amax)
        assert np.all(np.abs(L @ Um - A[perm_o, :]) <= 8 * max(m, n) * 128 * U * amax * growth), (m, n)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * growth), (m, n)
    # device-resident matrix + device permutation arrays
    import torch
    A = np.asfortranarray(rng.standard_normal((900, 900)))
    perm_o, pinv_o, nt_o = oracle.lu(A.copy(order="F"))
    dA = to_dev(A.copy(order="F"), cuda_dev)
    tdt = torch.int64 if idx == np.uint64 else torch.int32
    dp = torch.zeros(900, dtype=tdt, device=cuda_dev); dpi = torch.zeros(900, dtype=tdt, device=cuda_dev)
    info = la.lu_in_place(dA, dp, dpi)
    assert np.array_equal(dp.cpu().numpy().astype(np.int64), perm_o)
    assert np.array_equal(dpi.cpu().numpy().astype(np.int64), pinv_o)
    assert info.transposition_count == nt_o


def test_lu_pivot_rule_ties_zero_columns(fb, oracle):
    la = fb.linalg
    # ties -> lowest row index; all-zero column keeps imax = row and yields inf/nan without an error
    T = np.asfortranarray(np.array([[1.0, 2.0, 0.5], [-1.0, 5.0, 1.0], [1.0, 7.0, 3.0], [0.5, 1.0, 2.0]]))
    want = T.copy(order="F"); perm_o, _, nt_o = oracle.lu(want)
    got = T.copy(order="F"); p = np.zeros(4, np.uint64); pi = np.zeros(4, np.uint64)
    info = la.lu_in_place(got, p, pi)
    assert np.array_equal(p.astype(np.int64), perm_o) and info.transposition_count == nt_o
    assert np.allclose(got, want, rtol=1e-14, atol=0)
    Z = np.asfortranarray(np.array([[0.0, 1.0], [0.0, 2.0]]))
    want = Z.copy(order="F"); perm_o, _, _ = oracle.lu(want)
    got = Z.copy(order="F"); p = np.zeros(2, np.uint64); pi = np.zeros(2, np.uint64)
    la.lu_in_place(got, p, pi)
    assert np.array_equal(p.astype(np.int64), perm_o)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))


# ---- BASELINE.json full sizes: size-independent properties (the oracle would take minutes here) ----
def test_llt_full_size_property(fb, cuda_dev):
    """config[1]: f64 LLT n=16384. Check A x = L (L^T x) on random probes + upper triangle untouched."""
    import torch
    la = fb.linalg
    n = 16384
    torch.manual_seed(0)
    G = torch.randn((n, n), dtype=torch.float64, device=cuda_dev)
    A0 = torch.addmm(n * torch.eye(n, dtype=torch.float64, device=cuda_dev), G, G.T).T  # column-major view, symmetric
    del G
    A = A0.clone(memory_format=torch.preserve_format)
    info = la.cholesky_in_place(A)
    assert info.dynamic_regularization_count == 0
    assert torch.equal(torch.triu(A, 1), torch.triu(A0, 1))
    L = torch.tril(A)
    x = torch.randn((n, 8), dtype=torch.float64, device=cuda_dev)
    r = A0 @ x - L @ (L.T @ x)
    scale = float(A0.abs().max()) * float(x.abs().max()) * n
    assert float(r.abs().max()) <= 128 * U * scale


def test_lu_half_size_property(fb, cuda_dev):
    """config[2] at n=16384 on one GPU (n=32768 is exercised by bench/scale runs): P A x = L (U x) on probes,
    perm is a permutation, perm_inv its inverse."""
    import torch
    la = fb.linalg
    n = 16384
    torch.manual_seed(1)
    A0 = torch.randn((n, n), dtype=torch.float64, device=cuda_dev).T
    A = A0.clone(memory_format=torch.preserve_format)
    p = torch.zeros(n, dtype=torch.int64, device=cuda_dev); pi = torch.zeros(n, dtype=torch.int64, device=cuda_dev)
    la.lu_in_place(A, p, pi)
    assert torch.equal(torch.sort(p).values, torch.arange(n, device=cuda_dev))
    assert torch.equal(pi[p], torch.arange(n, device=cuda_dev))
    x = torch.randn((n, 4), dtype=torch.float64, device=cuda_dev)
    Lx = torch.tril(A, -1) @ (torch.triu(A) @ x) + torch.triu(A) @ x
    r = A0[p, :] @ x - Lx
    growth = max(1.0, float(torch.triu(A).abs().max()) / float(A0.abs().max()))
    assert float(r.abs().max()) <= 128 * 8 * U * n * float(A0.abs().max()) * float(x.abs().max()) * growth
    assert float(torch.tril(A, -1).abs().max()) <= 1.0  # partial pivoting bounds the multipliers


def test_c64_matmul_vs_oracle(fb, oracle, cuda_dev):
    """c64 matmul / triangular matmul (4M formulation on the f64 DMMA kernel) vs the oracle: shapes of the reference's
    test_matmul, complex alpha, all dst structures; tolerance = the forward bound of appendix B on |a||b|."""
    import itertools
    la = fb.linalg
    rng = np.random.default_rng(21)

    def crandn(shape, order):
        return np.array(rng.standard_normal(shape) + 1j * rng.standard_normal(shape), order=order)

    # (600, 520, 300): large enough for the planar-operand path on the TMA-fed kernel (gemm_c64.cu)
    for (m, n, k) in [(1, 1, 1), (2, 2, 2), (17, 17, 17), (127, 129, 65), (128, 128, 128), (16, 1, 16), (4, 4, 0), (100, 63, 9), (600, 520, 300)]:
        for layout in [("F", "F", "F"), ("C", "F", "C"), ("F", "C", "F")]:
            for add, alpha in [(False, 1.0), (True, -1.0), (True, 0.5 - 2.0j), (False, 1.5j)]:
                A = crandn((m, k), layout[0]); B = crandn((k, n), layout[1]); C0 = crandn((m, n), layout[2])
                want = C0.copy(order="K")
                if not add:
                    want[...] = np.nan
                oracle.matmul(want, add, A, B, alpha)
                got = C0.copy(order="K")
                if not add:
                    got[...] = np.nan
                la.matmul(got, la.Accum.Add if add else la.Accum.Replace, A, B, alpha)
                bound = 8 * abs(alpha) * max(k, 1) * (2 * U) * (np.abs(A) @ np.abs(B)) + 8 * U * np.abs(want) + 1e-300
                assert np.all(np.abs(got - want) <= bound), (m, n, k, layout, add, alpha)
    # structured: a few representative combos incl. unit diagonals (real part 1, imaginary part 0)
    for ds, ls, rs in [(0, 5, 0), (1, 0, 0), (0, 0, 6), (6, 6, 5), (2, 1, 2), (3, 4, 0)]:
        n = 70
        A = crandn((n, n), "F"); B = crandn((n, n), "F"); C0 = crandn((n, n), "F")
        want = C0.copy(order="F"); oracle.matmul_triangular(want, ds, True, A, ls, B, rs, 0.5 + 1j)
        got = C0.copy(order="F"); la.matmul_triangular(got, ds, la.Accum.Add, A, ls, B, rs, 0.5 + 1j)
        assert np.all(np.abs(got - want) <= 1e-10 * np.maximum(1.0, np.abs(want))), (ds, ls, rs)
    # device-resident c64, n = 1024
    import torch
    n = 1024
    A = crandn((n, n), "F"); B = crandn((n, n), "F")
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda_dev).T
    dB = torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda_dev).T
    dC = torch.empty((n, n), dtype=torch.complex128, device=cuda_dev).T
    la.matmul(dC, la.Accum.Replace, dA, dB, 1.0)
    ref = A @ B
    assert np.all(np.abs(dC.cpu().numpy() - ref) <= 8 * n * (2 * U) * (np.abs(A) @ np.abs(B)))


def test_plu_large_path_pivots_match_oracle(fb, oracle):
    """n >= 4096 takes the block-column look-ahead driver (SM-partitioned, cluster panel kernel; dist.cu / lu_f64.cu): its
    permutation must still be the reference's, bit for bit, and the factors must agree with the oracle's."""
    la = fb.linalg
    n = 4096 + 40  # ragged last block
    rng = np.random.default_rng(77)
    A = np.asfortranarray(rng.standard_normal((n, n)))
    want = A.copy(order="F")
    perm_o, pinv_o, nt_o = oracle.lu(want)
    got = A.copy(order="F")
    perm = np.zeros(n, dtype=np.uint64); pinv = np.zeros(n, dtype=np.uint64)
    info = la.lu_in_place(got, perm, pinv)
    assert np.array_equal(perm.astype(np.int64), perm_o)
    assert np.array_equal(pinv.astype(np.int64), pinv_o)
    assert info.transposition_count == nt_o
    growth = max(1.0, np.abs(np.triu(want)).max() / np.abs(A).max())
    assert np.allclose(got, want, rtol=1e-8, atol=1e-8 * growth)
