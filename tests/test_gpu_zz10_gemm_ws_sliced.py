"""GPU parity tests of the two round-2 f64 GEMM kernels, through the C ABI, against the oracle:

  * the TMA-fed, warp-specialised DMMA kernel (csrc/gemm_f64_ws.cuh), forced on (`gemm_ws` = 2) so that small and ragged
    shapes reach it too: all operand layouts (MN-major / K-major shared-memory tiles), Replace (dst pre-filled with NaN: it
    must not be read, matmul/mod.rs:1580-1582) and Add, triangular destinations with the untouched part checked
    bit-for-bit (matmul/triangular.rs:641-680), sizes that are not multiples of the 128 x 64 x 16 tile;
    tolerance: the forward bound 2 k u (|A||B|)_ij of SURVEY appendix B;
  * the opt-in int8-sliced tcgen05 product (csrc/gemm_f64_sliced.cuh, `f64_gemm_mode` = 1): its stated contract
    |dC_ij| <= 16 u (|alpha||A||B|)_ij, u = 2^-53, including rows / columns with very different scales.
"""
import itertools

import numpy as np
import pytest

from test_gpu_parity import S_LOW, S_RECT, S_SLOW, S_SUP, S_ULOW, S_UP, S_UUP, U, gemm_bound, to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture
def opts(fb):
    lib = fb.load()
    saved = {k: lib.faer_b200_get_option(k) for k in (b"gemm_ws", b"f64_gemm_mode")}
    yield lib
    for k, v in saved.items():
        lib.faer_b200_set_option(k, v)


WS_SHAPES = [(128, 64, 16), (129, 65, 17), (256, 256, 256), (300, 520, 260), (1000, 900, 300), (1031, 517, 129), (2048, 2048, 64),
             (64, 1500, 40), (1500, 64, 2000), (130, 70, 1), (5, 3, 100)]


def test_ws_gemm_vs_oracle_all_layouts(fb, oracle, cuda_dev, opts):
    la = fb.linalg
    assert opts.faer_b200_set_option(b"gemm_ws", 2) == 0
    rng = np.random.default_rng(110)
    n0 = opts.faer_b200_launch_count()
    for (m, n, k) in WS_SHAPES:
        for layout in itertools.product("CF", repeat=3):
            for add, alpha in [(False, 1.0), (True, -0.75)]:
                A = np.array(rng.standard_normal((m, k)), order=layout[0])
                B = np.array(rng.standard_normal((k, n)), order=layout[1])
                C0 = np.array(rng.standard_normal((m, n)), order=layout[2])
                want = C0.copy(order="K")
                if not add:
                    want[...] = np.nan
                oracle.matmul(want, add, A, B, alpha)
                bound = abs(alpha) * gemm_bound(A, B, k) + (2 * U) * np.abs(want) * 2
                dA, dB, dC = to_dev(A, cuda_dev), to_dev(B, cuda_dev), to_dev(C0.copy(order="K"), cuda_dev)
                if not add:
                    dC.fill_(float("nan"))
                la.matmul(dC, la.Accum.Add if add else la.Accum.Replace, dA, dB, alpha)
                got = dC.cpu().numpy()
                assert np.all(np.abs(got - want) <= bound), (m, n, k, layout, add, float(np.nanmax(np.abs(got - want))))
    assert opts.faer_b200_launch_count() > n0


def test_ws_gemm_triangular_destinations(fb, oracle, cuda_dev, opts):
    la = fb.linalg
    assert opts.faer_b200_set_option(b"gemm_ws", 2) == 0
    rng = np.random.default_rng(111)
    for n, k in [(128, 16), (200, 33), (517, 129), (1030, 256), (1536, 70)]:
        for ds in (S_LOW, S_UP, S_SLOW, S_SUP, S_ULOW, S_UUP):
            for la_, lb_ in itertools.product("CF", repeat=2):
                A = np.array(rng.standard_normal((n, k)), order=la_)
                B = np.array(rng.standard_normal((k, n)), order=lb_)
                C0 = np.asfortranarray(rng.standard_normal((n, n)))
                for add in (False, True):
                    want = C0.copy(order="F")
                    oracle.matmul_triangular(want, ds, add, np.asfortranarray(A), S_RECT, np.asfortranarray(B), S_RECT, -1.0)
                    dA, dB, dC = to_dev(A, cuda_dev), to_dev(B, cuda_dev), to_dev(C0.copy(order="F"), cuda_dev)
                    la.matmul_triangular(dC, ds, la.Accum.Add if add else la.Accum.Replace, dA, S_RECT, dB, S_RECT, -1.0)
                    got = dC.cpu().numpy()
                    sel = np.tril(np.ones((n, n), bool)) if ds in (S_LOW, S_SLOW, S_ULOW) else np.triu(np.ones((n, n), bool))
                    if ds >= S_SLOW:
                        np.fill_diagonal(sel, False)
                    assert np.array_equal(got[~sel], C0[~sel]), (n, k, ds, la_, lb_, add)
                    bound = gemm_bound(A, B, k) + (4 * U) * np.abs(want)
                    assert np.all(np.abs(got - want)[sel] <= bound[sel]), (n, k, ds, la_, lb_, add)


def test_ws_gemm_syrk_property_large(fb, cuda_dev, opts):
    """The LLT trailing update at a BASELINE-like size (lower destination, lhs column-major, rhs = lhs^T): the kernel
    against the cp.async DMMA kernel (`gemm_ws` = 0) on the same device data — both sum the same products, in different
    orders — and the untouched strict upper triangle."""
    import torch
    la = fb.linalg
    torch.manual_seed(7)
    n, k = 6016, 256
    A = torch.randn((k, n), dtype=torch.float64, device=cuda_dev).T  # column-major n x k
    C0 = torch.randn((n, n), dtype=torch.float64, device=cuda_dev).T
    outs = []
    for mode in (0, 2):
        assert opts.faer_b200_set_option(b"gemm_ws", mode) == 0
        Cm = C0.clone(memory_format=torch.preserve_format)
        la.matmul_triangular(Cm, S_LOW, la.Accum.Add, A, S_RECT, A.T, S_RECT, -1.0)
        outs.append(Cm)
    assert torch.equal(torch.triu(outs[1], 1), torch.triu(C0, 1))
    absprod = A.abs() @ A.abs().T
    assert bool((torch.tril(outs[0] - outs[1]).abs() <= 2 * k * 2 * U * absprod).all())


def test_sliced_gemm_accuracy_contract(fb, oracle, cuda_dev, opts):
    la = fb.linalg
    assert opts.faer_b200_set_option(b"f64_gemm_mode", 1) == 0
    rng = np.random.default_rng(112)
    for (m, n, k, graded) in [(256, 256, 128, False), (512, 384, 1000, False), (700, 300, 2500, True), (1024, 1024, 4096, False)]:
        A = np.asfortranarray(rng.standard_normal((m, k)))
        B = np.asfortranarray(rng.standard_normal((k, n)))
        if graded:
            A *= np.exp(20 * rng.standard_normal((m, 1)))   # rows of A / columns of B on very different scales
            B *= np.exp(20 * rng.standard_normal((1, n)))
        C0 = np.asfortranarray(rng.standard_normal((m, n)))
        for add, alpha in [(False, 1.0), (True, -0.5)]:
            want = C0.copy(order="F")
            oracle.matmul(want, add, A, B, alpha)
            dA, dB, dC = to_dev(A, cuda_dev), to_dev(B, cuda_dev), to_dev(C0.copy(order="F"), cuda_dev)
            if not add:
                dC.fill_(float("nan"))
            la.matmul(dC, la.Accum.Add if add else la.Accum.Replace, dA, dB, alpha)
            got = dC.cpu().numpy()
            absprod = np.abs(A) @ np.abs(B)
            # the oracle's own k-ordered sum carries up to k u (|A||B|): the contract is checked against a long-double
            # reference on a sample of columns, the oracle comparison uses the ordinary forward bound
            cols = rng.choice(n, size=min(n, 40), replace=False)
            ref = alpha * (A.astype(np.longdouble) @ B[:, cols].astype(np.longdouble)) + (C0[:, cols] if add else 0)
            err = np.abs(got[:, cols].astype(np.longdouble) - ref).astype(np.float64)
            lim = 16 * U * abs(alpha) * absprod[:, cols] + 2 * U * np.abs(want[:, cols])
            assert np.all(err <= lim), (m, n, k, graded, add, float((err / absprod[:, cols]).max()))
            assert np.all(np.abs(got - want) <= abs(alpha) * gemm_bound(A, B, k) + 2 * U * np.abs(want)), (m, n, k, graded, add)
