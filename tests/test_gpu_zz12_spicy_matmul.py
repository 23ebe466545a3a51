"""GPU test of `spicy_matmul` (faer/src/linalg/matmul/internal/mod.rs:45-379) through `faer_b200_spicy_matmul_f64`:
C[row_idx[i], col_idx[j]] (+)= alpha (A diag(D) B)[i, j], masked by the block structure of the product — against the
definition evaluated with numpy (the reference's own fallback composition, internal/mod.rs:206-379: scale the columns of A,
structured product into a temporary, masked scatter), for both code paths: the warp-specialised kernel with the diagonal
folded into the lhs fragments and the scatter into the store (large, TMA-readable operands; forced with gemm_ws = 2 for the
small ones too), and the composition on the cp.async kernel (gemm_ws = 0). Untouched entries of C are compared bit for bit."""
import itertools

import numpy as np
import pytest

from test_gpu_parity import S_LOW, S_RECT, S_SLOW, S_SUP, S_ULOW, S_UP, S_UUP, U

pytestmark = pytest.mark.gpu


def reference(C0, blk, ri, ci, add, A, B, D, alpha):
    m, n = A.shape[0], B.shape[1]
    P = alpha * ((A * D[None, :]) if D is not None else A) @ B
    i, j = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
    keep = np.ones((m, n), bool)
    if blk in (S_LOW, S_SLOW, S_ULOW):
        keep = i >= j if blk == S_LOW else i > j
    if blk in (S_UP, S_SUP, S_UUP):
        keep = i <= j if blk == S_UP else i < j
    rr = np.arange(m) if ri is None else ri.astype(np.int64)
    cc = np.arange(n) if ci is None else ci.astype(np.int64)
    want = C0.copy()
    touched = np.zeros(C0.shape, bool)
    for a in range(m):
        for b in range(n):
            if keep[a, b]:
                want[rr[a], cc[b]] = (want[rr[a], cc[b]] if add else 0.0) + P[a, b]
                touched[rr[a], cc[b]] = True
    return want, touched


@pytest.mark.parametrize("ws", [0, 2])
def test_spicy_matmul_vs_definition(fb, ws):
    la = fb.linalg
    lib = fb.load()
    saved = lib.faer_b200_get_option(b"gemm_ws")
    lib.faer_b200_set_option(b"gemm_ws", ws)
    try:
        rng = np.random.default_rng(130 + ws)
        for (m, n, k) in [(5, 7, 3), (64, 64, 16), (130, 70, 33), (300, 300, 40), (200, 257, 129)]:
            for blk, use_ri, use_ci, use_d, add in itertools.product([S_RECT, S_LOW, S_SUP, S_ULOW, S_UP], [False, True], [False, True],
                                                                    [False, True], [False, True]):
                if blk != S_RECT and m != n and not (use_ri or use_ci):
                    pass  # non-square structured products are legal (internal/mod.rs:264-300)
                R = m + 11 if use_ri else m
                Cc = n + 5 if use_ci else n
                ri = rng.permutation(R)[:m].astype(np.uint64) if use_ri else None
                ci = rng.permutation(Cc)[:n].astype(np.uint64) if use_ci else None
                A = np.asfortranarray(rng.standard_normal((m, k)))
                B = np.asfortranarray(rng.standard_normal((k, n)))
                D = rng.standard_normal(k) if use_d else None
                C0 = np.asfortranarray(rng.standard_normal((R, Cc)))
                want, touched = reference(C0, blk, ri, ci, add, A, B, D, -0.5)
                got = C0.copy(order="F")
                la.spicy_matmul(got, blk, ri, ci, la.Accum.Add if add else la.Accum.Replace, A, B, D, -0.5)
                key = (m, n, k, blk, use_ri, use_ci, use_d, add)
                assert np.array_equal(got[~touched], C0[~touched]), key
                bound = 4 * k * 2 * U * 0.5 * (np.abs(A) * (np.abs(D)[None, :] if D is not None else 1.0)) @ np.abs(B)
                full = np.zeros(C0.shape); 
                rr = np.arange(m) if ri is None else ri.astype(np.int64); cc = np.arange(n) if ci is None else ci.astype(np.int64)
                full[np.ix_(rr, cc)] = bound
                assert np.all(np.abs(got - want)[touched] <= (full + 4 * U * np.abs(want))[touched]), key
    finally:
        lib.faer_b200_set_option(b"gemm_ws", saved)
