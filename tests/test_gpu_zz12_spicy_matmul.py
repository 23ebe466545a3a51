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


def test_inner_seam_gemm(fb):
    """`faer_b200_gemm` (the parameter list of private_gemm_x86::gemm at matmul/mod.rs:1373-1411, triangular.rs:641-680,
    internal/mod.rs:143-201): scatter through u32 / u64 indices, strided diagonal, DstKind Lower / Upper / Full on f64 against the
    definition; the plain product for f32 / c32 / c64 with conjugation flags against numpy."""
    la = fb.linalg
    rng = np.random.default_rng(140)
    for (m, n, k), kind, itype, add in itertools.product([(40, 40, 9), (130, 130, 33), (70, 45, 20)],
                                                         [la.GemmDstKind.Lower, la.GemmDstKind.Upper, la.GemmDstKind.Full],
                                                         [np.uint32, np.uint64], [False, True]):
        if kind != la.GemmDstKind.Full and m != n:
            continue
        ri = rng.permutation(m + 9)[:m].astype(itype)
        ci = rng.permutation(n + 4)[:n].astype(itype)
        A = np.asfortranarray(rng.standard_normal((m, k)))
        B = rng.standard_normal((k, n))                      # row-major rhs: strides are part of the seam
        Dbuf = rng.standard_normal(2 * k); D = Dbuf[::2]     # diag_stride = 2
        C0 = np.asfortranarray(rng.standard_normal((m + 9, n + 4)))
        blk = {la.GemmDstKind.Lower: S_LOW, la.GemmDstKind.Upper: S_UP, la.GemmDstKind.Full: S_RECT}[kind]
        want, touched = reference(C0, blk, ri, ci, add, A, B, D, 0.75)
        got = C0.copy(order="F")
        la.gemm(got, ri, ci, kind, 1 if add else 0, A, False, D, B, False, 0.75)
        key = (m, n, k, kind, itype.__name__, add)
        assert np.array_equal(got[~touched], C0[~touched]), key
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12), key
    # plain products of the other scalar types, with the conjugation flags of the seam
    for dt, tol, (m, n, k) in [(np.float32, 2e-4, (96, 70, 50)), (np.complex64, 5e-4, (96, 70, 50)), (np.complex128, 1e-11, (96, 70, 50)),
                               (np.complex128, 1e-11, (530, 512, 260))]:   # the last one: planar-operand path of gemm_c64.cu
        cplx = np.issubdtype(dt, np.complexfloating)
        mk = (lambda *sh: (rng.standard_normal(sh) + 1j * rng.standard_normal(sh)).astype(dt)) if cplx else (lambda *sh: rng.standard_normal(sh).astype(dt))
        A = np.asfortranarray(mk(m, k)); B = np.asfortranarray(mk(k, n)); C0 = np.asfortranarray(mk(m, n))
        for cl, cr in ([(False, False), (True, False), (False, True), (True, True)] if cplx else [(False, False)]):
            alpha = (0.5 - 0.25j) if cplx else 0.5
            got = C0.copy(order="F")
            la.gemm(got, None, None, la.GemmDstKind.Full, 1, A, cl, None, B, cr, alpha)
            want = C0 + alpha * ((A.conj() if cl else A).astype(np.complex128 if cplx else np.float64) @ (B.conj() if cr else B))
            assert np.allclose(got, want, rtol=tol, atol=tol * k), (dt.__name__, cl, cr)
