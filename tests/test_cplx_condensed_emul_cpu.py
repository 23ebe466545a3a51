"""The launch sequences and per-thread bodies of the complex condensed-form kernels (csrc/cplx_condensed_core.cuh), compiled for
the host (tools/emul/cplx_condensed_host.cpp) and run thread by thread: same source as the CUDA build, the launcher replaced by
loops over the launch's index space. Checked against the numpy model (tests/cplx_condensed_model.py, itself pinned on the
oracle's restatement of the reference in test_cplx_condensed_model_cpu.py), in forward and reverse thread order (bitwise equal:
no intra-launch dependence), on strided / adjoint inputs, on inputs with zero columns (tau = inf reflectors) and on the
end-to-end identities with LAPACK standing in for the real condensed solvers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cplx_condensed_model as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U = np.finfo(np.float64).eps
I64 = C.c_longlong
P = C.c_void_p


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cce") / "libcce.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, os.path.join(ROOT, "tools", "emul", "cplx_condensed_host.cpp")])
    lib = C.CDLL(out)
    lib.cc_emul_tridiag.argtypes = [P, I64, I64, I64, P, P, P, P, P, P, C.c_int]
    lib.cc_emul_tridiag.restype = I64
    lib.cc_emul_bidiag.argtypes = [P, I64, I64, I64, I64, C.c_int, P, P, P, P, P, P, P, P, P, C.c_int]
    lib.cc_emul_bidiag.restype = I64
    lib.cc_emul_scale_rows_embed.argtypes = [P, I64, I64, P, P, I64, I64, C.c_int]
    lib.cc_emul_transpose_corner.argtypes = [P, I64, P, I64, C.c_int]
    lib.cc_emul_copy_out_f64.argtypes = [P, I64, I64, P, I64, I64, C.c_int]
    lib.cc_emul_copy_out_f32.argtypes = [P, I64, I64, P, I64, I64, C.c_int]
    lib.cc_emul_copy_values_f64.argtypes = [P, I64, P, I64]
    lib.cc_emul_widen_c32.argtypes = [P, I64, I64, P, I64, I64]
    return lib


def crandn(rng, shape):
    return np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


def strides(a):
    return a.strides[0] // a.itemsize, a.strides[1] // a.itemsize


def run_tridiag(lib, A, reverse=0):
    n = A.shape[0]
    W = np.zeros((n, n), dtype=np.complex128, order="F")
    tau = np.zeros(max(n - 1, 1)); d = np.zeros(n); e = np.zeros(max(n - 1, 1))
    ph = np.zeros(n, dtype=np.complex128); tauc = np.zeros(max(n - 1, 1), dtype=np.complex128)
    rs, cs = strides(A)
    lib.cc_emul_tridiag(A.ctypes.data, rs, cs, n, W.ctypes.data, tau.ctypes.data, d.ctypes.data, e.ctypes.data, ph.ctypes.data,
                        tauc.ctypes.data, reverse)
    return W, tau[:n - 1], d, e[:n - 1], ph, tauc[:n - 1]


def run_bidiag(lib, A, adjoint=0, reverse=0):
    m, n = (A.shape[1], A.shape[0]) if adjoint else A.shape
    W = np.zeros((m, n), dtype=np.complex128, order="F")
    tl = np.zeros(n); tr = np.zeros(max(n - 1, 1)); d = np.zeros(n); f = np.zeros(max(n - 1, 1))
    l = np.zeros(n, dtype=np.complex128); r = np.zeros(n, dtype=np.complex128)
    tlc = np.zeros(n, dtype=np.complex128); trc = np.zeros(max(n - 1, 1), dtype=np.complex128)
    rs, cs = strides(A)
    lib.cc_emul_bidiag(A.ctypes.data, rs, cs, m, n, adjoint, W.ctypes.data, tl.ctypes.data, tr.ctypes.data, d.ctypes.data, f.ctypes.data,
                       l.ctypes.data, r.ctypes.data, tlc.ctypes.data, trc.ctypes.data, reverse)
    return W, tl, tr[:n - 1], d, f[:n - 1], l, r, tlc, trc[:n - 1]


@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 130, 300])
def test_tridiag_emulation_matches_the_model(emul, n):
    rng = np.random.default_rng(900 + n)
    G = crandn(rng, (n, n))
    A = np.asfortranarray(G + G.conj().T)
    A[np.triu_indices(n, 1)] = 77.0 - 3.0j          # the strict upper triangle is never read
    A[np.diag_indices(n)] += 5.0j                    # nor the imaginary part of the diagonal
    Wm, taum = cm.tridiag_unblocked(A)
    W, tau, d, e, ph, tauc = run_tridiag(emul, A)
    Wr = run_tridiag(emul, A, reverse=1)
    for a, b in zip((W, tau, d, e, ph, tauc), Wr):
        assert np.array_equal(a, b)                  # thread order within a launch does not matter
    scale = np.abs(Wm).max() * n
    low = np.tril_indices(n)
    assert np.abs(np.tril(W, -2) - np.tril(Wm, -2)).max(initial=0.0) <= 4096 * n * U       # essentials
    assert np.abs(np.diag(W) - np.diag(Wm)).max() <= 4096 * U * scale
    assert np.abs(np.diag(W, -1) - np.diag(Wm, -1)).max(initial=0.0) <= 4096 * U * scale
    assert np.all(np.isfinite(tau) == np.isfinite(taum)) and np.allclose(tau[np.isfinite(tau)], taum[np.isfinite(taum)], rtol=4096 * n * U)
    assert np.array_equal(d, np.diag(W).real) and np.allclose(e, np.abs(np.diag(W, -1)), rtol=4 * U, atol=0)
    assert np.array_equal(tauc, tau.astype(np.complex128))
    # the phases make the tridiagonal real: conj(ph[k+1]) T[k+1, k] ph[k] = |T[k+1, k]|
    assert np.abs(np.abs(ph) - 1).max() <= 8 * U
    if n > 1:
        sub = np.diag(W, -1)
        assert np.abs(ph[1:].conj() * sub * ph[:-1] - np.abs(sub)).max() <= 64 * U * max(1.0, np.abs(sub).max())


def test_tridiag_emulation_strided_and_zero_columns(emul):
    rng = np.random.default_rng(950)
    n = 40
    G = crandn(rng, (n, n))
    H = G + G.conj().T
    H[20:, :20] = 0; H[:20, 20:] = 0                 # block diagonal: column 19 has a zero tail (tau = inf, no update, phase 1)
    big = np.zeros((2 * n, 3 * n), dtype=np.complex128)
    view = big[::2, ::3][:n, :n]                     # row-major storage, strides (6n, 3) in elements
    view[...] = H
    W1 = run_tridiag(emul, np.asfortranarray(H))
    W2 = run_tridiag(emul, view)
    for a, b in zip(W1, W2):
        assert np.array_equal(a, b)
    W, tau, d, e, ph, _ = W1
    assert np.isinf(tau[19]) and e[19] == 0.0 and abs(ph[20] - ph[19]) <= 4 * U
    # end to end with LAPACK as the real tridiagonal solver
    T = np.diag(d) + np.diag(e, -1) + np.diag(e, 1)
    lam, Q = np.linalg.eigh(T)
    Uv = ph[:, None] * Q
    cm.apply_sequence(W[1:, :n - 1], tau, Uv[1:, :])
    assert np.abs(Uv @ np.diag(lam) @ Uv.conj().T - H).max() <= 256 * n * U * np.abs(H).max()
    assert np.abs(Uv.conj().T @ Uv - np.eye(n)).max() <= 256 * n * U


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (5, 3), (17, 17), (64, 20), (130, 130), (300, 70)])
def test_bidiag_emulation_matches_the_model(emul, shape):
    m, n = shape
    rng = np.random.default_rng(1000 + m + n)
    A = crandn(rng, (m, n))
    Wm, tlm, trm = cm.bidiag_unblocked(A)
    out = run_bidiag(emul, A)
    outr = run_bidiag(emul, A, reverse=1)
    for a, b in zip(out, outr):
        assert np.array_equal(a, b)
    W, tl, tr, d, f, l, r, tlc, trc = out
    scale = np.abs(A).max() * max(m, n)
    assert np.abs(np.diag(W) - np.diag(Wm)).max() <= 4096 * U * scale
    assert np.abs(np.diag(W[:n, :n], 1) - np.diag(Wm[:n, :n], 1)).max(initial=0.0) <= 4096 * U * scale
    assert np.abs(np.tril(W, -1) - np.tril(Wm, -1)).max(initial=0.0) <= 4096 * max(m, n) * U
    assert np.abs(np.triu(W[:n, :n], 2) - np.triu(Wm[:n, :n], 2)).max(initial=0.0) <= 4096 * max(m, n) * U
    assert np.allclose(tl, tlm, rtol=4096 * max(m, n) * U) and np.allclose(tr, trm, rtol=4096 * max(m, n) * U)
    assert np.array_equal(tlc, tl.astype(np.complex128)) and np.array_equal(trc, tr.astype(np.complex128))
    # the phases make the bidiagonal real and nonnegative
    B = np.diag(np.diag(W)) + np.diag(np.diag(W[:n, :n], 1), 1)
    Breal = np.diag(l.conj()) @ B @ np.diag(r)
    assert np.abs(Breal - (np.diag(d) + np.diag(f, 1))).max() <= 64 * U * max(1.0, np.abs(B).max())
    # end to end with LAPACK as the real bidiagonal solver (thin vectors)
    Ub, S, Vbt = np.linalg.svd(np.diag(d) + np.diag(f, 1))
    Uv = np.zeros((m, n), dtype=np.complex128); Uv[:n, :] = l[:, None] * Ub
    cm.apply_sequence(W, tl, Uv)
    Vv = (r[:, None] * Vbt.T).astype(np.complex128)
    if n > 1:
        cm.apply_sequence(W[:n, :].T[1:, :n - 1], tr, Vv[1:, :], conj=True)
    tol = U * 128 * np.sqrt(8 * max(m, n)) * max(1.0, np.abs(A).max())
    assert np.abs(Uv @ np.diag(S) @ Vv.conj().T - A).max() <= tol
    assert np.abs(Uv.conj().T @ Uv - np.eye(n)).max() <= tol and np.abs(Vv.conj().T @ Vv - np.eye(n)).max() <= tol


def test_bidiag_emulation_adjoint_strided_and_rank_deficient(emul):
    rng = np.random.default_rng(1100)
    A = crandn(rng, (12, 30))                        # wide: the driver works on A^H
    ref = run_bidiag(emul, np.asfortranarray(A.conj().T))
    got = run_bidiag(emul, A, adjoint=1)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    got = run_bidiag(emul, np.ascontiguousarray(A), adjoint=1)   # row-major input
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    Z = crandn(rng, (20, 8)); Z[:, 3] = 0; Z[:, 6] = Z[:, 1]     # a zero column and a repeated column
    W, tl, tr, d, f, l, r, _, _ = run_bidiag(emul, Z)
    Ub, S, Vbt = np.linalg.svd(np.diag(d) + np.diag(f, 1))
    Uv = np.zeros((20, 8), dtype=np.complex128); Uv[:8, :] = l[:, None] * Ub
    cm.apply_sequence(W, tl, Uv)
    Vv = (r[:, None] * Vbt.T).astype(np.complex128)
    cm.apply_sequence(W[:8, :].T[1:, :7], tr, Vv[1:, :], conj=True)
    assert np.abs(Uv @ np.diag(S) @ Vv.conj().T - Z).max() <= 1e-13 * np.abs(Z).max() * 20
    assert S[-1] <= 1e-13 and S[-2] <= 1e-13


def test_assembly_bodies(emul):
    rng = np.random.default_rng(1200)
    nq, rows, cols = 5, 9, 7
    Q = np.asfortranarray(rng.standard_normal((nq, nq)))
    ph = np.exp(1j * rng.standard_normal(nq))
    for rev in (0, 1):
        out = np.full((rows, cols), np.nan, dtype=np.complex128, order="F")
        emul.cc_emul_scale_rows_embed(Q.ctypes.data, nq, nq, ph.ctypes.data, out.ctypes.data, rows, cols, rev)
        want = np.zeros((rows, cols), dtype=np.complex128); want[:nq, :nq] = ph[:, None] * Q
        for j in range(nq, cols):
            want[j, j] = 1.0
        assert np.array_equal(out, want)
        W = crandn(rng, (8, 6)); T = np.zeros((6, 6), dtype=np.complex128, order="F")
        emul.cc_emul_transpose_corner(W.ctypes.data, 8, T.ctypes.data, 6, rev)
        assert np.array_equal(T, W[:6, :6].T)
        src = crandn(rng, (rows, cols))
        big = np.full((2 * rows, 2 * cols), np.nan, dtype=np.complex128)      # row-major, every other row / column
        emul.cc_emul_copy_out_f64(big.ctypes.data, 4 * cols, 2, src.ctypes.data, rows, cols, rev)
        assert np.array_equal(big[::2, ::2], src) and np.all(np.isnan(big[1::2, :])) and np.all(np.isnan(big[:, 1::2]))
        o32 = np.zeros((rows, cols), dtype=np.complex64, order="F")
        emul.cc_emul_copy_out_f32(o32.ctypes.data, 1, rows, src.ctypes.data, rows, cols, rev)
        assert np.array_equal(o32, src.astype(np.complex64))
    vals = rng.standard_normal(6); S = np.full(12, np.nan, dtype=np.complex128)
    emul.cc_emul_copy_values_f64(S.ctypes.data, 2, vals.ctypes.data, 6)
    assert np.array_equal(S[::2], vals.astype(np.complex128)) and np.all(np.isnan(S[1::2]))
    a32 = (rng.standard_normal((4, 3)) + 1j * rng.standard_normal((4, 3))).astype(np.complex64)     # row-major c32
    w = np.zeros((4, 3), dtype=np.complex128, order="F")
    emul.cc_emul_widen_c32(a32.ctypes.data, 3, 1, w.ctypes.data, 4, 3)
    assert np.array_equal(w, a32.astype(np.complex128))


def _svd_from_emulation(emul, A):
    m, n = A.shape
    W, tl, tr, d, f, l, r, _, _ = run_bidiag(emul, np.asfortranarray(A))
    Ub, S, Vbt = np.linalg.svd(np.diag(d) + np.diag(f, 1))
    Uv = np.zeros((m, n), dtype=np.complex128); Uv[:n, :] = l[:, None] * Ub
    cm.apply_sequence(W, tl, Uv)
    Vv = (r[:, None] * Vbt.T).astype(np.complex128)
    if n > 1:
        cm.apply_sequence(W[:n, :].T[1:, :n - 1], tr, Vv[1:, :], conj=True)
    return S, Uv, Vv, (W, tl, tr, d, f)


@pytest.mark.parametrize("shape", [(6, 6), (12, 7), (40, 40), (64, 10), (150, 150)])
def test_special_matrices_through_the_emulation(emul, shape):
    """svd/mod.rs:773-982 specials. The all-ones matrix is the hard one: after the first column the trailing matrix is rounding
    noise that shrinks like eps^k per step, through the subnormal range down to exact zeros — the three-accumulator norm and
    the min_positive tests of make_householder (householder.rs:59-107) keep every reflector unitary on the way."""
    m, n = shape
    for A in (np.zeros((m, n)), np.ones((m, n)), np.eye(m, n), 1e-170 * np.ones((m, n)), 1e160 * np.eye(m, n)):
        A = A.astype(np.complex128)
        S, Uv, Vv, (W, tl, tr, d, f) = _svd_from_emulation(emul, A)
        assert np.all(np.isfinite(W)) and np.all(np.isfinite(d)) and np.all(np.isfinite(f))
        scale = max(np.abs(A).max(), np.finfo(float).tiny)
        tol = U * 128 * np.sqrt(8 * max(m, n))
        assert np.abs(Uv.conj().T @ Uv - np.eye(n)).max() <= tol and np.abs(Vv.conj().T @ Vv - np.eye(n)).max() <= tol
        assert np.abs((Uv * (S / scale)[None, :]) @ Vv.conj().T - A / scale).max() <= tol


def test_non_finite_input_reaches_the_condensed_form(emul):
    """The entry points report NoConvergence from a finiteness test of the real condensed entries (svd/mod.rs:282-286): a NaN or
    Inf anywhere in the input must reach them."""
    rng = np.random.default_rng(1300)
    for bad in (np.nan, np.inf):
        A = crandn(rng, (30, 20)); A[7, 3] = bad
        _, _, _, d, f, _, _, _, _ = run_bidiag(emul, A)
        assert not (np.all(np.isfinite(d)) and np.all(np.isfinite(f)))
        G = crandn(rng, (25, 25)); H = np.asfortranarray(G + G.conj().T); H[9, 2] = bad
        _, _, d, e, _, _ = run_tridiag(emul, H)
        assert not (np.all(np.isfinite(d)) and np.all(np.isfinite(e)))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 8, 16, 45, 130])
def test_hessenberg_emulation(emul, oracle, n):
    """evd/hessenberg.rs tests (test_hessenberg_cplx: n in {1..16}, block size 3): the emulated launch sequence against the model,
    and the reference's own check — Q^H A Q through the block-Householder sequences with the T blocks equals the Hessenberg part."""
    emul.cc_emul_hessenberg.argtypes = [P, I64, I64, I64, P, P, P, I64, C.c_int]
    rng = np.random.default_rng(1800 + n)
    A = crandn(rng, (n, n))
    bs = 3
    outs = []
    for rev in (0, 1):
        W = np.zeros((n, n), dtype=np.complex128, order="F"); tau = np.zeros(max(n - 1, 1)); Tf = np.zeros((bs, max(n - 1, 1)), dtype=np.complex128, order="F")
        rs, cs = strides(A)
        emul.cc_emul_hessenberg(A.ctypes.data, rs, cs, n, W.ctypes.data, tau.ctypes.data, Tf.ctypes.data, bs, rev)
        outs.append((W, tau[:n - 1], Tf[:, :n - 1]))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    W, tau, Tf = outs[0]
    Wm, taum = cm.hessenberg_unblocked(A)
    scale = np.abs(A).max() * n
    assert np.abs(np.triu(W, -1) - np.triu(Wm, -1)).max() <= 4096 * U * scale
    assert np.abs(np.tril(W, -2) - np.tril(Wm, -2)).max(initial=0.0) <= 4096 * n * U
    if n > 1:
        assert np.allclose(tau[np.isfinite(tau)], taum[np.isfinite(taum)], rtol=4096 * n * U) and np.array_equal(np.isfinite(tau), np.isfinite(taum))
        V = np.asfortranarray(W[1:, :n - 1])
        assert np.allclose(Tf, cm.t_blocks(V, tau, bs), rtol=1e-12, atol=1e-12 * n)
        B = A.copy(order="F")
        oracle.apply_q_transpose_sequence(V, np.asfortranarray(Tf), B[1:, :], conj_lhs=True)          # Q^H A
        oracle.apply_q_transpose_sequence(V, np.asfortranarray(Tf), B.T[1:, :], conj_lhs=False)       # (Q^H A) Q
        assert np.abs(B - np.triu(W, -1)).max() <= 256 * n * U * np.abs(A).max()
