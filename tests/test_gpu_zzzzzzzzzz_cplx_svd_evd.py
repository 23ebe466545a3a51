"""`svd` and `self_adjoint_evd` for complex T through the C ABI (`libfaer_v0_23_svd_{c64,c32}`,
`libfaer_v0_23_self_adjoint_evd_{c64,c32}`; csrc/cplx_condensed.cu), restating the reference's own tests, which draw complex
matrices too:

  svd/mod.rs:773-982   shapes up to 150^2 incl. wide / tall, zeros / ones / identity specials; thin, full and one-sided variants
                       agree; tolerance eps * 128 * sqrt(8 max(m, n)) (780-783) on U S V^H ~ A, plus unitarity of U and V
  evd/mod.rs (tests)   self-adjoint: U S U^H ~ A, U unitary, nondecreasing S; only the lower triangle is read
and the non-finite-input contract (SvdError / EvdError::NoConvergence). The checker is LAPACK (numpy) on the same matrices; the
launch sequences underneath are the ones test_cplx_condensed_emul_cpu.py runs thread by thread on the CPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CDTYPES = [np.complex128, np.complex64]


def crandn(rng, shape, dtype):
    return np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype))


def approx(a, b, tol):
    d = np.abs(a - b)
    return bool(np.all((d <= tol) | (d <= tol * np.maximum(np.abs(a), np.abs(b)))))


def check_svd(la, A, dtype):
    m, n = A.shape
    size = min(m, n)
    rdtype = np.float64 if dtype == np.complex128 else np.float32
    eps = np.finfo(rdtype).eps
    tol = eps * 128 * np.sqrt(8 * max(m, n, 1))
    scale = max(1.0, float(np.abs(A).max()) if A.size else 1.0)
    ref = np.linalg.svd(A.astype(np.complex128), compute_uv=False) if size else np.zeros(0)
    outs = {}
    for kind in ("full", "thin", "u_only", "v_only"):
        S = np.zeros(size, dtype=dtype)                 # T-typed, as the ABI has it: (value, 0)
        U = np.zeros((m, m if kind == "full" else size), dtype=dtype, order="F") if kind != "v_only" else None
        V = np.zeros((n, n if kind == "full" else size), dtype=dtype, order="F") if kind != "u_only" else None
        la.svd(A, S, U, V)
        assert np.all(S.imag == 0), (m, n, kind)
        Sr = S.real
        outs[kind] = Sr
        assert np.all(np.diff(Sr) <= 0) and np.all(Sr >= 0), (m, n, kind)
        assert np.abs(Sr - ref).max(initial=0) <= tol * scale * max(1.0, ref.max(initial=0) / scale), (m, n, kind)
        if U is not None:
            assert np.abs(U.conj().T @ U - np.eye(U.shape[1])).max(initial=0) <= tol, (m, n, kind, "U unitarity")
        if V is not None:
            assert np.abs(V.conj().T @ V - np.eye(V.shape[1])).max(initial=0) <= tol, (m, n, kind, "V unitarity")
        if U is not None and V is not None:
            rec = (U[:, :size] * Sr[None, :]) @ V[:, :size].conj().T
            assert approx(rec, A, tol * scale), (m, n, kind, float(np.abs(rec - A).max(initial=0)))
    vals = la.singular_values(A)
    assert not np.iscomplexobj(vals)
    for k, Sr in outs.items():
        assert np.abs(vals - Sr).max(initial=0) <= tol * scale * max(1.0, ref.max(initial=0) / scale), (m, n, k)
    # a REAL S is accepted by the Python mirror too
    Sreal = np.zeros(size, dtype=rdtype)
    la.svd(A, Sreal)
    assert np.abs(Sreal - outs["thin"]).max(initial=0) <= tol * scale * max(1.0, ref.max(initial=0) / scale)


@pytest.mark.parametrize("dtype", CDTYPES)
def test_cplx_svd_reference_shapes(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(320)
    for (m, n) in [(3, 2), (2, 2), (4, 4), (15, 10), (10, 10), (15, 15), (50, 50), (100, 100), (150, 150), (150, 20), (20, 150),
                   (110, 60), (60, 110), (1, 1), (1, 7), (7, 1), (33, 32), (257, 130)]:
        check_svd(la, crandn(rng, (m, n), dtype), dtype)
    for (m, n) in [(6, 6), (12, 7), (7, 12), (40, 40), (64, 10)]:
        check_svd(la, np.zeros((m, n), dtype=dtype, order="F"), dtype)
        check_svd(la, np.ones((m, n), dtype=dtype, order="F"), dtype)
        check_svd(la, np.asfortranarray(np.eye(m, n, dtype=dtype)), dtype)
    # rank deficient, row-major and strided inputs
    A = (crandn(rng, (90, 7), np.complex128) @ crandn(rng, (7, 70), np.complex128)).astype(dtype)
    check_svd(la, np.asfortranarray(A), dtype)
    check_svd(la, np.ascontiguousarray(A), dtype)
    big = np.zeros((180, 140), dtype=dtype)
    big[::2, ::2] = A
    check_svd(la, big[::2, ::2], dtype)


@pytest.mark.parametrize("dtype", CDTYPES)
def test_cplx_self_adjoint_evd(fb, cuda_dev, dtype):
    la = fb.linalg
    rng = np.random.default_rng(321)
    rdtype = np.float64 if dtype == np.complex128 else np.float32
    eps = np.finfo(rdtype).eps
    for n in [1, 2, 3, 4, 8, 16, 31, 33, 64, 100, 150, 257]:
        G = crandn(rng, (n, n), np.complex128)
        A = np.asfortranarray(((G + G.conj().T) / 2).astype(dtype))
        poisoned = A.copy(order="F"); poisoned[np.triu_indices(n, 1)] = np.nan  # only the lower triangle may be read
        S = np.zeros(n, dtype=dtype); U = np.zeros((n, n), dtype=dtype, order="F")
        la.self_adjoint_evd(poisoned, S, U)
        assert np.all(S.imag == 0)
        Sr = S.real
        tol = eps * 128 * np.sqrt(8 * n) * max(1.0, float(np.abs(A).max()))
        assert np.all(np.diff(Sr) >= 0)
        assert np.abs(U.conj().T @ U - np.eye(n)).max() <= tol, n
        rec = (U * Sr[None, :]) @ U.conj().T
        assert approx(rec, A, tol), (n, float(np.abs(rec - A).max()))
        ref = np.linalg.eigvalsh(A.astype(np.complex128))
        assert np.abs(Sr - ref).max() <= tol * max(1.0, np.abs(ref).max()), n
        vals = la.self_adjoint_eigenvalues(poisoned)
        assert not np.iscomplexobj(vals) and np.abs(vals - Sr).max() <= tol * max(1.0, np.abs(ref).max())
    # clustered / repeated eigenvalues, zero matrix, identity, a block-diagonal matrix (zero subdiagonal entry in the middle)
    G = crandn(rng, (20, 20), np.complex128); H = G + G.conj().T
    blk = np.zeros((40, 40), dtype=np.complex128); blk[:20, :20] = H; blk[20:, 20:] = H.conj()
    for A in [np.zeros((40, 40)), np.eye(40), np.diag(np.repeat([1.0, 2.0, 3.0, 4.0], 10)), np.ones((50, 50)), blk]:
        A = np.asfortranarray(A.astype(dtype)); n = A.shape[0]
        S = np.zeros(n, dtype=dtype); U = np.zeros((n, n), dtype=dtype, order="F")
        la.self_adjoint_evd(A, S, U)
        tol = eps * 128 * np.sqrt(8 * n) * max(1.0, float(np.abs(A).max()))
        assert np.abs(U.conj().T @ U - np.eye(n)).max() <= tol and approx((U * S.real[None, :]) @ U.conj().T, A, tol)


def test_cplx_non_finite_input_is_no_convergence(fb, cuda_dev):
    la = fb.linalg
    rng = np.random.default_rng(322)
    A = crandn(rng, (30, 20), np.complex128)
    A[7, 3] = np.nan
    S = np.zeros(20, dtype=np.complex128); U = np.zeros((30, 20), dtype=np.complex128, order="F"); V = np.zeros((20, 20), dtype=np.complex128, order="F")
    with pytest.raises(RuntimeError, match="NoConvergence"):
        la.svd(A, S, U, V)
    with pytest.raises(RuntimeError, match="NoConvergence"):
        la.singular_values(A)
    G = crandn(rng, (25, 25), np.complex128); H = np.asfortranarray(G + G.conj().T)
    H[9, 2] = np.inf
    S = np.zeros(25, dtype=np.complex128); U = np.zeros((25, 25), dtype=np.complex128, order="F")
    with pytest.raises(RuntimeError, match="NoConvergence"):
        la.self_adjoint_evd(H, S, U)


def test_cplx_svd_evd_device_resident(fb, cuda_dev):
    import torch
    la = fb.linalg
    torch.manual_seed(5)
    m, n = 300, 120
    A = torch.randn((n, m), dtype=torch.complex128, device=cuda_dev).T          # column-major m x n on the device
    S = torch.zeros(n, dtype=torch.complex128, device=cuda_dev)
    U = torch.zeros((n, m), dtype=torch.complex128, device=cuda_dev).T
    V = torch.zeros((n, n), dtype=torch.complex128, device=cuda_dev).T
    la.svd(A, S, U, V)
    tol = np.finfo(float).eps * 128 * np.sqrt(8 * m)
    eye = torch.eye(n, dtype=torch.complex128, device=cuda_dev)
    assert float((U.conj().T @ U - eye).abs().max()) <= tol and float((V.conj().T @ V - eye).abs().max()) <= tol
    assert float(((U * S[None, :]) @ V.conj().T - A).abs().max()) <= tol * max(1.0, float(A.abs().max()))
    assert float((S.real - torch.linalg.svdvals(A)).abs().max()) <= tol * float(S.real.max())
    H = torch.randn((n, n), dtype=torch.complex128, device=cuda_dev)
    H = (H + H.conj().T).T.contiguous().T
    E = torch.zeros(n, dtype=torch.complex128, device=cuda_dev)
    Q = torch.zeros((n, n), dtype=torch.complex128, device=cuda_dev).T
    la.self_adjoint_evd(H, E, Q)
    tol = np.finfo(float).eps * 128 * np.sqrt(8 * n) * max(1.0, float(H.abs().max()))
    assert float((Q.conj().T @ Q - eye).abs().max()) <= tol
    assert float(((Q * E[None, :]) @ Q.conj().T - H).abs().max()) <= tol


def test_cplx_solvers_svd_and_eigen(fb, cuda_dev):
    """`Svd::new` / `new_thin`, `SelfAdjointEigen::new` (both sides) and the pseudo-inverse on complex matrices
    (solvers.rs:1324-1520; test_pinv svd/mod.rs:1055-...)."""
    sv = fb.solvers
    rng = np.random.default_rng(323)
    A = crandn(rng, (6, 36), np.complex128)
    d = sv.Svd.new(A)
    assert d.U().shape == (6, 6) and d.V().shape == (36, 36) and d.S().shape == (6,) and not np.iscomplexobj(d.S())
    t = sv.Svd.new_thin(A)
    assert t.U().shape == (6, 6) and t.V().shape == (36, 6)
    assert np.abs((t.U() * t.S()[None, :]) @ t.V().conj().T - A).max() <= 1e-13
    pinv = t.pseudoinverse()
    assert np.abs(pinv - np.linalg.pinv(A)).max() <= 1e-12
    assert np.abs(A @ pinv @ A - A).max() <= 1e-12
    G = crandn(rng, (30, 30), np.complex128); H = np.asfortranarray(G + G.conj().T)
    for side, tri in ((sv.Side.Lower, np.tril), (sv.Side.Upper, np.triu)):
        P = np.asfortranarray(tri(H))                      # only the chosen triangle is given
        e = sv.SelfAdjointEigen.new(P, side)
        assert not np.iscomplexobj(e.S())
        assert np.abs((e.U() * e.S()[None, :]) @ e.U().conj().T - H).max() <= 1e-12
