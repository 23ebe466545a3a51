"""Shared checks for faer_b200.solvers, restating the reference's `test_all_solvers`
(faer/src/linalg/solvers.rs:2919-2977: n = 50, three right-hand sides, eight solve / rsolve identities per decomposition,
tolerance eps * 128 * n) for f64 and — as the reference itself runs it — c64 (`cplx=True`), plus the accessors' contracts (split_LU, thin_R, compute_Q, P,
reconstruct, inverse, least squares, Side::Upper, LltError).

Run twice: on the GPU through the C ABI (tests/test_gpu_zz3_solvers.py) and on the CPU with `solvers.la` swapped for an
oracle-backed stand-in (tests/test_solvers_host_logic_cpu.py), which checks the host-side logic of solvers.py itself
(buffer ownership, factor splitting, call order) without a GPU. The stand-in lives in the tests: the product never
routes through the oracle.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def approx(a, b, n, scale=1.0):
    tol = EPS * 128.0 * n * scale
    return bool(np.all(np.abs(np.asarray(a) - np.asarray(b)) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))))


def randn(rng, shape, cplx):
    x = rng.standard_normal(shape)
    if cplx:
        x = x + 1j * rng.standard_normal(shape)
    return np.asfortranarray(x)


def check_solver(A, dec, cond):
    """test_solver_imp (solvers.rs:2921-2946): the eight identities, with conjugate / adjoint meaning what they say for c64."""
    rng = np.random.default_rng(0xC0FFEE)
    n, k = A.shape[0], 3
    cplx = np.iscomplexobj(A)
    R = randn(rng, (n, k), cplx)
    L = randn(rng, (k, n), cplx)
    s = max(1.0, cond) * max(1.0, np.abs(A).max())
    assert approx(A @ dec.solve(R), R, n, s)
    assert approx(A.conj() @ dec.solve_conjugate(R), R, n, s)
    assert approx(A.T @ dec.solve_transpose(R), R, n, s)
    assert approx(A.conj().T @ dec.solve_adjoint(R), R, n, s)
    assert approx(dec.rsolve(L) @ A, L, n, s)
    assert approx(dec.rsolve_conjugate(L) @ A.conj(), L, n, s)
    assert approx(dec.rsolve_transpose(L) @ A.T, L, n, s)
    assert approx(dec.rsolve_adjoint(L) @ A.conj().T, L, n, s)
    # in-place forms and vectors
    X = R.copy(order="F"); dec.solve_in_place(X)
    assert np.allclose(X, dec.solve(R), rtol=1e-12, atol=1e-14)  # same calls: the two forms agree
    v = R[:, 0].copy(); dec.solve_in_place(v)
    assert approx(A @ v, R[:, 0], n, s)
    Y = L.copy(order="F"); dec.rsolve_in_place(Y)
    assert approx(Y @ A, L, n, s)
    assert approx(dec.inverse() @ A, np.eye(n), n, s)
    assert approx(dec.reconstruct(), A, n, s)


def run_all(sv, cplx=False):
    rng = np.random.default_rng(0)
    n = 50
    A = randn(rng, (n, n), cplx)
    cond = np.linalg.cond(A)

    # ---- PartialPivLu (solvers.rs:981-1034) ----
    lu = sv.partial_piv_lu(A)
    assert (lu.nrows(), lu.ncols()) == (n, n)
    Lf, Uf = lu.L(), lu.U()
    assert np.all(np.triu(Lf, 1) == 0) and np.all(np.diag(Lf) == 1) and np.all(np.tril(Uf, -1) == 0)
    # partial pivoting: |l| <= 1 for reals; the complex pivot rule is abs1 = |re| + |im|, hence |l| <= sqrt(2)
    assert np.all(np.abs(np.tril(Lf, -1)) <= (np.sqrt(2.0) * (1 + 1e-14) if cplx else 1.0))
    fwd, bwd = lu.P()
    assert sorted(int(i) for i in fwd) == list(range(n)) and all(int(bwd[int(fwd[i])]) == i for i in range(n))
    assert approx(Lf @ Uf, A[np.asarray(fwd, dtype=np.int64)], n, np.abs(A).max() * n)
    check_solver(A, lu, cond)
    # rectangular factorizations: split_LU's two branches and reconstruct
    for (m2, n2) in [(70, 30), (30, 70)]:
        B = randn(rng, (m2, n2), cplx)
        d = sv.PartialPivLu.new(B)
        size = min(m2, n2)
        assert d.L().shape == ((m2, n2) if m2 >= n2 else (size, size))
        assert d.U().shape == ((size, size) if m2 >= n2 else (m2, n2))
        assert np.all(np.diag(d.L()) == 1) and np.all(np.triu(d.L(), 1) == 0) and np.all(np.tril(d.U(), -1) == 0)
        assert approx(d.reconstruct(), B, max(m2, n2), np.abs(B).max() * size)
    assert np.array_equal(A, np.asfortranarray(A))  # inputs are never modified

    # ---- Qr (solvers.rs:1106-1205) ----
    qr = sv.qr(A)
    assert qr.Q_coeff().shape[1] == n and qr.R().shape == (n, n) and np.all(np.tril(qr.R(), -1) == 0)
    assert np.all(np.diag(qr.Q_basis()) == 1) and np.all(np.triu(qr.Q_basis(), 1) == 0)
    Q = qr.compute_Q()
    assert approx(Q.conj().T @ Q, np.eye(n), n) and approx(Q @ qr.R(), A, n, np.abs(A).max() * n)
    check_solver(A, qr, cond)
    m2, n2 = 120, 40
    B = randn(rng, (m2, n2), cplx)
    rhs = randn(rng, (m2, 4), cplx)
    d = sv.Qr.new(B)
    assert d.Q_basis().shape == (m2, n2) and d.R().shape == (n2, n2) and d.thin_R().shape == (n2, n2)
    tq = d.compute_thin_Q()
    assert tq.shape == (m2, n2) and approx(tq.conj().T @ tq, np.eye(n2), m2) and approx(tq @ d.thin_R(), B, m2, np.abs(B).max() * n2)
    assert approx(d.compute_Q()[:, :n2], tq, m2)
    x = d.solve_lstsq(rhs)
    assert x.shape == (n2, 4)
    assert approx(x, np.linalg.lstsq(B, rhs, rcond=None)[0], m2, np.linalg.cond(B))
    xc = d.solve_conjugate_lstsq(rhs)  # least squares with conj(B)
    assert approx(xc, np.linalg.lstsq(B.conj(), rhs, rcond=None)[0], m2, np.linalg.cond(B))
    if not cplx:
        assert np.allclose(xc, x, rtol=1e-12, atol=1e-14)
    assert approx(d.reconstruct(), B, m2, np.abs(B).max() * n2)
    wide = randn(rng, (30, 70), cplx)
    d = sv.Qr.new(wide)
    assert d.Q_basis().shape == (30, 30) and d.R().shape == (30, 70) and d.thin_R().shape == (30, 70)
    assert approx(d.compute_Q() @ d.R(), wide, 70, np.abs(wide).max() * 30)
    assert approx(d.reconstruct(), wide, 70, np.abs(wide).max() * 30)

    # ---- Llt (solvers.rs:770-816) ----
    S = np.asfortranarray(A @ A.conj().T)
    llt = sv.llt(S, sv.Side.Lower)
    assert np.all(np.triu(llt.L(), 1) == 0) and np.all(np.diag(llt.L()).real > 0)
    # (the factorization reads Re(a_jj) and scales the stored element: a rounding-level imaginary part of S's diagonal stays one)
    assert np.all(np.abs(np.diag(llt.L()).imag) <= 1e-14 * np.diag(llt.L()).real)
    assert approx(llt.L() @ llt.L().conj().T, S, n, np.abs(S).max())
    check_solver(S, llt, np.linalg.cond(S))
    # only the chosen triangle is read
    poisoned = S.copy(order="F"); poisoned[np.triu_indices(n, 1)] = np.nan
    assert np.allclose(sv.Llt.new(poisoned, sv.Side.Lower).L(), llt.L(), rtol=1e-12, atol=1e-14)
    poisoned = S.copy(order="F"); poisoned[np.tril_indices(n, -1)] = np.nan
    assert np.allclose(sv.Llt.new(poisoned, sv.Side.Upper).L(), llt.L(), rtol=1e-12, atol=1e-14)
    # NonPositivePivot { index } (llt/factor.rs:21-24)
    bad = S.copy(order="F"); bad[7, 7] = -1.0
    try:
        sv.llt(bad)
        raise AssertionError("expected LltError")
    except sv.LltError as e:
        assert e.index == 7

    # ---- `&A * &B` ----
    C = sv.mul(A, B[:n, :])
    assert C.shape == (n, n2) and approx(C, A @ B[:n, :], n, np.abs(A).max() * np.abs(B).max() * n)


def run_ldlt(sv):
    """Ldlt (solvers.rs:818-872) on a symmetric indefinite matrix with a dominant diagonal: test_solver identities, L / D
    contracts, both sides, ZeroPivot. Separate from run_all (its own kernel: csrc/ldlt_f64.cu)."""
    rng = np.random.default_rng(7)
    n = 50
    G = rng.standard_normal((n, n))
    s = np.where(rng.random(n) < 0.4, -1.0, 1.0)
    A = np.asfortranarray((G + G.T) / np.sqrt(n) + np.diag(4.0 * s))
    dec = sv.ldlt(A, sv.Side.Lower)
    L, D = dec.L(), dec.D()
    assert np.all(np.diag(L) == 1) and np.all(np.triu(L, 1) == 0) and D.shape == (n,)
    assert approx(L @ np.diag(D) @ L.T, A, n, np.abs(A).max())
    assert (np.asarray(D) < 0).sum() == (np.linalg.eigvalsh(A) < 0).sum()
    check_solver(A, dec, np.linalg.cond(A))
    poisoned = A.copy(order="F"); poisoned[np.tril_indices(n, -1)] = np.nan
    up = sv.Ldlt.new(poisoned, sv.Side.Upper)
    assert np.allclose(up.L(), L, rtol=1e-12, atol=1e-14) and np.allclose(up.D(), D, rtol=1e-12, atol=1e-14)
    bad = A.copy(order="F")
    bad[:4, :4] = np.diag([2.0, 4.0, 8.0, 0.0]); bad[3, :3] = bad[:3, 3] = [2.0, 4.0, 8.0]; bad[3, 3] = 14.0
    try:
        sv.ldlt(bad)
        raise AssertionError("expected LdltError")
    except sv.LdltError as e:
        assert e.index == 3
