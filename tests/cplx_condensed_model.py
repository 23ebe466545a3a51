"""Host model (numpy) of the complex condensed-form drivers of csrc/cplx_condensed.cu, statement by statement: the unblocked
Householder tridiagonalization / bidiagonalization with the reference's reflector definition, the phase normalisation that makes
the condensed matrix real, and the back-transforms. The CUDA file is a transcription of these functions (one kernel per
numbered step); tests/test_cplx_condensed_model_cpu.py checks the model against the oracle's restatement of the reference's
fused algorithms (same reflectors, same condensed entries up to rounding) and the end-to-end identities.

Reference: householder.rs:59-107 (make_householder_in_place), evd/tridiag.rs:274-529, svd/bidiag.rs:47-256 (both produce the
reflectors H_k = I - v_k v_k^H / tau_k, tau real, so H_k is Hermitian and unitary), svd/mod.rs:171-273 (phase normalisation of
the complex bidiagonal), svd/mod.rs:403-429 and evd/mod.rs:411-418 (back-transforms)."""
import numpy as np

MIN_POS = np.finfo(np.float64).tiny


def make_householder(x):
    """householder.rs:59-107 on the vector x (head x[0], tail x[1:]). Returns (beta, essential, tau); tau = inf: H = I."""
    head = complex(x[0])
    tail = x[1:]
    # scaled, like the reference's three-accumulator norm (norm_l2.rs:6-172): the tails of a rank-deficient problem shrink like
    # eps^k and their squares underflow
    mx = float(np.abs(tail).max()) if tail.size else 0.0
    if not np.isfinite(mx):
        tail_norm = np.nan
    elif mx < MIN_POS:
        tail_norm = mx               # subnormal tail: below min_positive either way (tau = inf)
    else:
        tail_norm = mx * float(np.linalg.norm(tail / mx))
    head_norm = abs(head)
    if head_norm < MIN_POS:
        head, head_norm = 0.0 + 0.0j, 0.0
    if tail_norm < MIN_POS:
        return head, tail.copy(), np.inf
    norm = np.hypot(head_norm, tail_norm)
    sign = head / head_norm if head_norm != 0.0 else 1.0 + 0.0j
    signed_norm = sign * norm
    inv = 1.0 / (head + signed_norm)
    tau = 0.5 * (1.0 + (tail_norm * abs(inv)) ** 2)
    return -signed_norm, tail * inv, tau


def tridiag_unblocked(A):
    """Step list of cc_tridiag: W = full Hermitian copy of A (lower triangle read). Returns (W, taus): T on W's diagonal /
    subdiagonal (subdiagonal complex), reflector k below the subdiagonal of column k."""
    n = A.shape[0]
    L = np.tril(A)
    W = L + np.tril(L, -1).conj().T
    W[np.diag_indices(n)] = W[np.diag_indices(n)].real
    taus = np.full(max(n - 1, 0), np.inf)
    for k in range(n - 1):
        # (1) reflector of x = W[k+1:, k]
        beta, ess, tau = make_householder(W[k + 1:, k])
        W[k + 1, k] = beta
        W[k + 2:, k] = ess
        taus[k] = tau
        if not np.isfinite(tau):
            continue
        v = np.concatenate([[1.0], ess])
        A22 = W[k + 1:, k + 1:]
        # (2) p = A22 v / tau      (3) K = (v^H p) / (2 tau), w = p - K v      (4) A22 -= v w^H + w v^H
        p = (A22 @ v) / tau
        K = np.vdot(v, p) / (2.0 * tau)
        w = p - K * v
        A22 -= np.outer(v, w.conj()) + np.outer(w, v.conj())
        A22[np.diag_indices(A22.shape[0])] = A22[np.diag_indices(A22.shape[0])].real  # the diagonal stays exactly real
    return W, taus


def tridiag_phases(d, e):
    """T = D T_real D^H with D = diag(ph): ph[0] = 1, ph[k+1] = ph[k] e[k] / |e[k]| (1 for e[k] = 0)."""
    n = d.shape[0]
    ph = np.ones(n, dtype=np.complex128)
    for k in range(n - 1):
        a = abs(e[k])
        ph[k + 1] = ph[k] * (e[k] / a if a != 0.0 else 1.0)
    return ph, np.abs(e)


def apply_sequence(W, taus, M, conj=False):
    """M <- H_0 H_1 ... H_{s-1} M with reflector k = [0...0, 1, W[k+1:, k]] (rows k..), H_k = I - v v^H / tau (conj: the
    conjugated reflectors). The block size 1 case of householder.rs:724-765."""
    for k in range(len(taus) - 1, -1, -1):
        if not np.isfinite(taus[k]):
            continue
        v = np.concatenate([[1.0], W[k + 1:, k]])
        if conj:
            v = v.conj()
        M[k:, :] -= np.outer(v, (v.conj() @ M[k:, :]) / taus[k])


def self_adjoint_evd(A):
    """Returns (lam ascending, U) with A = U diag(lam) U^H."""
    n = A.shape[0]
    W, taus = tridiag_unblocked(A)
    d = W[np.diag_indices(n)].real.copy()
    e = np.array([W[k + 1, k] for k in range(n - 1)], dtype=np.complex128)
    ph, e_abs = tridiag_phases(d, e)
    T = np.diag(d) + np.diag(e_abs, -1) + np.diag(e_abs, 1)
    lam, Q = np.linalg.eigh(T)              # the GPU path: tridiag_dc_f64
    U = ph[:, None] * Q.astype(np.complex128)
    if n > 1:
        apply_sequence(W[1:, :n - 1], taus, U[1:, :])
    return lam, U


def bidiag_unblocked(A):
    """Step list of cc_bidiag (m >= n). Returns (W, taus_l, taus_r): B on W's diagonal / superdiagonal, left reflector k below
    the diagonal of column k, right reflector k right of the superdiagonal of row k (stored unconjugated)."""
    W = np.array(A, dtype=np.complex128)
    m, n = W.shape
    tl = np.full(n, np.inf)
    tr = np.full(max(n - 1, 0), np.inf)
    for k in range(n):
        # (1) left reflector of W[k:, k], (2) y = v^H W[k:, k+1:] / tau, (3) W[k:, k+1:] -= v y
        beta, ess, tau = make_householder(W[k:, k])
        W[k, k] = beta
        W[k + 1:, k] = ess
        tl[k] = tau
        if np.isfinite(tau) and k + 1 < n:
            v = np.concatenate([[1.0], ess])
            y = (v.conj() @ W[k:, k + 1:]) / tau
            W[k:, k + 1:] -= np.outer(v, y)
        if k + 1 >= n:
            break
        # (4) right reflector of the row W[k, k+1:], (5) z = W[k+1:, k+1:] conj(v) / tau, (6) W[k+1:, k+1:] -= z v^T
        beta, ess, tau = make_householder(W[k, k + 1:])
        W[k, k + 1] = beta
        W[k, k + 2:] = ess
        tr[k] = tau
        if np.isfinite(tau):
            v = np.concatenate([[1.0], ess])
            z = (W[k + 1:, k + 1:] @ v.conj()) / tau
            W[k + 1:, k + 1:] -= np.outer(z, v)
    return W, tl, tr


def bidiag_phases(d, f):
    """B = Dl B_real Dr^H: r[0] = 1, l[k] = phase(d[k] r[k]), r[k+1] = conj(phase(conj(l[k]) f[k])) (phase(0) = 1)."""
    n = d.shape[0]
    ph = lambda z: z / abs(z) if abs(z) != 0.0 else 1.0 + 0.0j
    l = np.ones(n, dtype=np.complex128)
    r = np.ones(n, dtype=np.complex128)
    for k in range(n):
        l[k] = ph(d[k] * r[k])
        if k + 1 < n:
            r[k + 1] = np.conj(ph(np.conj(l[k]) * f[k]))
    return l, r, np.abs(d), np.abs(f)


def svd(A, full=False):
    """Returns (S non-increasing, U, V) with A = U[:, :size] diag(S) V[:, :size]^H; thin (size columns) or full vectors."""
    transpose = A.shape[1] > A.shape[0]
    M = A.conj().T if transpose else A
    m, n = M.shape
    W, tl, tr = bidiag_unblocked(M)
    d = W[np.diag_indices(n)].copy()
    f = np.array([W[k, k + 1] for k in range(n - 1)], dtype=np.complex128)
    l, r, d_abs, f_abs = bidiag_phases(d, f)
    B = np.diag(d_abs) + np.diag(f_abs, 1)
    Ub, S, Vbt = np.linalg.svd(B)            # the GPU path: bidiag_svd_vectors
    ku = m if full else n
    U = np.zeros((m, ku), dtype=np.complex128)
    U[:n, :n] = l[:, None] * Ub
    for j in range(n, ku):
        U[j, j] = 1.0
    apply_sequence(W, tl, U)
    V = (r[:, None] * Vbt.T).astype(np.complex128)
    if n > 1:
        # right reflectors: basis = transpose of the rows right of the superdiagonal, applied conjugated (svd/mod.rs:413-428)
        Wt = W[:n, :].T  # (n x n): column k holds row k
        apply_sequence(Wt[1:, :n - 1], tr, V[1:, :], conj=True)
    if transpose:
        U, V = V, U
    return S, U, V


def hessenberg_unblocked(A):
    """Step list of cc::hessenberg_unblocked: H = Q^H A Q, Q = H_0 ... H_{n-2}; reflector k below the subdiagonal of column k."""
    W = np.array(A, dtype=np.complex128)
    n = W.shape[0]
    taus = np.full(max(n - 1, 0), np.inf)
    for k in range(n - 1):
        beta, ess, tau = make_householder(W[k + 1:, k])
        W[k + 1, k] = beta
        W[k + 2:, k] = ess
        taus[k] = tau
        if not np.isfinite(tau):
            continue
        v = np.concatenate([[1.0], ess])
        M = W[k + 1:, k + 1:]
        M -= np.outer(v, (v.conj() @ M) / tau)
        N = W[:, k + 1:]
        N -= np.outer((N @ v) / tau, v.conj())
    return W, taus


def t_blocks(V, taus, bs):
    """The bs x s factor of T blocks (diag tau, strict upper part of V^H V inside each block) — householder.rs:132-272."""
    m, s = V.shape
    Tf = np.zeros((bs, s), dtype=np.complex128)
    for j0 in range(0, s, bs):
        b = min(bs, s - j0)
        Vb = np.tril(V[j0:, j0:j0 + b], -1) + np.eye(m - j0, b)
        G = Vb.conj().T @ Vb
        Tf[:b, j0:j0 + b] = np.triu(G, 1) + np.diag(taus[j0:j0 + b])
    return Tf
