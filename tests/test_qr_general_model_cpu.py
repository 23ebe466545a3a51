"""Host model of the GPU column-skipping QR path (faer-rs_b200/csrc/qr.cu: `qr_panel_general_kernel` + `qr_general_from`
+ `qr_coeff_fixup`) against the oracle's restatement of the reference (qr/no_pivoting/factor.rs:11-301).

The model follows the kernel's own data flow — panel columns and reflector slots as separate arrays, pivot row `lr`
and panel column `j` as separate counters, the zero fill of factor.rs:46-48, the write-back rules — and the driver's
blocking (sub-panels of <= 16 columns, T blocks indexed by reflector number, immediate application to every column on
the right). It pins the DESIGN of the device code on the CPU: rank exact, R staircase, compacted V and the T blocks
equal to the oracle's within rounding, on the reference's `test_qr` inputs (rank-deficient products A0 * A1).
"""
import numpy as np
import pytest

QRG_PW = 16


def panel_general(A, row0, col0, w, max_refl, taus_out):
    """One launch of qr_panel_general_kernel on A (modified in place). Returns `local`."""
    m = A.shape[0]
    mp = m - row0
    d0 = col0 - row0
    S = A[row0:, col0:col0 + w].copy()
    SV = np.zeros((mp, w))
    above2 = (A[:row0, col0:col0 + w] ** 2).sum(axis=0)
    fi = np.finfo(A.dtype)
    min_pos, eps = fi.tiny, fi.eps
    lr, jend = 0, 0
    for j in range(w):
        if lr >= max_refl:
            break
        jend = j + 1
        tail = S[lr + 1:, j]
        tail_norm = np.sqrt((tail ** 2).sum())
        dots = tail @ S[lr + 1:, :]
        above = (S[:lr, j] ** 2).sum()
        rowj = S[lr, :].copy()
        head = rowj[j]
        head_norm = abs(head)
        if head_norm < min_pos:
            head, head_norm = 0.0, 0.0
        no_tail = tail_norm < min_pos
        inv, new_head = 0.0, head
        if no_tail:
            tau, norm = np.inf, head_norm
        else:
            norm = np.hypot(head_norm, tail_norm)
            sign = head / head_norm if head_norm != 0 else 1.0
            inv = 1.0 / (head + sign * norm)
            new_head = -sign * norm
            tau = 0.5 * (1.0 + (tail_norm * abs(inv)) ** 2)
        total = np.hypot(norm, np.sqrt(above2[j] + above))
        threshold = eps * ((mp - lr) * 16.0) * total
        tau_inv = 1.0 / tau
        apply, advance = False, False
        if tau_inv < min_pos:
            advance = norm > 0
        elif norm > threshold:
            apply = advance = True
        taus_out[lr] = tau
        kc = np.zeros(w)
        if apply:
            kc[j + 1:] = -((rowj[j + 1:] + inv * dots[j + 1:]) * tau_inv)
        gap = d0 + j - lr
        S[lr, j] = head if no_tail else new_head
        S[lr, j + 1:] += kc[j + 1:]
        v = np.zeros(mp - lr - 1) if no_tail else S[lr + 1:, j] * inv
        SV[lr + 1:, lr] = v
        S[lr + 1:lr + 1 + gap, j] = 0.0
        S[lr + 1:, j + 1:] += np.outer(v, kc[j + 1:])
        if advance:
            lr += 1
    local = lr
    for c in range(w):
        hi = mp if c >= jend else min(mp, d0 + c + 1)
        A[row0:row0 + hi, col0 + c] = S[:hi, c]
    for l in range(local):
        A[row0 + l + 1:, row0 + l] = SV[l + 1:, l]
    return local


def build_t(V, T):
    """striu(T) <- striu(V^T V) for unit-lower V; diagonal untouched (householder.cu: householder_build_t)."""
    k = V.shape[1]
    Vu = np.tril(V, -1).copy()
    Vu[np.arange(k), np.arange(k)] = 1.0
    G = Vu.T @ Vu
    iu = np.triu_indices(k, 1)
    T[iu] = G[iu]


def apply_left(V, T, M):
    """M <- (I - V T^-T V^T) M, V unit-lower (householder.rs:370-620, forward = true)."""
    k = V.shape[1]
    Vu = np.tril(V, -1).copy()
    Vu[np.arange(k), np.arange(k)] = 1.0
    W = Vu.T @ M
    Tt = np.triu(T).T
    for i in range(k):
        W[i] = (W[i] - Tt[i, :i] @ W[:i]) / Tt[i, i]
    M -= Vu @ W


def qr_general_model(A, bs):
    m, n = A.shape
    size = min(m, n)
    H = np.zeros((bs, size), order="F")
    row = col = 0
    while row < size and col < n:
        start, offset, pieces = row, 0, 0
        blk = min(bs, size - row, n - col)
        while offset < blk and col < n:
            w = min(QRG_PW, blk - offset, n - col)
            taus = np.zeros(w + 1)
            local = panel_general(A, row, col, w, min(w, size - row), taus)
            for l in range(local):
                H[offset + l, row + l] = taus[l]
            if local > 0:
                Vs = A[row:, row:row + local]
                Tss = H[offset:offset + local, row:row + local]
                build_t(Vs, Tss)
                if n - (col + w) > 0:
                    apply_left(Vs, Tss, A[row:, col + w:])
                pieces += 1
            offset += local; row += local; col += w
        if pieces > 1:
            build_t(A[start:, start:start + offset], H[:offset, start:start + offset])
    rank = row
    for c in range(rank, size):
        H[:, c] = 0.0
        H[c % bs, c] = np.inf
    return H, rank


@pytest.mark.parametrize("shape", [(2, 2), (3, 3), (8, 8), (24, 24), (32, 32), (128, 128), (255, 255), (257, 257), (8, 4),
                                   (128, 20), (257, 20), (300, 64), (20, 50), (90, 200)])
def test_model_matches_oracle_on_rank_deficient_products(oracle, shape):
    rng = np.random.default_rng(7)
    m, n = shape
    size = min(m, n)
    for rank_true in sorted({1, 2, 3, 5, 17, 100} & set(range(1, size))) + [size]:
        if rank_true < size:
            A = np.asfortranarray(rng.standard_normal((m, rank_true)) @ rng.standard_normal((rank_true, n)))
        else:
            A = np.asfortranarray(rng.standard_normal((m, n)))
        for bs in sorted({1, min(15, size), min(40, size), oracle.qr_recommended_block_size(m, n)}):
            QRo = A.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=bs)
            QR = A.copy(order="F"); H, rank = qr_general_model(QR, bs)
            # the reference's own criterion (factor.rs:376, 385-404): rank >= true rank and Q R ~ A (1e-10)
            assert rank >= min(rank_true, size), (shape, rank_true, bs, rank)
            Q = np.eye(m)
            for j in reversed(range(0, size, bs)):
                b = min(bs, size - j)
                Vb = np.tril(QR[j:, j:j + b], -1); Vb[np.arange(min(b, m - j)), np.arange(min(b, m - j))] = 1.0
                W = Vb.T @ Q[j:, :]
                Tb = np.triu(H[:b, j:j + b])
                for i in reversed(range(b)):
                    W[i] = (W[i] - Tb[i, i + 1:] @ W[i + 1:]) / Tb[i, i]
                Q[j:, :] -= Vb @ W
            sc = max(1.0, np.abs(A).max())
            assert np.all(np.abs(Q @ np.triu(QR) - A) <= 1e-10 * sc * max(m, n)), (shape, rank_true, bs)
            if rank != rank_o:
                # both are legal outcomes: a column whose tail is EXACTLY zero after the previous reflectors but whose head
                # is rounding noise yields an identity reflector that still advances `row` (factor.rs:60-63); whether the
                # tail is exactly zero depends on the summation order of the dot products. Tiny matrices only.
                assert max(m, n) <= 4 and abs(rank - rank_o) <= 2, (shape, rank_true, bs, rank, rank_o)
                continue
            tol = 1e-9 * max(m, n)
            assert np.allclose(np.triu(QR), np.triu(QRo), rtol=tol, atol=tol * sc), (shape, rank_true, bs)
            assert np.array_equal(np.isinf(H), np.isinf(Ho)), (shape, rank_true, bs)
            # reflectors with tau = +inf are the identity: their v (stale data in the reference: the early return of
            # make_householder, householder.rs:73-79, writes nothing) and their T row / column are never interpreted
            # ... and reflectors beyond the true rank are built from rounding noise: their direction is arbitrary
            live = np.array([np.isfinite(Ho[c % bs, c]) and c < rank_true for c in range(rank)], dtype=bool)
            V = np.tril(QR, -1)[:, :rank][:, live]; Vo = np.tril(QRo, -1)[:, :rank][:, live]
            assert np.allclose(V, Vo, rtol=tol, atol=tol), (shape, rank_true, bs)
            for j in range(0, rank, bs):
                b = min(bs, rank - j)
                lv = live[j:j + b]
                Tg = np.triu(H[:b, j:j + b])[np.ix_(lv, lv)]; To = np.triu(Ho[:b, j:j + b])[np.ix_(lv, lv)]
                assert np.allclose(Tg, To, rtol=tol, atol=tol), (shape, rank_true, bs, j)
