"""Dev aid: per-shape error report of bidiag / tridiag against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import faer_b200  # noqa: E402
import oracle  # noqa: E402
la = faer_b200.linalg
rng = np.random.default_rng(7)
np.set_printoptions(linewidth=200, precision=4)
for dtype in (np.float64,):
    for (m, n, bl, br) in [(8, 4, 4, 3), (8, 8, 4, 3), (1, 1, 1, 1), (2, 2, 1, 1), (5, 1, 2, 1), (33, 17, 8, 8), (64, 64, 16, 5),
                           (130, 97, 32, 32), (300, 300, 32, 16), (1000, 37, 8, 8), (513, 512, 64, 64), (3000, 2048, 64, 64)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        want = A.copy(order="F")
        if m * n * n < 3e8:
            Hl_w, Hr_w = oracle.bidiag(want, bl, br)
        else:
            want = None
        got = A.copy(order="F")
        Hl = np.zeros((bl, n), dtype=dtype, order="F")
        Hr = np.zeros((br, max(n - 1, 0)), dtype=dtype, order="F")
        la.bidiag_in_place(got, Hl, Hr)
        d = np.diagonal(got).copy(); e = np.diagonal(got, 1).copy()
        B = np.zeros((n, n)); B[np.arange(n), np.arange(n)] = d
        if n > 1:
            B[np.arange(n - 1), np.arange(1, n)] = e
        sv_a = np.linalg.svd(A.astype(np.float64), compute_uv=False); sv_b = np.linalg.svd(B, compute_uv=False)
        msg = f"bidiag {m}x{n}: finite={np.all(np.isfinite(got))} sv_err={np.abs(sv_a - sv_b).max():.2e}"
        if want is not None:
            diff = np.abs(got - want)
            diff[~np.isfinite(diff)] = 1e300
            i, j = np.unravel_index(np.argmax(diff), diff.shape)
            fh = np.isfinite(Hl_w)
            msg += f" |A-oracle|max={diff.max():.2e} at {(i, j)} Hl_err={(np.abs(Hl - Hl_w)[fh].max() if fh.any() else 0):.2e}"
            if n > 1:
                fr = np.isfinite(Hr_w)
                msg += f" Hr_err={np.abs(Hr - Hr_w)[fr].max() if fr.any() else 0:.2e} inf_eq={np.array_equal(np.isinf(Hr), np.isinf(Hr_w))}"
            firstbad = np.argwhere(diff > 1e-9)
            if firstbad.size:
                msg += f" first_bad_col={firstbad[:, 1].min()} first_bad_row={firstbad[:, 0].min()}"
        print(msg, flush=True)
    for n, b in [(2, 3), (3, 3), (4, 3), (8, 3), (16, 3), (1, 1), (45, 8), (100, 32), (257, 16), (700, 64)]:
        Gm = rng.standard_normal((n, n)).astype(dtype)
        A = np.asfortranarray(Gm + Gm.T)
        want = A.copy(order="F"); H_w = oracle.tridiag(want, b)
        got = A.copy(order="F"); H = np.zeros((b, max(n - 1, 0)), dtype=dtype, order="F")
        la.tridiag_in_place(got, H)
        lo = np.tril_indices(n)
        diff = np.abs(np.tril(got) - np.tril(want)); diff[~np.isfinite(diff)] = 1e300
        i, j = np.unravel_index(np.argmax(diff), diff.shape)
        dd = np.diagonal(got); ee = np.diagonal(got, -1)
        T = np.diag(dd) + np.diag(ee, -1) + np.diag(ee, 1)
        ev = np.abs(np.linalg.eigvalsh(A) - np.linalg.eigvalsh(T)).max() if np.all(np.isfinite(T)) else np.nan
        fh = np.isfinite(H_w)
        herr = np.abs(H - H_w)[fh].max() if fh.any() else 0.0
        firstbad = np.argwhere(diff > 1e-9)
        print(f"tridiag n={n}: ev_err={ev:.2e} |A-oracle|max={diff.max():.2e} at {(i, j)} H_err={herr:.2e} inf_eq={np.array_equal(np.isinf(H), np.isinf(H_w))}"
              + (f" first_bad_col={firstbad[:, 1].min()}" if firstbad.size else ""), flush=True)
        if n <= 4 and diff.max() > 1e-9:
            print("got\n", np.tril(got), "\nwant\n", np.tril(want), "\nH", H, "\nH_w", H_w)
