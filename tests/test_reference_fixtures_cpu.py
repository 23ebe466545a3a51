"""Pins the oracle (and the host/device bisection code) on the fixtures the reference itself holds for the path
(tests/golden/make_reference_fixtures.py extracted them from /root/reference):

  * `test_rank_deficient` (qr/no_pivoting/factor.rs:540-4787): the 100 x 40 c64 matrix of numerical rank 33, block size 20,
    `Q R ~ A` with ApproxEq{abs 1e-10, rel 1e-10} — run through the oracle's c64 QR (the column-skipping path of
    factor.rs:40-83 is what this matrix exercises);
  * `test_zink` (svd/mod.rs:985-1054): the graded 20-point bidiagonal whose smallest singular value must not collapse to 0;
  * `faer/test_data/svd/*.txt` (bidiag_svd.rs:1526-1606): bidiagonals of order 64 ... 8660; the singular values of the
    n x n part (subdiag[n-1] = 0, as test_qr_algorithm sets it) against LAPACK.
"""
import json
import os

import numpy as np
import pytest

from test_bidiag_sv_cpu import bsv  # noqa: F401  (fixture: csrc/bidiag_sv.cuh compiled for the host)

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def approx_eq(a, b, abs_tol, rel_tol):
    """utils/approx.rs:48-57: |a - b| <= abs_tol or <= rel_tol * max(|a|, |b|), element-wise."""
    d = np.abs(a - b)
    return np.all((d <= abs_tol) | (d <= rel_tol * np.maximum(np.abs(a), np.abs(b))))


def test_rank_deficient_c64_matrix(oracle):
    A = np.asfortranarray(np.load(os.path.join(GOLD, "qr_rank_deficient_c64.npz"))["A"])
    m, n = A.shape
    assert (m, n) == (100, 40) and A.dtype == np.complex128
    QR = A.copy(order="F")
    H, rank = oracle.qr(QR, block_size=20)
    # singular values 26..40 decay from 1.8e-10 to 7.7e-12 (|A| ~ 50): the rank is fuzzy; the reference's test only
    # requires Q R ~ A. The oracle stops at 35 reflectors.
    assert 25 <= rank < 40, rank
    Q = np.asfortranarray(np.eye(m, dtype=A.dtype))
    oracle.apply_q_sequence(QR, H, Q)
    R = np.triu(QR)
    assert approx_eq(Q @ R, A, 1e-10, 1e-10)
    assert approx_eq(Q.conj().T @ Q, np.eye(m), 1e-10, 1e-10)
    # Q_coeff beyond the rank: zero columns with +inf on the block diagonals (factor.rs:287-299)
    for c in range(rank, n):
        col = H[:, c]
        assert np.isinf(col[c % 20].real) and np.count_nonzero(col) == 1
    # the same matrix as a real problem (the GPU path's dtype): the real embedding [[Re, -Im], [Im, Re]], 200 x 80
    E = np.asfortranarray(np.block([[A.real, -A.imag], [A.imag, A.real]]))
    QRe = E.copy(order="F")
    He, rank_e = oracle.qr(QRe, block_size=20)
    assert 50 <= rank_e < 80, rank_e
    Qe = np.asfortranarray(np.eye(2 * m))
    oracle.apply_q_sequence(QRe, He, Qe)
    assert approx_eq(Qe @ np.triu(QRe), E, 1e-10, 1e-10)


def test_zink_bidiagonal(bsv):  # noqa: F811
    fx = json.load(open(os.path.join(GOLD, "svd_zink.json")))
    d = np.array(fx["diag"]); s = np.array(fx["subdiag"])
    assert s[-1] == 0.0
    sv = bsv(d, s[:-1])
    assert sv[-1] != 0.0 and np.all(sv > 0) and np.all(np.diff(sv) <= 0)
    # LAPACK on the same lower-bidiagonal matrix (normwise accurate only): agreement to eps * |B|
    B = np.diag(d) + np.diag(s[:-1], -1)
    ref = np.linalg.svd(B, compute_uv=False)
    assert np.all(np.abs(sv - ref) <= 64 * np.finfo(float).eps * ref[0])
    # determinant identity pins the SMALL values too: prod(sigma) = prod |d_i| for a bidiagonal matrix
    assert abs(np.sum(np.log(sv)) - np.sum(np.log(np.abs(d)))) <= 1e-9


@pytest.mark.parametrize("name", ["svd64", "svd128", "svd512", "svd1024_0", "svd1024_1", "svd1024_2", "svd_josef"])
def test_reference_bidiagonals(bsv, name):  # noqa: F811
    fx = np.load(os.path.join(GOLD, f"svd_bidiag_{name}.npz"))
    d, s = fx["diag"], fx["subdiag"]
    n = d.size
    assert s.size == n
    if n > 2048:
        d, s, n = d[:2048].copy(), s[:2048].copy(), 2048  # keep the CPU suite short; the GPU test runs the whole file
    sv = bsv(d, s[:-1])
    B = np.diag(d) + np.diag(s[:-1], -1)
    ref = np.linalg.svd(B, compute_uv=False)
    assert np.all(np.diff(sv) <= 0) and np.all(sv >= 0)
    assert np.all(np.abs(sv - ref) <= 8 * n * np.finfo(float).eps * ref[0]), name
