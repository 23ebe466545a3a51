"""Host-side logic of faer_b200.solvers on the CPU: `solvers.la` (the C-ABI mirror, which needs a GPU) is swapped for a
stand-in with the same function signatures backed by the oracle, and the shared cases of tests/solvers_cases.py run
against it. This checks what solvers.py itself does (ownership, split_LU, triangle selection, call order, shapes,
error propagation); tests/test_gpu_zz3_solvers.py runs the same cases through libfaer_b200.so on the GPU."""
import types

import numpy as np
import pytest

from solvers_cases import run_all, run_ldlt


def oracle_backed_la(fb, oracle):
    real = fb.linalg

    def cholesky_in_place(A, regularization=(0.0, 0.0), par=None, params=None):
        fail, count = oracle.llt(A, regularization[0], regularization[1])
        if fail >= 0:
            raise real.LltError(fail)
        return count

    def llt_solve_in_place(L, rhs, conj=0, par=None):
        # conj?(L) y = rhs, then conj?(L)^H x = y (llt/solve.rs:12-35)
        oracle.solve_triangular(L, rhs, lower=True, unit=False, conj=bool(conj))
        oracle.solve_triangular(L.T, rhs, lower=False, unit=False, conj=not conj)

    def ldlt_in_place(A, regularization=(0.0, 0.0), signs=None, par=None, params=None):
        fail, count = oracle.ldlt(A, regularization[0], regularization[1], signs)
        if fail >= 0:
            raise real.LdltError(fail)
        return count

    def ldlt_solve_in_place(LD, rhs, conj=0, par=None, D=None):
        packed = np.array(LD, order="F", copy=True)
        if D is not None:
            np.fill_diagonal(packed, D)
        oracle.ldlt_solve(packed, rhs, conj_lhs=bool(conj))

    def lu_in_place(A, perm, perm_inv, par=None, params=None):
        p, pi, _ = oracle.lu(A)
        perm[...] = p
        perm_inv[...] = pi

    def lu_solve_in_place(LU, perm, perm_inv, rhs, conj=0, par=None, U=None):
        rhs[...] = rhs[np.asarray(perm, dtype=np.int64)]
        oracle.solve_triangular(LU, rhs, lower=True, unit=True, conj=bool(conj))
        oracle.solve_triangular(LU if U is None else U, rhs, lower=False, unit=False, conj=bool(conj))

    def lu_solve_transpose_in_place(LU, perm, perm_inv, rhs, conj=0, par=None, U=None):
        oracle.solve_triangular((LU if U is None else U).T, rhs, lower=True, unit=False, conj=bool(conj))
        oracle.solve_triangular(LU.T, rhs, lower=False, unit=True, conj=bool(conj))
        rhs[...] = rhs[np.asarray(perm_inv, dtype=np.int64)]

    def qr_in_place(A, Q_coeff, par=None, params=None):
        H, rank = oracle.qr(A, block_size=Q_coeff.shape[0])
        Q_coeff[...] = H
        return real.QrInfo(rank)

    def _packed(Q_basis, R):
        # the oracle's solves take the packed QR matrix; rebuild it from the split factors
        QR = np.array(Q_basis, order="F", copy=True)
        size = min(QR.shape)
        QR[:size, :] = np.tril(QR[:size, :], -1) + np.triu(np.asarray(R)[:size, :QR.shape[1]])
        return QR

    def qr_solve_lstsq_in_place(Q_basis, Q_coeff, R, rhs, conj=0, par=None):
        oracle.qr_solve_lstsq(_packed(Q_basis, R), Q_coeff, rhs, conj_QR=bool(conj))

    def qr_solve_in_place(Q_basis, Q_coeff, R, rhs, conj=0, par=None):
        oracle.qr_solve(_packed(Q_basis, R), Q_coeff, rhs, conj_QR=bool(conj))

    def qr_solve_transpose_in_place(Q_basis, Q_coeff, R, rhs, conj=0, par=None):
        oracle.qr_solve_transpose(_packed(Q_basis, R), Q_coeff, rhs, conj_QR=bool(conj))

    def apply_seq(basis, factor, rhs, conj=0, par=None):
        oracle.apply_q_sequence(basis, factor, rhs, conj_lhs=bool(conj))

    def matmul(dst, accum, lhs, rhs, alpha, par=None):
        oracle.matmul(dst, accum == real.Accum.Add, lhs, rhs, alpha)

    def matmul_triangular(dst, ds, accum, lhs, ls, rhs, rs, alpha, par=None):
        oracle.matmul_triangular(dst, ds, accum == real.Accum.Add, lhs, ls, rhs, rs, alpha)

    return types.SimpleNamespace(
        Accum=real.Accum, BlockStructure=real.BlockStructure, LltError=real.LltError, CONJ_NO=real.CONJ_NO, CONJ_YES=real.CONJ_YES,
        cholesky_in_place=cholesky_in_place, llt_solve_in_place=llt_solve_in_place, lu_in_place=lu_in_place,
        ldlt_in_place=ldlt_in_place, ldlt_solve_in_place=ldlt_solve_in_place, LdltError=real.LdltError,
        lu_solve_in_place=lu_solve_in_place, lu_solve_transpose_in_place=lu_solve_transpose_in_place,
        qr_recommended_block_size=oracle.qr_recommended_block_size, qr_in_place=qr_in_place,
        qr_solve_lstsq_in_place=qr_solve_lstsq_in_place, qr_solve_in_place=qr_solve_in_place,
        qr_solve_transpose_in_place=qr_solve_transpose_in_place,
        apply_block_householder_sequence_on_the_left_in_place=apply_seq, matmul=matmul, matmul_triangular=matmul_triangular)


def test_solvers_host_logic_against_oracle_backend(fb, oracle, monkeypatch):
    sv = fb.solvers
    monkeypatch.setattr(sv, "la", oracle_backed_la(fb, oracle))
    run_all(sv)
    run_all(sv, cplx=True)  # the reference's test_all_solvers runs on c64 (solvers.rs:2919-2977)
    run_ldlt(sv)


def test_ldlt_class_on_complex_and_f32(fb, oracle, monkeypatch):
    """The host logic of `solvers.Ldlt` beyond f64 (conjugation flags of the four solves, the adjoint side, L D L^H): the GPU
    test's own function against the oracle-backed stand-in."""
    import importlib
    sv = fb.solvers
    monkeypatch.setattr(sv, "la", oracle_backed_la(fb, oracle))
    gpu_case = importlib.import_module("test_gpu_zzzzzzzzz_3_ldlt_types").test_ldlt_solver_class_other_dtypes
    for dtype in (np.complex128, np.float32):
        gpu_case(fb, None, dtype)


def test_split_lu_contract(fb):
    """solvers.rs:955-980, pure host logic."""
    sv = fb.solvers
    rng = np.random.default_rng(5)
    for (m, n) in [(6, 4), (4, 6), (5, 5), (1, 3), (3, 1)]:
        LU = np.asfortranarray(rng.standard_normal((m, n)))
        keep = LU.copy()
        L, U = sv.split_LU(LU)
        size = min(m, n)
        assert L.shape == ((m, n) if m >= n else (size, size)) and U.shape == ((size, size) if m >= n else (m, n))
        assert np.array_equal(np.tril(L, -1), np.tril(keep, -1)[:L.shape[0], :L.shape[1]])
        assert np.all(np.diag(L) == 1) and np.all(np.triu(L, 1) == 0)
        assert np.array_equal(np.triu(U), np.triu(keep)[:U.shape[0], :U.shape[1]]) and np.all(np.tril(U, -1) == 0)
        assert L.flags.f_contiguous and U.flags.f_contiguous
