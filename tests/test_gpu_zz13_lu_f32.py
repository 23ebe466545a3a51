"""f32 partial-pivoting LU through the C ABI (`libfaer_v0_23_partial_piv_lu_{factor,solve,solve_transpose}_in_place_{u32,u64}_f32`).
The entry points widen the matrix to f64 on the device, factor it with the f64 drivers and round the factors back
(csrc/ffi.cu), so the contract is: a valid partial-pivoting factorization of the f32 matrix — P A = L U within the f32
backward bound, |l_ij| <= 1, permutation arrays inverse to each other — whose pivots agree with the oracle's all-f32
elimination (lu/partial_pivoting/factor.rs:19-295) except where two candidates coincide to f32 precision; where the
permutations agree the factors and the transposition count are compared with the oracle's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
U32 = 2.0 ** -24


def _factor(la, A, idx):
    m = A.shape[0]
    LU = A.copy(order="F")
    p = np.zeros(m, dtype=idx); pi = np.zeros(m, dtype=idx)
    info = la.lu_in_place(LU, p, pi)
    return LU, p.astype(np.int64), pi.astype(np.int64), info


def test_lu_f32_vs_definition_and_oracle(fb, oracle):
    la = fb.linalg
    rng = np.random.default_rng(131)
    same = total = 0
    for (m, n) in [(1, 1), (7, 7), (50, 50), (64, 200), (200, 130), (333, 333), (1000, 1000)]:
        for idx in (np.uint32, np.uint64):
            A = np.asfortranarray(rng.standard_normal((m, n)).astype(np.float32))
            LU, p, pi, info = _factor(la, A, idx)
            size = min(m, n)
            assert sorted(p.tolist()) == list(range(m)) and np.array_equal(pi[p], np.arange(m)), (m, n)
            L = np.tril(LU[:, :size].astype(np.float64), -1) + np.eye(m, size)
            Uf = np.triu(LU[:size, :].astype(np.float64))
            assert np.all(np.abs(np.tril(LU[:, :size], -1)) <= 1.0 + 4 * U32), (m, n)
            PA = A[p, :].astype(np.float64)
            bound = 8 * size * U32 * (np.abs(L) @ np.abs(Uf)) + 1e-30
            assert np.all(np.abs(PA - L @ Uf) <= bound), (m, n, float(np.max(np.abs(PA - L @ Uf) / bound)))
            want = A.copy(order="F"); po, pio, nt = oracle.lu(want)
            total += 1
            if np.array_equal(po, p):
                same += 1
                assert info.transposition_count == nt, (m, n)
                assert np.allclose(LU, want, rtol=2e-3, atol=2e-3 * np.abs(A).max()), (m, n)
    assert same >= total - 2, (same, total)  # pivots of an f64 and an f32 elimination part ways only at f32-level ties


def test_lu_f32_large_and_device_resident(fb, cuda_dev):
    """n = 4608 (the look-ahead driver on the SM partition underneath), device tensors in place."""
    import torch
    la = fb.linalg
    torch.manual_seed(132)
    n = 4608
    A = torch.randn((n, n), dtype=torch.float32, device=cuda_dev).T
    LU = A.clone(memory_format=torch.preserve_format)
    p = torch.zeros(n, dtype=torch.int64, device=cuda_dev); pi = torch.zeros(n, dtype=torch.int64, device=cuda_dev)
    la.lu_in_place(LU, p, pi)
    assert bool((pi[p] == torch.arange(n, device=cuda_dev)).all())
    x = torch.randn((n, 3), dtype=torch.float64, device=cuda_dev)
    L = torch.tril(LU.double(), -1) + torch.eye(n, dtype=torch.float64, device=cuda_dev)
    Uf = torch.triu(LU.double())
    lhs = A.double()[p, :] @ x
    rhs = L @ (Uf @ x)
    scale = float((L.abs() @ (Uf.abs() @ x.abs())).max())
    assert float((lhs - rhs).abs().max()) <= 16 * n * U32 * scale


def test_lu_f32_solves(fb):
    la = fb.linalg
    rng = np.random.default_rng(133)
    for n, k in [(40, 3), (300, 17), (900, 5)]:
        A = np.asfortranarray((rng.standard_normal((n, n)) + 4 * np.eye(n)).astype(np.float32))
        LU, p, pi, _ = _factor(la, A, np.uint64)
        pu, piu = p.astype(np.uint64), pi.astype(np.uint64)
        B = np.asfortranarray(rng.standard_normal((n, k)).astype(np.float32))
        X = B.copy(order="F"); la.lu_solve_in_place(LU, pu, piu, X)
        Xt = B.copy(order="F"); la.lu_solve_transpose_in_place(LU, pu, piu, Xt)
        A64 = A.astype(np.float64)
        cond = np.linalg.cond(A64)
        tol = 64 * n * U32 * cond
        assert np.max(np.abs(A64 @ X - B)) <= tol * np.max(np.abs(B)), (n, k)
        assert np.max(np.abs(A64.T @ Xt - B)) <= tol * np.max(np.abs(B)), (n, k)
