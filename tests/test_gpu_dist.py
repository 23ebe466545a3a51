"""GPU tests of the distributed (1-D block-column-cyclic) LLT driver, csrc/dist.cu.

Single process (P = 1, no communicator): the same code path the multi-GPU runs use (panel pack, per-block-column updates,
look-ahead stream), compared with the oracle. The P = 2 run over NCCL is `tools/dist_check.py` under torchrun (used with
`gpurun --gpus 2`), and the layout/schedule logic is covered on CPU by tests/test_dist_cpu.py (gloo, world size 2).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def spd(rng, n):
    G = rng.standard_normal((n, n))
    return np.asfortranarray(G @ G.T + n * np.eye(n))


@pytest.mark.parametrize("lookahead", [True, False])
def test_dist_llt_single_rank_vs_oracle(fb, oracle, cuda_dev, lookahead):
    import torch
    rng = np.random.default_rng(11)
    U = np.finfo(np.float64).eps / 2
    for n, nb in [(64, 16), (100, 32), (512, 128), (1000, 256), (1536, 512)]:
        A = spd(rng, n)
        want = A.copy(order="F"); fail, _ = oracle.llt(want); assert fail == -1
        dA = torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda_dev).T  # column-major device matrix
        fail, cnt = fb.dist.cholesky_in_place(dA, n, nb=nb, lookahead=lookahead)
        assert fail == -1 and cnt == 0
        got = dA.cpu().numpy()
        assert np.array_equal(np.triu(got, 1), np.triu(A, 1))
        L = np.tril(got)
        assert np.all(np.abs(L @ L.T - A) <= 8 * n * 128 * U * np.abs(A).max()), (n, nb)
        assert np.allclose(L, np.tril(want), rtol=1e-10, atol=1e-10 * np.sqrt(np.abs(A).max())), (n, nb)


def test_dist_llt_failure_index(fb, cuda_dev):
    import torch
    rng = np.random.default_rng(12)
    n = 700
    A = spd(rng, n); A[600, 600] = -5.0
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda_dev).T
    fail, _ = fb.dist.cholesky_in_place(dA, n, nb=128)
    assert fail == 600


@pytest.mark.parametrize("lookahead", [True, False])
def test_dist_lu_single_rank_vs_oracle(fb, oracle, cuda_dev, lookahead):
    """Permutations bit-exact vs the oracle (and hence vs the single-GPU path), factors within tolerance."""
    import torch
    rng = np.random.default_rng(13)
    U = np.finfo(np.float64).eps / 2
    for n, nb in [(64, 16), (100, 32), (512, 128), (1000, 256), (1536, 512)]:
        A = np.asfortranarray(rng.standard_normal((n, n)))
        want = A.copy(order="F"); perm_o, pinv_o, nt_o = oracle.lu(want)
        dA = torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda_dev).T
        perm, pinv, nt = fb.dist.lu_in_place(dA, n, nb=nb, lookahead=lookahead)
        assert np.array_equal(perm, perm_o) and np.array_equal(pinv, pinv_o) and nt == nt_o, (n, nb)
        got = dA.cpu().numpy()
        L = np.tril(got, -1) + np.eye(n); Um = np.triu(got)
        growth = max(1.0, np.abs(Um).max() / np.abs(A).max())
        assert np.all(np.abs(L @ Um - A[perm_o, :]) <= 8 * n * 128 * U * np.abs(A).max() * growth), (n, nb)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * growth), (n, nb)
