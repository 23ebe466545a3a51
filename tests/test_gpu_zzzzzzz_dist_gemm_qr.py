"""SURVEY.md 8e rows beyond LLT / LU on one rank (the P > 1 runs: tools/dist_parity.py under torchrun, tests/test_gpu_dist_multi.py;
the schedules on CPU: tests/test_dist_cpu.py): the distributed QR driver (csrc/dist.cu::dist_qr_impl) as a single-rank run against
the single-GPU entry point and the oracle, the rank-deficient refusal, and the column-split GEMM front end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dist_qr_single_rank_vs_single_gpu_entry(fb, oracle, cuda_dev, dtype):
    import torch
    la, lay = fb.linalg, fb.dist
    rng = np.random.default_rng(171)
    u = np.finfo(dtype).eps
    for (m, n, bs) in [(300, 200, 32), (1000, 130, 64), (512, 512, 128), (700, 90, 16)]:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(dtype))
        loc = torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda_dev).T  # column-major device matrix
        H = lay.qr_in_place(loc, m, n, bs)
        got = loc.cpu().numpy(); Hg = H.cpu().numpy()
        QR = A.copy(order="F"); H1 = np.zeros((bs, n), dtype=dtype, order="F")
        assert la.qr_in_place(QR, H1).rank == n
        loose = 2e3 * u * max(m, n)
        assert np.allclose(got, QR, rtol=loose, atol=loose * np.abs(A).max()), (m, n, bs)
        QRo = A.copy(order="F"); Ho, rank_o = oracle.qr(QRo, block_size=bs)
        assert rank_o == n
        for j in range(0, n, bs):
            b = min(bs, n - j)
            assert np.allclose(np.triu(Hg[:b, j:j + b]), np.triu(H1[:b, j:j + b]), rtol=loose, atol=loose), (m, n, bs, j)
            assert np.allclose(np.triu(Hg[:b, j:j + b]), np.triu(Ho[:b, j:j + b]), rtol=loose, atol=loose), (m, n, bs, j)
        # Q R = A through the library's block-Householder sequence on the distributed driver's output
        Q = np.asfortranarray(np.eye(m, dtype=dtype))
        la.apply_block_householder_sequence_on_the_left_in_place(np.asfortranarray(got), np.asfortranarray(Hg), Q)
        tol = 128 * u * np.sqrt(8 * max(m, n)) * max(1.0, float(np.abs(A).max()))
        assert np.all(np.abs(Q @ np.triu(got) - A) <= tol), (m, n, bs)


def test_dist_qr_refuses_rank_deficient_blocks(fb, cuda_dev):
    import torch
    lay = fb.dist
    rng = np.random.default_rng(172)
    A = rng.standard_normal((200, 20)) @ rng.standard_normal((20, 96))
    loc = torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda_dev).T
    with pytest.raises(RuntimeError):
        lay.qr_in_place(loc, 200, 96, 32)


def test_dist_matmul_single_rank(fb, cuda_dev):
    import torch
    la, lay = fb.linalg, fb.dist
    rng = np.random.default_rng(173)
    m, k, n = 300, 200, 170
    A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n)); C0 = rng.standard_normal((m, n))
    a, b = lay.column_slab(n, 1, 0)
    assert (a, b) == (0, n)
    dA = torch.from_numpy(A).to(cuda_dev); dB = torch.from_numpy(B).to(cuda_dev)
    dC = torch.from_numpy(np.ascontiguousarray(C0.T)).to(cuda_dev).T
    lay.matmul(dC, la.Accum.Add, dA, dB, 0.5, src_rank=0)
    bound = 2 * k * 2.0 ** -53 * 0.5 * (np.abs(A) @ np.abs(B)) + 4 * 2.0 ** -53 * np.abs(C0)
    assert np.all(np.abs(dC.cpu().numpy() - (C0 + 0.5 * A @ B)) <= bound)
