"""CPU tests of the multi-GPU host logic with a world-size-2 gloo group (no GPU needed).

What is covered here: the block-column-cyclic layout helpers (faer-rs_b200/dist.py) and the step/ownership/broadcast
schedule of the distributed LLT and LU, executed with the CPU ORACLE as the per-block compute (test infrastructure) and
gloo as the transport. The schedule below is the one csrc/dist.cu runs on the GPUs (same panel order, same owner rule,
same update order per block column), so a layout/ownership/ordering bug shows up here. The CUDA implementation itself is
exercised by tests/test_gpu_dist.py (-m gpu) and by the multi-rank runs of bench.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nb, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import faer_b200
    from oracle import oracle as orc
    lay = faer_b200._pkg.dist if hasattr(faer_b200, "_pkg") else None
    from faer_rs_b200 import dist as lay  # noqa: F811  (registered by the faer_b200 shim)

    rng = np.random.default_rng(0)  # same seed on every rank: replicated global input
    G = rng.standard_normal((n, n))
    A = np.asfortranarray(G @ G.T + n * np.eye(n))
    loc = lay.scatter_block_cyclic(A, nb, world, rank)
    assert loc.shape == (n, lay.local_cols(n, nb, world, rank))

    # ---- distributed LLT schedule (mirrors csrc/dist.cu::dist_llt_f64) ----
    nblk = lay.num_blocks(n, nb)
    for k in range(nblk):
        k0 = k * nb; kb = min(nb, n - k0); owner = lay.owner_of_block(k, world)
        W = np.zeros((n - k0, kb), order="F")
        if owner == rank:
            off = lay.local_col_offset(k, nb, world)
            panel = loc[k0:, off:off + kb]
            diag = np.asfortranarray(panel[:kb, :])
            fail, _ = orc.llt(diag)
            assert fail == -1
            panel[:kb, :] = diag
            if n - k0 > kb:
                below_t = np.ascontiguousarray(panel[kb:, :]).T  # (kb x rows) view, solve conj(L) X = A10^T
                orc.solve_triangular(np.asfortranarray(np.tril(diag)), below_t, lower=True, unit=False)
                panel[kb:, :] = below_t.T
            W[:, :] = panel
            W[:kb, :] = np.tril(W[:kb, :])
        t = torch.from_numpy(np.ascontiguousarray(W))
        dist.broadcast(t, src=owner)
        W = np.asfortranarray(t.numpy())
        for j in range(k + 1, nblk):
            if lay.owner_of_block(j, world) != rank:
                continue
            j0 = j * nb; jb = min(nb, n - j0); off = lay.local_col_offset(j, nb, world)
            Wj = W[j0 - k0:j0 - k0 + jb, :]
            blk = loc[j0:, off:off + jb]
            full = blk - W[j0 - k0:, :] @ Wj.T
            # diagonal block: lower triangle only (strict upper part untouched)
            d = blk[:jb, :].copy()
            blk[:, :] = full
            iu = np.triu_indices(jb, 1)
            blk[:jb, :][iu] = d[iu]
    np.save(os.path.join(out_dir, f"llt_{rank}.npy"), loc)

    # ---- distributed LU schedule (panel factor on the owner, broadcast panel + transpositions, swaps everywhere) ----
    rng = np.random.default_rng(1)
    M = np.asfortranarray(rng.standard_normal((n, n)))
    loc = lay.scatter_block_cyclic(M, nb, world, rank)
    trans_all = np.zeros(n, dtype=np.int64)
    for k in range(nblk):
        k0 = k * nb; kb = min(nb, n - k0); owner = lay.owner_of_block(k, world)
        W = np.zeros((n - k0, kb), order="F"); piv = np.zeros(kb, dtype=np.int64)
        if owner == rank:
            off = lay.local_col_offset(k, nb, world)
            panel = np.asfortranarray(loc[k0:, off:off + kb])
            perm, _, _ = orc.lu(panel)
            # recover the transposition sequence from the permutation (replay: perm.swap(i, i + t_i))
            cur = np.arange(n - k0)
            for i in range(kb):
                src = int(np.where(cur == perm[i])[0][0])
                piv[i] = src - i
                cur[i], cur[src] = cur[src], cur[i]
            loc[k0:, off:off + kb] = panel
            W[:, :] = panel
        tw = torch.from_numpy(np.ascontiguousarray(W)); tp = torch.from_numpy(piv)
        dist.broadcast(tw, src=owner); dist.broadcast(tp, src=owner)
        W = np.asfortranarray(tw.numpy()); piv = tp.numpy()
        trans_all[k0:k0 + kb] = piv
        # apply the swaps to every local column except the panel itself
        mycols = np.ones(loc.shape[1], bool)
        if owner == rank:
            off = lay.local_col_offset(k, nb, world); mycols[off:off + kb] = False
        for i in range(kb):
            a, b = k0 + i, k0 + i + piv[i]
            if a != b:
                tmp = loc[a, mycols].copy(); loc[a, mycols] = loc[b, mycols]; loc[b, mycols] = tmp
        L11 = np.tril(W[:kb, :], -1) + np.eye(kb)
        for j in range(k + 1, nblk):
            if lay.owner_of_block(j, world) != rank:
                continue
            j0 = j * nb; jb = min(nb, n - j0); off = lay.local_col_offset(j, nb, world)
            U = np.asfortranarray(loc[k0:k0 + kb, off:off + jb])
            orc.solve_triangular(np.asfortranarray(L11), U, lower=True, unit=True)
            loc[k0:k0 + kb, off:off + jb] = U
            loc[k0 + kb:, off:off + jb] -= W[kb:, :] @ U
    np.save(os.path.join(out_dir, f"lu_{rank}.npy"), loc)
    np.save(os.path.join(out_dir, f"lu_trans_{rank}.npy"), trans_all)
    dist.destroy_process_group()


@pytest.mark.parametrize("n,nb", [(96, 16), (100, 24)])
def test_block_cyclic_llt_and_lu_schedule_gloo_world2(tmp_path, oracle, n, nb):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, nb, str(tmp_path)), nprocs=world, join=True)
    import faer_b200  # noqa: F401
    from faer_rs_b200 import dist as lay
    # LLT: gather and compare with the single-process oracle
    locs = [np.load(tmp_path / f"llt_{r}.npy") for r in range(world)]
    got = lay.gather_block_cyclic(locs, n, nb, world)
    rng = np.random.default_rng(0)
    G = rng.standard_normal((n, n)); A = np.asfortranarray(G @ G.T + n * np.eye(n))
    want = A.copy(order="F"); fail, _ = oracle.llt(want); assert fail == -1
    assert np.allclose(np.tril(got), np.tril(want), rtol=1e-11, atol=1e-11)
    assert np.array_equal(np.triu(got, 1), np.triu(A, 1))  # strict upper triangle untouched
    # LU: permutation identical to the oracle's, factors within tolerance
    locs = [np.load(tmp_path / f"lu_{r}.npy") for r in range(world)]
    got = lay.gather_block_cyclic(locs, n, nb, world)
    t0 = np.load(tmp_path / "lu_trans_0.npy"); t1 = np.load(tmp_path / "lu_trans_1.npy")
    assert np.array_equal(t0, t1)
    perm = np.arange(n)
    for i, t in enumerate(t0):
        perm[i], perm[i + t] = perm[i + t], perm[i]
    rng = np.random.default_rng(1)
    M = np.asfortranarray(rng.standard_normal((n, n)))
    want = M.copy(order="F"); perm_o, _, _ = oracle.lu(want)
    assert np.array_equal(perm, perm_o)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-10)


def test_layout_helpers():
    import faer_b200  # noqa: F401
    from faer_rs_b200 import dist as lay
    for n, nb, P in [(100, 16, 3), (64, 16, 4), (17, 8, 2), (5, 8, 4)]:
        cols = [lay.global_col_indices(n, nb, P, r) for r in range(P)]
        allc = np.sort(np.concatenate(cols))
        assert np.array_equal(allc, np.arange(n))
        for r in range(P):
            assert len(cols[r]) == lay.local_cols(n, nb, P, r)
            for b in lay.local_blocks(n, nb, P, r):
                assert lay.owner_of_block(b, P) == r
                off = lay.local_col_offset(b, nb, P)
                assert cols[r][off] == b * nb
        A = np.arange(n * n, dtype=np.float64).reshape(n, n)
        locs = [lay.scatter_block_cyclic(A, nb, P, r) for r in range(P)]
        assert np.array_equal(lay.gather_block_cyclic(locs, n, nb, P), A)


def test_column_slabs_partition_the_columns():
    import faer_b200  # noqa: F401
    from faer_rs_b200 import dist as lay
    for ncols in [0, 1, 7, 64, 100, 4097]:
        for P in [1, 2, 3, 4, 8]:
            slabs = [lay.column_slab(ncols, P, r) for r in range(P)]
            assert slabs[0][0] == 0 and slabs[-1][1] == ncols
            assert all(slabs[r][1] == slabs[r + 1][0] for r in range(P - 1))
            widths = [b - a for a, b in slabs]
            assert max(widths) - min(widths) <= 1 and sorted(widths, reverse=True) == widths


def _gemm_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import faer_b200  # noqa: F401
    from faer_rs_b200 import dist as lay
    from faer_rs_b200 import linalg as la
    from oracle import oracle as orc

    # the per-rank product is the library's single-GPU matmul; on the CPU the oracle stands in for it (test infrastructure)
    def cpu_matmul(dst, accum, lhs, rhs, alpha, par=None):
        d = np.asfortranarray(dst.numpy().copy()); orc.matmul(d, accum == la.Accum.Add, np.asfortranarray(lhs.numpy()),
                                                              np.asfortranarray(rhs.numpy()), alpha)
        dst.copy_(torch.from_numpy(d))
    la.matmul = cpu_matmul
    m, k, n = 37, 29, 45
    rng = np.random.default_rng(3)  # replicated inputs, same seed on every rank
    A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n)); C0 = rng.standard_normal((m, n))
    a, b = lay.column_slab(n, world, rank)
    At = torch.from_numpy(A.copy()) if rank == 1 else torch.zeros((m, k), dtype=torch.float64)  # A only valid on rank 1
    Cl = torch.from_numpy(np.ascontiguousarray(C0[:, a:b]))
    lay.matmul(Cl, la.Accum.Add, At, torch.from_numpy(np.ascontiguousarray(B[:, a:b])), 0.5, src_rank=1)
    np.save(os.path.join(out_dir, f"c{rank}.npy"), Cl.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_column_split_gemm_gloo_world2(tmp_path):
    """SURVEY.md 8e GEMM row: 1-D column split, A broadcast from the rank that holds it, no collective in the product; the slabs
    concatenate to the single-process result."""
    world, port = 2, _free_port()
    mp.spawn(_gemm_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    A = rng.standard_normal((37, 29)); B = rng.standard_normal((29, 45)); C0 = rng.standard_normal((37, 45))
    got = np.concatenate([np.load(tmp_path / f"c{r}.npy") for r in range(world)], axis=1)
    assert np.allclose(got, C0 + 0.5 * A @ B, rtol=1e-13, atol=1e-13)


def _qr_worker(rank, world, port, m, n, bs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import faer_b200  # noqa: F401
    from faer_rs_b200 import dist as lay
    from oracle import oracle as orc

    rng = np.random.default_rng(11)  # replicated global input
    A = np.asfortranarray(rng.standard_normal((m, n)))
    loc = lay.scatter_block_cyclic(A, bs, world, rank)
    H = np.zeros((bs, n), order="F")
    # ---- distributed QR schedule (mirrors csrc/dist.cu::dist_qr_impl) ----
    for k in range(lay.num_blocks(n, bs)):
        j0 = k * bs; jb = min(bs, n - j0); owner = lay.owner_of_block(k, world)
        W = np.zeros((m - j0, jb), order="F"); Tk = np.zeros((bs, jb), order="F")
        if owner == rank:
            off = lay.local_col_offset(k, bs, world)
            panel = np.asfortranarray(loc[j0:, off:off + jb])
            Hk, r = orc.qr(panel, block_size=jb)
            assert r == jb
            loc[j0:, off:off + jb] = panel
            W[:, :] = panel; Tk[:jb, :] = Hk
        for buf in (W, Tk):  # the two broadcasts of the step: factored panel, T block
            t = torch.from_numpy(np.ascontiguousarray(buf))
            dist.broadcast(t, src=owner)
            buf[...] = t.numpy()
        H[:, j0:j0 + jb] = Tk
        t0 = sum(min(bs, n - b * bs) for b in range(rank, k + 1, world))  # my columns to the right of block k are the local tail
        if loc.shape[1] > t0:
            M = np.asfortranarray(loc[j0:, t0:])
            orc.apply_block_householder_on_the_left(W, np.asfortranarray(Tk[:jb, :jb]), M, forward=True)
            loc[j0:, t0:] = M
    np.save(os.path.join(out_dir, f"qr{rank}.npy"), loc)
    if rank == 0:
        np.save(os.path.join(out_dir, "qrH.npy"), H)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("m,n,bs,world", [(70, 50, 8, 2), (64, 64, 16, 2), (90, 33, 8, 2), (100, 77, 8, 3), (60, 40, 4, 4)])
def test_block_cyclic_qr_schedule_gloo(tmp_path, oracle, m, n, bs, world):
    """SURVEY.md 8e QR row: owner factors its block column, broadcasts (V panel, T), every rank updates its own columns; the result
    is the single-process blocked QR with the same block size (factors and T blocks to rounding)."""
    import faer_b200  # noqa: F401
    from faer_rs_b200 import dist as lay
    port = _free_port()
    mp.spawn(_qr_worker, args=(world, port, m, n, bs, str(tmp_path)), nprocs=world, join=True)
    got = lay.gather_block_cyclic([np.load(tmp_path / f"qr{r}.npy") for r in range(world)], n, bs, world)
    H = np.load(tmp_path / "qrH.npy")
    rng = np.random.default_rng(11)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    want = A.copy(order="F"); Ho, rank = oracle.qr(want, block_size=bs)
    assert rank == n
    assert np.allclose(got, want, rtol=1e-10, atol=1e-10)
    for j in range(0, n, bs):
        b = min(bs, n - j)
        assert np.allclose(np.triu(H[:b, j:j + b]), np.triu(Ho[:b, j:j + b]), rtol=1e-10, atol=1e-10)
