"""Multi-rank parity check (run under torchrun, one rank per GPU): the distributed LLT and LU on P ranks against the SAME
factorizations on one GPU (rank 0, communicator ignored): permutations and status bit-exact, factors to rounding (different GEMM
kernels may serve different local widths), plus the reconstruction probes. Prints one line per check and exits non-zero on failure.
With `qr-gemm` as third argument the run checks the distributed QR and the column-split GEMM instead (same exit convention).
usage: torchrun --nproc-per-node P tools/dist_parity.py [n] [nb] [qr-gemm]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
what = sys.argv[3] if len(sys.argv) > 3 else "llt-lu"
assert what in ("llt-lu", "qr-gemm", "all")
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    faer_b200.dist.init_from_torch_distributed()
lay = faer_b200.dist
ok = True


def gather_cols(loc):
    """global n x n matrix on rank 0 from the block-column-cyclic pieces (None elsewhere)"""
    if world == 1:
        return loc
    parts = [torch.empty((n, lay.local_cols(n, nb, world, r)), dtype=torch.float64, device=dev) for r in range(world)] if rank == 0 else None
    # pieces may differ in width: send one by one
    full = torch.zeros((n, n), dtype=torch.float64, device=dev) if rank == 0 else None
    for r in range(world):
        w = lay.local_cols(n, nb, world, r)
        buf = loc.contiguous() if r == rank else torch.empty((n, w), dtype=torch.float64, device=dev)
        dist.broadcast(buf, src=r)
        if rank == 0:
            full[:, torch.as_tensor(lay.global_col_indices(n, nb, world, r), device=dev)] = buf
    return full


torch.manual_seed(0)
G = torch.randn((n, n), dtype=torch.float64, device=dev)
S = torch.addmm(n * torch.eye(n, dtype=torch.float64, device=dev), G, G.T)
cols = torch.as_tensor(lay.global_col_indices(n, nb, world, rank), device=dev)
for name in (("llt", "lu") if what in ("llt-lu", "all") else ()):
    A = S if name == "llt" else G
    loc = A[:, cols].T.contiguous().T
    if name == "llt":
        fail, cnt = lay.cholesky_in_place(loc, n, nb=nb)
        perm = None
    else:
        perm, pinv, nt = lay.lu_in_place(loc, n, nb=nb)
    full = gather_cols(loc)
    if rank == 0:
        one = A.T.contiguous().T.clone(memory_format=torch.preserve_format)
        if name == "llt":
            # single-rank run with the communicator ignored (lookahead bit 1)
            import ctypes as C
            from faer_b200 import capi
            lib = capi.load()
            d0 = C.c_double(0.0); e0 = C.c_double(0.0)
            reg = capi.LltRegularization(C.cast(C.pointer(d0), C.c_void_p), C.cast(C.pointer(e0), C.c_void_p))
            st = lib.faer_b200_dist_llt_factor_in_place_f64(one.data_ptr(), n, n, nb, reg, 3)
            same_status = (st.tag == 0) == (fail == -1)
            a = torch.tril(full); b = torch.tril(one)
        else:
            p1, _, nt1 = lay.lu_in_place(one, n, nb=nb, lookahead=3)
            same_status = bool(np.array_equal(p1, perm)) and nt1 == nt
            a = full; b = one
        diff = float((a - b).abs().max()); scale = float(b.abs().max())
        x = torch.randn((n, 3), dtype=torch.float64, device=dev)
        if name == "llt":
            L = torch.tril(full)
            resid = float((S @ x - L @ (L.T @ x)).abs().max()) / (float(S.abs().max()) * n)
        else:
            L = torch.tril(full, -1) + torch.eye(n, dtype=torch.float64, device=dev); U = torch.triu(full)
            resid = float((G[torch.as_tensor(perm, device=dev)] @ x - L @ (U @ x)).abs().max()) / (float(G.abs().max()) * n)
        good = same_status and diff <= 1e-9 * scale and resid < 1e-12
        ok = ok and good
        print(f"[dist_parity] {name} world={world} n={n} nb={nb}: status/perm identical {same_status}, "
              f"max |factor(P) - factor(1)| = {diff:.3e} (scale {scale:.1f}), probe residual {resid:.3e} -> {'OK' if good else 'FAIL'}", flush=True)
if what in ("qr-gemm", "all"):
    # ---- distributed QR (broadcast (V panel, T)): P ranks against the same driver run locally on rank 0 and against the
    # reconstruction probe A x = Q (R x) through the library's own block-Householder sequence ----
    mq, nq = n + n // 2, n
    bsq = int(faer_b200.linalg.qr_recommended_block_size(mq, nq))
    torch.manual_seed(2)
    Aq = torch.randn((mq, nq), dtype=torch.float64, device=dev)
    colsq = torch.as_tensor(lay.global_col_indices(nq, bsq, world, rank), device=dev)
    locq = Aq[:, colsq].T.contiguous().T
    Hq = lay.qr_in_place(locq, mq, nq, bsq)
    if world > 1:
        fullq = torch.zeros((mq, nq), dtype=torch.float64, device=dev) if rank == 0 else None
        for r in range(world):
            w = lay.local_cols(nq, bsq, world, r)
            buf = locq.T.contiguous() if r == rank else torch.empty((w, mq), dtype=torch.float64, device=dev)
            dist.broadcast(buf, src=r)
            if rank == 0:
                fullq[:, torch.as_tensor(lay.global_col_indices(nq, bsq, world, r), device=dev)] = buf.T
    else:
        fullq = locq
    if rank == 0:
        oneq = Aq.T.contiguous().T.clone(memory_format=torch.preserve_format)
        H1 = lay.qr_in_place(oneq, mq, nq, bsq, local_only=True)
        diff = float((fullq - oneq).abs().max()); scale = float(oneq.abs().max())
        hdiff = max(float((torch.triu(Hq[:min(bsq, nq - j), j:j + bsq]) - torch.triu(H1[:min(bsq, nq - j), j:j + bsq])).abs().max())
                    for j in range(0, nq, bsq))
        x = torch.randn((nq, 3), dtype=torch.float64, device=dev)
        y = torch.zeros((mq, 3), dtype=torch.float64, device=dev).T.contiguous().T
        y[:nq] = torch.triu(fullq[:nq]) @ x
        yc = y.T.contiguous().T
        faer_b200.linalg.apply_block_householder_sequence_on_the_left_in_place(fullq.T.contiguous().T, Hq, yc)
        resid = float((Aq @ x - yc).abs().max()) / (float(Aq.abs().max()) * mq)
        good = diff <= 1e-9 * scale and hdiff <= 1e-9 and resid < 1e-12
        ok = ok and good
        print(f"[dist_parity] qr world={world} {mq}x{nq} bs={bsq}: max |factor(P) - factor(1)| = {diff:.3e} (scale {scale:.1f}), "
              f"max |T(P) - T(1)| = {hdiff:.3e}, probe residual {resid:.3e} -> {'OK' if good else 'FAIL'}", flush=True)

    # ---- column-split GEMM: every rank forms its slab of C = C0 + alpha A B with A broadcast from the last rank; the
    # slabs are compared with the same product done on one GPU (rank 0). Different local widths may pick different kernels, so the
    # comparison is to the forward bound 2 k u sum|a||b|, not bitwise ----
    m2 = k2 = n
    torch.manual_seed(1)
    Af = torch.randn((m2, k2), dtype=torch.float64, device=dev); Bf = torch.randn((k2, n), dtype=torch.float64, device=dev)
    C0 = torch.randn((m2, n), dtype=torch.float64, device=dev)
    a0, a1 = lay.column_slab(n, world, rank)
    A_in = Af.T.contiguous().T if rank == world - 1 else torch.zeros((k2, m2), dtype=torch.float64, device=dev).T
    Cl = C0[:, a0:a1].T.contiguous().T
    lay.matmul(Cl, faer_b200.linalg.Accum.Add, A_in, Bf[:, a0:a1].T.contiguous().T, 0.75, src_rank=world - 1)
    if world > 1:
        pieces = []
        for r in range(world):
            b0, b1 = lay.column_slab(n, world, r)
            buf = Cl.T.contiguous() if r == rank else torch.empty((b1 - b0, m2), dtype=torch.float64, device=dev)
            dist.broadcast(buf, src=r)
            pieces.append(buf.T)
        Cfull = torch.cat(pieces, dim=1)
    else:
        Cfull = Cl
    if rank == 0:
        one = C0.T.contiguous().T.clone(memory_format=torch.preserve_format)
        faer_b200.linalg.matmul(one, faer_b200.linalg.Accum.Add, Af.T.contiguous().T, Bf.T.contiguous().T, 0.75)
        bound = 2 * k2 * 2.0 ** -53 * 0.75 * (Af.abs() @ Bf.abs()) + 4 * 2.0 ** -53 * C0.abs()
        worst = float(((Cfull - one).abs() / bound).max())
        good = worst <= 1.0
        ok = ok and good
        print(f"[dist_parity] gemm world={world} n={n}: max |C(P) - C(1)| / (2 k u |A||B|) = {worst:.3e} -> {'OK' if good else 'FAIL'}", flush=True)
if world > 1:
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(t, src=0)
    ok = bool(t.item())
    faer_b200.dist.finalize()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
