"""Multi-GPU check + timing of the distributed LLT (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [n] [nb]
Every rank builds the same global SPD matrix (same seed), keeps its block columns, factors, and rank 0 gathers and
checks  A = L L^T  on random probes; then times a device-resident run (max over ranks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    faer_b200.dist.init_from_torch_distributed()
lay = faer_b200.dist

torch.manual_seed(0)
G = torch.randn((n, n), dtype=torch.float64, device=dev)
A = torch.addmm(n * torch.eye(n, dtype=torch.float64, device=dev), G, G.T)
del G
cols = torch.as_tensor(lay.global_col_indices(n, nb, world, rank), device=dev)
loc0 = A[:, cols].T.contiguous().T  # column-major n x local_cols
loc = loc0.clone(memory_format=torch.preserve_format)
fail, cnt = lay.cholesky_in_place(loc, n, nb=nb)
assert fail == -1, fail
# gather L on rank 0
if world > 1:
    parts = [torch.empty((n, lay.local_cols(n, nb, world, r)), dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(parts, loc.contiguous()) if len({p.shape for p in parts}) == 1 else None
    if len({p.shape for p in parts}) != 1:
        parts = None
else:
    parts = [loc]
if rank == 0 and parts is not None:
    Lfull = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for r in range(world):
        Lfull[:, torch.as_tensor(lay.global_col_indices(n, nb, world, r), device=dev)] = parts[r]
    L = torch.tril(Lfull)
    x = torch.randn((n, 4), dtype=torch.float64, device=dev)
    resid = float((A @ x - L @ (L.T @ x)).abs().max()) / (float(A.abs().max()) * n)
    up_ok = bool(torch.equal(torch.triu(Lfull, 1), torch.triu(A, 1)))
    print(f"[dist_check] world={world} n={n} nb={nb}: probe residual {resid:.3e}, upper untouched {up_ok}", flush=True)
    assert resid < 1e-13 and up_ok

# timing
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

for la in (1, 0):
    best = 1e30
    for it in range(4):
        loc.copy_(loc0)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        lay.cholesky_in_place(loc, n, nb=nb, lookahead=bool(la))
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            best = min(best, float(t.item()))
    if rank == 0:
        print(f"[dist_check] world={world} n={n} nb={nb} lookahead={la}: {best:.2f} ms  {n**3/3/best/1e9:.2f} TFLOP/s aggregate", flush=True)
# ---- distributed LU: P A = L U on probes, permutation identical on all ranks, timing ----
del A, loc, loc0
torch.manual_seed(1)
M = torch.randn((n, n), dtype=torch.float64, device=dev)
loc0 = M[:, cols].T.contiguous().T
loc = loc0.clone(memory_format=torch.preserve_format)
perm, pinv, nt = lay.lu_in_place(loc, n, nb=nb)
tperm = torch.as_tensor(perm, device=dev)
if world > 1:
    t0 = tperm.clone(); dist.broadcast(t0, src=0)
    assert torch.equal(t0, tperm), "permutation differs across ranks"
    parts = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(parts, loc.contiguous())
else:
    parts = [loc]
if rank == 0:
    LU = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for r in range(world):
        LU[:, torch.as_tensor(lay.global_col_indices(n, nb, world, r), device=dev)] = parts[r]
    x = torch.randn((n, 4), dtype=torch.float64, device=dev)
    Ux = torch.triu(LU) @ x
    r_ = M[tperm, :] @ x - (torch.tril(LU, -1) @ Ux + Ux)
    growth = max(1.0, float(torch.triu(LU).abs().max()) / float(M.abs().max()))
    resid = float(r_.abs().max()) / (float(M.abs().max()) * n * growth)
    print(f"[dist_check] LU world={world} n={n} nb={nb}: probe residual {resid:.3e}, max|L| {float(torch.tril(LU, -1).abs().max()):.3f}, "
          f"transpositions {nt}", flush=True)
    assert resid < 1e-12
    del LU
for la in (1, 0):
    best = 1e30
    for it in range(3):
        loc.copy_(loc0)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        lay.lu_in_place(loc, n, nb=nb, lookahead=bool(la))
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            best = min(best, float(t.item()))
    if rank == 0:
        print(f"[dist_check] LU world={world} n={n} nb={nb} lookahead={la}: {best:.2f} ms  {2*n**3/3/best/1e9:.2f} TFLOP/s aggregate", flush=True)
if world > 1:
    faer_b200.dist.finalize()
    dist.destroy_process_group()
