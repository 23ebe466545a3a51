"""Timing of the distributed Householder QR (csrc/dist.cu::dist_qr_impl, no look-ahead) on the BASELINE.json configs[3] shape,
f32 65536 x 4096, block size 256: P ranks (torchrun, one per GPU) against the same driver run locally and the single-GPU entry
point on rank 0. Device-resident, restore copy + factor per repetition, barrier + CUDA events, max over ranks.
usage: torchrun --nproc-per-node P tools/time_dist_qr.py [m] [n]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    faer_b200.dist.init_from_torch_distributed()
lay, la = faer_b200.dist, faer_b200.linalg
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
bs = int(la.qr_recommended_block_size(m, n))
flops = 2.0 * m * n * n - 2.0 * n ** 3 / 3.0

torch.manual_seed(0)
cols = torch.as_tensor(lay.global_col_indices(n, bs, world, rank), device=dev)
# the same global matrix on every rank, generated block column by block column to keep the footprint at the local size
gen = torch.Generator(device=dev); gen.manual_seed(1234)
A0 = torch.empty((len(cols), m), dtype=torch.float32, device=dev).T  # column-major m x local_cols
for b in range((n + bs - 1) // bs):
    blk = torch.randn((min(bs, n - b * bs), m), dtype=torch.float32, device=dev, generator=gen)
    if b % world == rank:
        off = lay.local_col_offset(b, bs, world)
        A0.T[off:off + blk.shape[0]].copy_(blk)
loc = A0.clone(memory_format=torch.preserve_format)


def timed(f, reps=3):
    best = 1e30
    for _ in range(reps + 1):  # first repetition = warm-up
        loc.copy_(A0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if _ > 0:
            best = min(best, float(t.item()))
    return best


ms = timed(lambda: lay.qr_in_place(loc, m, n, bs))
if rank == 0:
    print(f"[time_dist_qr] f32 {m}x{n} bs={bs} world={world}: {ms:.2f} ms = {flops / ms / 1e9:.2f} TFLOP/s (max over ranks, best of 3)", flush=True)
if world > 1:
    dist.barrier()
if rank == 0 and world > 1:
    # the whole matrix on one GPU: the distributed driver run locally, and the single-GPU entry point
    del loc, A0
    gen.manual_seed(1234)
    F0 = torch.empty((n, m), dtype=torch.float32, device=dev).T
    for b in range((n + bs - 1) // bs):
        F0.T[b * bs:b * bs + min(bs, n - b * bs)].copy_(torch.randn((min(bs, n - b * bs), m), dtype=torch.float32, device=dev, generator=gen))
    F = F0.clone(memory_format=torch.preserve_format)
    H = torch.zeros((n, bs), dtype=torch.float32, device=dev).T

    def t1(f):
        best = 1e30
        for i in range(4):
            F.copy_(F0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            if i > 0:
                best = min(best, e0.elapsed_time(e1))
        return best
    a = t1(lambda: lay.qr_in_place(F, m, n, bs, local_only=True))
    b_ = t1(lambda: la.qr_in_place(F, H))
    print(f"[time_dist_qr] on one GPU: distributed driver run locally {a:.2f} ms = {flops / a / 1e9:.2f} TFLOP/s; "
          f"libfaer_v0_23_qr_factor_in_place_f32 {b_:.2f} ms = {flops / b_ / 1e9:.2f} TFLOP/s", flush=True)
if world > 1:
    dist.barrier()
    faer_b200.dist.finalize()
    dist.destroy_process_group()
