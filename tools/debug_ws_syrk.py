"""Dev aid: locate disagreements between the cp.async and the TMA / warp-specialised f64 GEMM (lower / rectangular destination,
Add / Replace) against a torch reference; for bad entries reports the tile pattern and the ratio (C - ref) / (A B), which
tells missed (+1 for alpha = -1), doubled (-1) or partial contributions apart. usage: python tools/debug_ws_syrk.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(7)
cases = [(4096, 256, "low", 1), (6016, 256, "low", 1), (6016, 256, "low", 0), (6016, 256, "rect", 1), (6016, 256, "rect", 0),
         (6016, 64, "low", 1), (8192, 512, "low", 1), (8192, 512, "rectB", 1), (16128, 256, "low", 1)]
for (n, k, kind, add) in cases:
    A = torch.randn((k, n), dtype=torch.float64, device=dev).T  # n x k column-major
    B = A.T if kind != "rectB" else torch.randn((n, k), dtype=torch.float64, device=dev).T  # rectB: k x n column-major
    C0 = torch.randn((n, n), dtype=torch.float64, device=dev).T
    P = A @ B
    full = (C0 - P) if add else (-P)
    ref = (torch.tril(full) + torch.triu(C0, 1)) if kind == "low" else full
    for mode in (0, 2):
        lib.faer_b200_set_option(b"gemm_ws", mode)
        for rep in range(2):
            Cm = C0.clone(memory_format=torch.preserve_format)
            if kind == "low":
                la.matmul_triangular(Cm, 1, add, A, 0, B, 0, -1.0)
            else:
                la.matmul(Cm, add, A, B, -1.0)
            torch.cuda.synchronize()
            d = (Cm - ref).abs()
            bad = d > 1e-9
            nbad = int(bad.sum())
            msg = f"n={n} k={k} {kind} add={add} mode={mode} rep={rep}: max|C - ref| = {float(d.max()):.3e}, bad = {nbad}"
            if nbad:
                idx = bad.nonzero()
                r0, r1 = int(idx[:, 0].min()), int(idx[:, 0].max()); c0, c1 = int(idx[:, 1].min()), int(idx[:, 1].max())
                sub = idx[:: max(1, nbad // 4000)]
                tiles = sorted({(int(i) // 128, int(j) // 64) for i, j in sub.tolist()})
                ratio = ((Cm - ref)[bad] / P[bad])
                msg += (f" rows {r0}..{r1} cols {c0}..{c1} ntiles(128x64)~{len(tiles)} first {tiles[:10]} last {tiles[-4:]}"
                        f" ratio (C-ref)/(AB): median {float(ratio.median()):.3f} min {float(ratio.min()):.3f} max {float(ratio.max()):.3f}")
                # per-tile bad counts for the first few tiles
                for (ti, tj) in tiles[:3]:
                    blk = bad[ti * 128:(ti + 1) * 128, tj * 64:(tj + 1) * 64]
                    rr = blk.any(dim=1).nonzero().flatten().tolist(); cc = blk.any(dim=0).nonzero().flatten().tolist()
                    msg += f"\n    tile ({ti},{tj}): {int(blk.sum())} bad, rows {rr[:3]}..{rr[-3:]} cols {cc[:3]}..{cc[-3:]}"
            print(msg, flush=True)
    del A, B, C0, P, full, ref
lib.faer_b200_set_option(b"gemm_ws", 1)
