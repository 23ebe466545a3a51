"""Dev aid: locate disagreements between the cp.async and the TMA / warp-specialised f64 GEMM on the LLT trailing update
(lower destination, Add) and compare both with a torch reference. usage: python tools/debug_ws_syrk.py"""
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
for (n, k) in [(2048, 256), (4096, 256), (6016, 256), (6016, 64), (8192, 512)]:
    A = torch.randn((k, n), dtype=torch.float64, device=dev).T
    C0 = torch.randn((n, n), dtype=torch.float64, device=dev).T
    ref = torch.tril(C0 - A @ A.T) + torch.triu(C0, 1)
    outs = {}
    for mode in (0, 2):
        lib.faer_b200_set_option(b"gemm_ws", mode)
        for rep in range(3):
            Cm = C0.clone(memory_format=torch.preserve_format)
            la.matmul_triangular(Cm, 1, la.Accum.Add, A, 0, A.T, 0, -1.0)
            torch.cuda.synchronize()
            d = (Cm - ref).abs()
            bad = d > 1e-9
            nb = int(bad.sum())
            msg = f"n={n} k={k} mode={mode} rep={rep}: max|C - ref| = {float(d.max()):.3e}, bad = {nb}"
            if nb:
                idx = bad.nonzero()
                r0, r1 = int(idx[:, 0].min()), int(idx[:, 0].max()); c0, c1 = int(idx[:, 1].min()), int(idx[:, 1].max())
                tiles = sorted({(int(i) // 128, int(j) // 64) for i, j in idx[:: max(1, nb // 2000)].tolist()})
                msg += f" rows {r0}..{r1} cols {c0}..{c1} tiles(128x64) {tiles[:12]}{'...' if len(tiles) > 12 else ''}"
            print(msg, flush=True)
        outs[mode] = Cm
lib.faer_b200_set_option(b"gemm_ws", 1)
