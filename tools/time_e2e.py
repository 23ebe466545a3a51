"""Dev aid: LLT through the C ABI with a pinned HOST matrix (pipelined transfers) vs device-resident; checks they agree."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_b200  # noqa: E402
from faer_b200 import linalg as la  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda:0")
lib = faer_b200.load()
lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
G = torch.randn((n, n), dtype=torch.float64, device=dev)
A0 = torch.addmm(n * torch.eye(n, dtype=torch.float64, device=dev), G, G.T).T
del G
A = A0.clone(memory_format=torch.preserve_format)
la.cholesky_in_place(A)
torch.cuda.synchronize()
hA0 = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
hA0.copy_(A0.T)  # storage of the column-major matrix
hA = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
for it in range(3):
    hA.copy_(hA0)
    t0 = time.time()
    la.cholesky_in_place(hA.T)
    t1 = time.time()
    print(f"e2e host LLT n={n}: {(t1 - t0) * 1e3:.1f} ms  {n**3 / 3 / (t1 - t0) / 1e12:.2f} TFLOP/s", flush=True)
L_dev = torch.tril(A).cpu()
L_host = torch.tril(hA.T)
print("max |L_host - L_dev| =", float((L_host - L_dev).abs().max()), " upper untouched:",
      bool(torch.equal(torch.triu(hA.T, 1), torch.triu(hA0.T, 1))))
