#!/bin/bash
# round 2, call Z4 (2 GPUs): distributed QR + column-split GEMM: single-rank tests, then torchrun world 2 through the parity tool
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zzzzzzz_dist_gemm_qr.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_z4_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29555 tools/dist_parity.py 3072 256 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/r02_z4_parity_w2.log
