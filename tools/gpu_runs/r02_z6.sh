#!/bin/bash
# round 2, call Z6 (2 GPUs): timing of the distributed QR on the configs[3] shape (f32 65536 x 4096) at world 2 vs one GPU
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29577 tools/time_dist_qr.py 2>&1 | grep "time_dist_qr\|Error\|error\|Traceback" | tail -12 | tee gpurun_out/r02_z6_time_dist_qr.log
