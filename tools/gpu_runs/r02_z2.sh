#!/bin/bash
# round 2, call Z2: complex QR / Householder sequences / QR solves (cplx.cu) + the c32 file after the robust complex reciprocal
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzzzz_cplx_qr.py tests/test_gpu_zzzz_c32_llt_trsm_lu.py tests/test_gpu_zzz_c64_llt_trsm_lu.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r02_z2_tests.log
