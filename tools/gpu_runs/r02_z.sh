#!/bin/bash
# round 2, call Z: c32 twins of the complex TRSM / LLT / LU (cplx.cu templated over the real type) + the c64 file on the same source
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zzzz_c32_llt_trsm_lu.py tests/test_gpu_zzz_c64_llt_trsm_lu.py tests/test_gpu_zz13_lu_f32.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_z_tests.log
