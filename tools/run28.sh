mkdir -p gpurun_out
timeout 120 tools/tc_f32_test 2>&1 | tail -4 | tee gpurun_out/tc_f32_test.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 100 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee gpurun_out/llt_bulk.log
FAER_B200_LLT_SPLIT_BULK=1 timeout 100 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee -a gpurun_out/llt_bulk.log
timeout 100 python tools/time_other.py qr f32 2>&1 | tee gpurun_out/time_qr32.log
timeout 100 python tools/time_other.py gemm 2>&1 | grep f32 | tee -a gpurun_out/time_qr32.log
