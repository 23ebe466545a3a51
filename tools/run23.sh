mkdir -p gpurun_out
timeout 120 tools/tc_f32_test 2>&1 | tee gpurun_out/tc_f32_test.log
timeout 200 python tools/time_other.py qr 2>&1 | tee gpurun_out/time_qr.log
timeout 300 python -m pytest tests/test_gpu_qr.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_qr.log
