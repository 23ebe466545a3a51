mkdir -p gpurun_out
timeout 120 tools/tc_f32_test 2>&1 | tee gpurun_out/tc_f32_test.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python tools/time_other.py all 2>&1 | tee gpurun_out/time_other.log
