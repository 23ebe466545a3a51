mkdir -p gpurun_out
timeout 120 tools/tc_f32_test 2>&1 | tail -5 | tee gpurun_out/tc_f32_test.log
timeout 100 python tools/time_other.py qr f32 2>&1 | tee gpurun_out/time_qr32.log
timeout 200 python -m pytest tests/test_gpu_f32.py tests/test_gpu_qr.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_part.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/qr32_launches.csv python tools/time_other.py qr f32 > gpurun_out/qr32_ncu.log 2>&1
