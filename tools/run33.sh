mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_host_pipeline.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_part.log
timeout 200 python tools/time_e2e.py 16384 2>&1 | tail -5 | tee gpurun_out/e2e.log
FAER_B200_HOST_LEFT=0 timeout 200 python tools/time_e2e.py 16384 2>&1 | tail -5 | tee -a gpurun_out/e2e.log
