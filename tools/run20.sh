mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_qr.py tests/test_gpu_condensed.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_part.log
timeout 200 python tools/time_other.py qr 2>&1 | tee gpurun_out/time_qr.log
timeout 100 python tools/time_condensed.py 4096 8192 2>&1 | grep f64 | tee gpurun_out/time_condensed.log
rm -f gpurun_out/green_lu.log
echo "--- plain" | tee -a gpurun_out/green_lu.log
timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -1 | tee -a gpurun_out/green_lu.log
echo "--- GREEN_SMS=16 offload" | tee -a gpurun_out/green_lu.log
FAER_B200_GREEN_SMS=16 FAER_B200_TRACE=1 timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -34 | tee -a gpurun_out/green_lu.log
echo "--- GREEN_SMS=16 no offload" | tee -a gpurun_out/green_lu.log
FAER_B200_NO_OFFLOAD=1 FAER_B200_GREEN_SMS=16 timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -1 | tee -a gpurun_out/green_lu.log
echo "--- GREEN_SMS=8 offload" | tee -a gpurun_out/green_lu.log
FAER_B200_GREEN_SMS=8 timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -1 | tee -a gpurun_out/green_lu.log
echo "--- GREEN_SMS=16 offload n=32768" | tee -a gpurun_out/green_lu.log
FAER_B200_GREEN_SMS=16 timeout 300 python tools/time_factor.py lu 32768 2>&1 | tail -1 | tee -a gpurun_out/green_lu.log
