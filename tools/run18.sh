mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_condensed.py -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_condensed.log
timeout 100 python tools/time_condensed.py 4096 8192 2>&1 | grep f64 | tee gpurun_out/time_condensed.log
for P in 16 24 32; do
  echo "--- GREEN_SMS=$P" | tee -a gpurun_out/green_lu.log
  FAER_B200_GREEN_SMS=$P FAER_B200_TRACE=1 timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -5 | tee -a gpurun_out/green_lu.log
done
echo "--- GREEN_SMS=24 NB=256" | tee -a gpurun_out/green_lu.log
FAER_B200_NB=256 FAER_B200_GREEN_SMS=24 timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -1 | tee -a gpurun_out/green_lu.log
echo "--- GREEN_SMS=24 n=32768" | tee -a gpurun_out/green_lu.log
FAER_B200_GREEN_SMS=24 timeout 300 python tools/time_factor.py lu 32768 2>&1 | tail -1 | tee -a gpurun_out/green_lu.log
echo "--- plain" | tee -a gpurun_out/green_lu.log
timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -1 | tee -a gpurun_out/green_lu.log
