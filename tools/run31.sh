mkdir -p gpurun_out
rm -f gpurun_out/nb_sweep3.log
for NB in 256 512; do
  echo "--- LLT NB=$NB" | tee -a gpurun_out/nb_sweep3.log
  FAER_B200_NB=$NB timeout 100 python tools/time_factor.py llt 8192 16384 2>&1 | tail -2 | tee -a gpurun_out/nb_sweep3.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_part.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.log
