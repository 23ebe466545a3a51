# Round 2, call T: LLT with paired (k = 512) trailing updates: parity incl. host pipeline bit-identity, timing with / without, e2e, bench.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py tests/test_gpu_dist.py tests/test_gpu_solve.py tests/test_gpu_zz3_solvers.py tests/test_gpu_zz4_reconstruct_inverse.py tests/test_gpu_zz5_llt_f32.py -k "llt or LLT or cholesky or Llt or solvers" 2>&1 | tail -6 | tee gpurun_out/r02_t_llt_tests.log
for NP in 0 1; do
  echo "--- FAER_B200_LLT_NO_PAIR=$NP" | tee -a gpurun_out/r02_t_llt_time.log
  if [ $NP = 1 ]; then export FAER_B200_LLT_NO_PAIR=1; fi
  timeout 200 python tools/time_factor.py llt 8192 16384 2>&1 | tail -2 | tee -a gpurun_out/r02_t_llt_time.log
done
unset FAER_B200_LLT_NO_PAIR
timeout 300 python tools/time_e2e.py 16384 2>&1 | tail -3 | tee gpurun_out/r02_t_e2e.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_t_bench.log 2>&1; tail -1 gpurun_out/r02_t_bench.log | cut -c1-330
