# Round 2, call O: LLT with 128-wide blocks in the chain-bound tail (parity incl. the host pipeline, timing on / off), bench line,
# ncu capture of the LLT's dominant launch (first trailing update on the cp.async kernel) for roofline.traffic.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py tests/test_gpu_dist.py tests/test_gpu_solve.py tests/test_gpu_zz3_solvers.py tests/test_gpu_zz4_reconstruct_inverse.py -k "llt or LLT or cholesky or Llt or solvers" 2>&1 | tail -6 | tee gpurun_out/r02_o_llt_tests.log
for T in 8192 0 6144 10240; do
  echo "--- LLT tail=$T" | tee -a gpurun_out/r02_o_llt_time.log
  FAER_B200_LLT_TAIL=$T timeout 200 python tools/time_factor.py llt 8192 16384 2>&1 | tail -2 | tee -a gpurun_out/r02_o_llt_time.log
done
timeout 300 python tools/time_e2e.py 16384 2>&1 | tail -3 | tee gpurun_out/r02_o_e2e.log
timeout 600 python bench.py > gpurun_out/r02_o_bench.log 2>&1; tail -1 gpurun_out/r02_o_bench.log | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_llt_update_cp -f python tools/run_syrk_once.py 0 16128 256 > gpurun_out/r02_o_ncu.log 2>&1; tail -2 gpurun_out/r02_o_ncu.log
