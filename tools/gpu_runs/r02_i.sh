# Round 2, call I: register-resident fused LU sub-panel (parity, phases, timing), ws stagger, the inner-seam test.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
echo "--- fused LU sub-panel v2: recursive driver (cluster 16)"
FAER_B200_LOOKAHEAD_MIN_N=0 FAER_B200_LU_CLUSTER=16 timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py -k "plu or lu or pivot" 2>&1 | tail -8 | tee gpurun_out/r02_i_lu_rec_tests.log
echo "--- partitioned driver (default path)"
timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_solve.py tests/test_gpu_zz9_baseline_sizes.py -k "plu or lu or pivot" 2>&1 | tail -8 | tee gpurun_out/r02_i_lu_part_tests.log
timeout 200 $PYT tests/test_gpu_zz12_spicy_matmul.py 2>&1 | tail -8 | tee gpurun_out/r02_i_seam.log
for W in 128 256; do
  echo "--- LU timing FUSED_W=$W" | tee -a gpurun_out/r02_i_lu_time.log
  FAER_B200_LU_FUSED_W=$W FAER_B200_LU_SUBPANEL_PROF=1 timeout 200 python tools/time_factor.py lu 8192 16384 2>&1 | tail -3 | tee -a gpurun_out/r02_i_lu_time.log
done
FAER_B200_TRACE=1 timeout 200 python tools/time_factor.py lu 16384 > gpurun_out/r02_i_lu_trace.log 2>&1; tail -1 gpurun_out/r02_i_lu_trace.log
timeout 300 python tools/time_factor.py lu 32768 2>&1 | tail -1 | tee -a gpurun_out/r02_i_lu_time.log
for S in 0 1; do FAER_B200_WS_STAGGER=$S timeout 200 python tools/time_ws_shapes.py 2>&1 | tee -a gpurun_out/r02_i_ws_stagger.log; done
timeout 200 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee gpurun_out/r02_i_llt.log
FAER_B200_GEMM_WS=0 timeout 200 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee -a gpurun_out/r02_i_llt.log
