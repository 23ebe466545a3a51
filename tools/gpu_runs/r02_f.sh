# Round 2, call F: ws GEMM disagreement located (tile pattern / ratio), the f32 rank-deficient QR failures with tracebacks, the
# tests after the abort of call E, and the fused LU sub-panel kernel (parity with the recursive driver first, then timing).
mkdir -p gpurun_out
timeout 300 python tools/debug_ws_syrk.py > gpurun_out/r02_f_debug_ws.log 2>&1; tail -60 gpurun_out/r02_f_debug_ws.log
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 300 $PYT "tests/test_gpu_qr.py" -k "rank_deficient and float32" > gpurun_out/r02_f_qr32.log 2>&1; tail -60 gpurun_out/r02_f_qr32.log
export FAER_B200_GEMM_WS=0
timeout 900 $PYT tests/test_gpu_zz11_evd_svd_vectors.py tests/test_gpu_zz12_spicy_matmul.py tests/test_gpu_zz1_new_entry_points.py tests/test_gpu_zz2_qr_solve.py tests/test_gpu_zz3_solvers.py tests/test_gpu_zz4_reconstruct_inverse.py tests/test_gpu_zz5_llt_f32.py tests/test_gpu_zz6_ldlt.py tests/test_gpu_zz7_singular_values.py tests/test_gpu_zz8_self_adjoint_eigenvalues.py tests/test_gpu_zz9_baseline_sizes.py tests/test_gpu_parity.py > gpurun_out/r02_f_rest.log 2>&1; tail -40 gpurun_out/r02_f_rest.log
echo "--- fused LU sub-panel: parity on the recursive driver (cluster 16)"
FAER_B200_LOOKAHEAD_MIN_N=0 FAER_B200_LU_CLUSTER=16 timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py -k "plu or lu or pivot" > gpurun_out/r02_f_lu_fused_tests.log 2>&1; tail -15 gpurun_out/r02_f_lu_fused_tests.log
echo "--- partitioned driver (default path) parity"
timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_solve.py -k "plu or lu or pivot" > gpurun_out/r02_f_lu_part_tests.log 2>&1; tail -8 gpurun_out/r02_f_lu_part_tests.log
for W in 0 128 256; do
  echo "--- LU timing, FAER_B200_LU_FUSED_W=$W (gemm_ws=0)"
  FAER_B200_LU_FUSED_W=$W timeout 200 python tools/time_factor.py lu 8192 16384 2>&1 | tail -2
done
FAER_B200_TRACE=1 timeout 200 python tools/time_factor.py lu 16384 > gpurun_out/r02_f_lu_trace.log 2>&1; tail -1 gpurun_out/r02_f_lu_trace.log
FAER_B200_TRACE=1 timeout 300 python tools/time_factor.py lu 32768 > gpurun_out/r02_f_lu_trace32.log 2>&1; tail -1 gpurun_out/r02_f_lu_trace32.log
FAER_B200_DIST_NO_PARTITION=1 timeout 300 python tools/time_factor.py lu 32768 2>&1 | tail -1
