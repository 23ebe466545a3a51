# Round 2, call L: fused LU sub-panel v3 (redux arg-max, owner-warp push, in-kernel swap plans): parity, panel chain, LU timings.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
FAER_B200_LOOKAHEAD_MIN_N=0 FAER_B200_LU_CLUSTER=16 timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py -k "plu or lu or pivot" 2>&1 | tail -8 | tee gpurun_out/r02_l_lu_rec_tests.log
timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_solve.py tests/test_gpu_zz9_baseline_sizes.py tests/test_gpu_zz4_reconstruct_inverse.py -k "plu or lu or pivot" 2>&1 | tail -8 | tee gpurun_out/r02_l_lu_part_tests.log
FAER_B200_LU_CLUSTER=16 FAER_B200_LU_SUBPANEL_PROF=1 timeout 120 python tools/time_lu_panel.py 128 2>&1 | tee gpurun_out/r02_l_panel128.log
FAER_B200_LU_CLUSTER=16 timeout 120 python tools/time_lu_panel.py 512 2>&1 | tee gpurun_out/r02_l_panel512.log
timeout 300 python tools/time_factor.py lu 8192 16384 32768 2>&1 | tail -3 | tee gpurun_out/r02_l_lu_time.log
FAER_B200_TRACE=1 timeout 200 python tools/time_factor.py lu 16384 > gpurun_out/r02_l_lu_trace.log 2>&1
timeout 200 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee gpurun_out/r02_l_llt.log
FAER_B200_NB=512 timeout 200 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee -a gpurun_out/r02_l_llt.log
