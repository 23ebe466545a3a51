# Round 2, call Q: templated LLT (f64 bitwise leaf parity, f32), f32 LU entry points, LLT timing.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 900 $PYT tests/test_gpu_zz13_lu_f32.py tests/test_gpu_zz5_llt_f32.py tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py tests/test_gpu_solve.py tests/test_gpu_zz3_solvers.py tests/test_gpu_zz4_reconstruct_inverse.py tests/test_gpu_dist.py 2>&1 | tail -25 | tee gpurun_out/r02_q_tests.log
timeout 200 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee gpurun_out/r02_q_llt.log
