# Round 2, call C: GEMM mode timings first, then the new tests with per-test durations and a stack dump for anything slow.
mkdir -p gpurun_out
timeout 420 python tools/time_gemm_modes.py > gpurun_out/r02_gemm_modes.log 2>&1
tail -25 gpurun_out/r02_gemm_modes.log
PYT="python -m pytest -m gpu -x -v --durations=15 -o faulthandler_timeout=120 -p no:cacheprovider"
timeout 420 $PYT tests/test_gpu_zz10_gemm_ws_sliced.py > gpurun_out/r02_c_zz10.log 2>&1; tail -12 gpurun_out/r02_c_zz10.log
timeout 500 $PYT tests/test_gpu_qr.py > gpurun_out/r02_c_qr.log 2>&1; tail -30 gpurun_out/r02_c_qr.log
timeout 500 $PYT tests/test_gpu_zz11_evd_svd_vectors.py > gpurun_out/r02_c_zz11.log 2>&1; tail -30 gpurun_out/r02_c_zz11.log
