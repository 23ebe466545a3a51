# Round 2, call D: locate the ws / cp.async disagreement, profile the SYRK-shaped update on both kernels, re-run the new tests.
mkdir -p gpurun_out
timeout 200 python tools/debug_ws_syrk.py > gpurun_out/r02_d_debug.log 2>&1; tail -30 gpurun_out/r02_d_debug.log
NCU="ncu --set full --clock-control none --import-source on -k regex:gemm_f64 --launch-skip 2 --launch-count 1"
timeout 300 $NCU -o gpurun_out/r02_syrk_ws python tools/run_syrk_once.py 2 16128 256 > gpurun_out/r02_d_ncu1.log 2>&1
timeout 300 $NCU -o gpurun_out/r02_syrk_cp python tools/run_syrk_once.py 0 16128 256 > gpurun_out/r02_d_ncu2.log 2>&1
timeout 300 $NCU -o gpurun_out/r02_rect_ws python tools/run_syrk_once.py 2 16128 256 rect > gpurun_out/r02_d_ncu3.log 2>&1
tail -3 gpurun_out/r02_d_ncu1.log gpurun_out/r02_d_ncu2.log gpurun_out/r02_d_ncu3.log
PYT="python -m pytest -m gpu -v --durations=8 -o faulthandler_timeout=120 -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_qr.py tests/test_gpu_zz10_gemm_ws_sliced.py tests/test_gpu_zz11_evd_svd_vectors.py tests/test_gpu_zz12_spicy_matmul.py tests/test_gpu_zz6_ldlt.py tests/test_gpu_condensed.py tests/test_gpu_zz7_singular_values.py tests/test_gpu_zz8_self_adjoint_eigenvalues.py > gpurun_out/r02_d_tests.log 2>&1
grep -E "PASSED|FAILED|ERROR|passed|failed" gpurun_out/r02_d_tests.log | tail -60
