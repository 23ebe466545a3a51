# Round 2, call B: new QR rank-deficient path, the TMA/warp-specialised f64 GEMM and the sliced GEMM (parity first, then time).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_qr.py tests/test_gpu_zz10_gemm_ws_sliced.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_b_tests.log
timeout 400 python tools/time_gemm_modes.py 2>&1 | tee gpurun_out/r02_gemm_modes.log
