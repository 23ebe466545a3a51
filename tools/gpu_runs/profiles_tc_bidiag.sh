mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.log
timeout 300 python tools/time_other.py all 2>&1 | tee gpurun_out/time_other.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f32_tc_kernel --launch-skip 13 -c 1 -o gpurun_out/tc_gemm8192 -f tools/tc_f32_test > gpurun_out/ncu_tc.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:bidiag_kernel -c 1 -o gpurun_out/bidiag8192 -f python tools/run_condensed_once.py bidiag 8192 > gpurun_out/ncu_bidiag.log 2>&1
