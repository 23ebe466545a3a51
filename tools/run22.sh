mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 200 python tools/time_factor.py lu 8192 16384 2>&1 | tail -2 | tee gpurun_out/lu_default.log
FAER_B200_LU_WIDE=0 timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -1 | tee -a gpurun_out/lu_default.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qr_panel_kernel --launch-skip 300 -c 1 -o gpurun_out/qr_panel_f32 -f python tools/time_other.py qr > gpurun_out/ncu_qr.log 2>&1
