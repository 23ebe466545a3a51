# Round 2, call N: c64 planar path (parity + time), adaptive fused LU width, the whole GPU suite, bench, ncu launch list of the bench.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 1500 $PYT tests > gpurun_out/r02_n_tests.log 2>&1; tail -8 gpurun_out/r02_n_tests.log
FAER_B200_LU_CLUSTER=16 timeout 120 python tools/time_lu_panel.py 512 2>&1 | head -3 | tee gpurun_out/r02_n_panel.log
FAER_B200_LU_CLUSTER=16 FAER_B200_LU_FUSED_TALL=1000000 timeout 120 python tools/time_lu_panel.py 512 2>&1 | head -2 | tee -a gpurun_out/r02_n_panel.log
timeout 200 python tools/time_other.py all 2>&1 | tail -9 | tee gpurun_out/r02_n_time_other.log
timeout 600 python bench.py > gpurun_out/r02_n_bench.log 2>&1; tail -1 gpurun_out/r02_n_bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_n_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_n_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_n_bench_launches.csv | head -30 | tee gpurun_out/r02_n_bench_launch_summary.txt
