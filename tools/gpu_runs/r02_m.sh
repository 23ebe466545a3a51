# Round 2, call M (4 GPUs): LU parity on the final sub-panel kernel, panel chain timing, multi-rank parity at 2 and 4 ranks,
# configs[2] at N = 4 with rank 0's timeline.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
FAER_B200_LOOKAHEAD_MIN_N=0 FAER_B200_LU_CLUSTER=16 timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py -k "plu or lu or pivot" 2>&1 | tail -4 | tee gpurun_out/r02_m_lu_rec_tests.log
timeout 300 $PYT tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_solve.py -k "plu or lu or pivot" 2>&1 | tail -4 | tee gpurun_out/r02_m_lu_part_tests.log
FAER_B200_LU_CLUSTER=16 FAER_B200_LU_SUBPANEL_PROF=1 timeout 120 python tools/time_lu_panel.py 128 512 2>&1 | tee gpurun_out/r02_m_panel.log
timeout 200 python tools/time_factor.py lu 8192 16384 2>&1 | tail -2 | tee gpurun_out/r02_m_lu_time.log
timeout 600 $PYT tests/test_gpu_dist_multi.py 2>&1 | tail -6 | tee gpurun_out/r02_m_multi_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621"
FAER_B200_TRACE=1 timeout 900 $TR bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02_m_bench_n4.log 2> gpurun_out/r02_m_bench_n4.err; tail -1 gpurun_out/r02_m_bench_n4.log | cut -c1-900
awk '/dist LU/{c++} c==5' gpurun_out/r02_m_bench_n4.err | head -70 > gpurun_out/r02_m_trace_n4.log; sed -n 1,12p gpurun_out/r02_m_trace_n4.log; sed -n 50,66p gpurun_out/r02_m_trace_n4.log
