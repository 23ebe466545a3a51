# Round 2, call J (2 GPUs): multi-rank parity (LLT + LU vs the single-GPU run), then the configs[2] bench line at N = 2 with the
# event timeline of rank 0, and the old (no partition) schedule for comparison.
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r02_j_gpus.log
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_dist_multi.py 2>&1 | tail -15 | tee gpurun_out/r02_j_multi_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
FAER_B200_TRACE=1 timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_j_bench_n2.log 2> gpurun_out/r02_j_bench_n2.err; tail -1 gpurun_out/r02_j_bench_n2.log | cut -c1-1500
grep -A70 "dist LU" gpurun_out/r02_j_bench_n2.err | tail -70 > gpurun_out/r02_j_trace_n2.log; tail -25 gpurun_out/r02_j_trace_n2.log
FAER_B200_DIST_NO_PARTITION=1 timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e 2>/dev/null | tail -1 | cut -c1-400 | tee gpurun_out/r02_j_bench_n2_nopart.log
