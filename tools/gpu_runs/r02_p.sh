# Round 2, call P (8 GPUs): configs[2] at N = 8 with rank 0's timeline; smoke(); reference arm under torchrun.
mkdir -p gpurun_out
nvidia-smi -L | wc -l | tee gpurun_out/r02_p_gpus.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631"
FAER_B200_TRACE=1 timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_p_bench_n8.log 2> gpurun_out/r02_p_bench_n8.err; tail -1 gpurun_out/r02_p_bench_n8.log | cut -c1-1200
awk '/dist LU/{c++} c==5' gpurun_out/r02_p_bench_n8.err | head -70 > gpurun_out/r02_p_trace_n8.log; sed -n 1,14p gpurun_out/r02_p_trace_n8.log; sed -n 56,66p gpurun_out/r02_p_trace_n8.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
