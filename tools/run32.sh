mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.log
