# First GPU session of round 2: validate what was written after round 1's GPU budget ran out, in order of risk.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_runs/r02_first.sh'
mkdir -p gpurun_out
# 1. the whole GPU suite (new since the last hardware run: tests/test_gpu_zz*.py and the entry lock)
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
# 2. opt-in look-ahead QR on a partitioned GPU: parity first, then time against the default driver
FAER_B200_QR_LOOKAHEAD=48 timeout 300 python -m pytest tests/test_gpu_qr.py tests/test_gpu_zz2_qr_solve.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_qr_lookahead_tests.log
timeout 200 python tools/time_other.py qr 2>&1 | tee gpurun_out/r02_qr_default.log
for sms in 48 64; do
  FAER_B200_QR_LOOKAHEAD=$sms FAER_B200_VERBOSE=1 timeout 200 python tools/time_other.py qr 2>&1 | tee gpurun_out/r02_qr_lookahead_$sms.log
done
# 3. bring-up of the int8-sliced f64 GEMM on tcgen05 (accuracy table first; a hang is bounded by the timeout)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I faer-rs_b200/csrc -o gpurun_out/ozaki_test tools/next/ozaki_test.cu -lcuda \
  && timeout 120 gpurun_out/ozaki_test 2>&1 | tee gpurun_out/r02_ozaki_test.log
# 4. the bench line, unchanged code path: confirms the round-1 number on this box
timeout 500 python bench.py 2>&1 | tail -1 | tee gpurun_out/r02_bench_n1.log
