# Round 2, call A: opt-in look-ahead QR (parity + timing) and the int8-sliced f64 GEMM bring-up.
mkdir -p gpurun_out
FAER_B200_QR_LOOKAHEAD=48 timeout 300 python -m pytest tests/test_gpu_qr.py tests/test_gpu_zz2_qr_solve.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_qr_lookahead_tests.log
timeout 200 python tools/time_other.py qr 2>&1 | tee gpurun_out/r02_qr_default.log
for sms in 32 48 64; do
  FAER_B200_QR_LOOKAHEAD=$sms FAER_B200_VERBOSE=1 timeout 200 python tools/time_other.py qr 2>&1 | tee gpurun_out/r02_qr_lookahead_$sms.log
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I faer-rs_b200/csrc -o gpurun_out/ozaki_test tools/next/ozaki_test.cu -lcuda \
  && timeout 120 gpurun_out/ozaki_test 2>&1 | tee gpurun_out/r02_ozaki_test.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv | tee gpurun_out/r02_smi.log
