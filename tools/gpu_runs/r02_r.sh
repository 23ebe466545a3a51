# Round 2, call R: QR f32 65536 x 4096 with the panel grid capped (fewer, fatter CTAs) + parity at the best setting.
mkdir -p gpurun_out
for G in 0 112 96 80 64 48; do
  echo "--- FAER_B200_QR_PANEL_CTAS=$G" | tee -a gpurun_out/r02_r_qr_ctas.log
  FAER_B200_QR_PANEL_CTAS=$G timeout 120 python tools/time_other.py qr 2>&1 | tail -2 | tee -a gpurun_out/r02_r_qr_ctas.log
done
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
FAER_B200_QR_PANEL_CTAS=64 timeout 600 $PYT tests/test_gpu_qr.py tests/test_gpu_zz2_qr_solve.py 2>&1 | tail -4 | tee gpurun_out/r02_r_qr_tests.log
