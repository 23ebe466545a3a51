mkdir -p gpurun_out
rm -f gpurun_out/nb_sweep2.log
for NB in 512 768 1024 1536; do
  echo "--- LLT NB=$NB" | tee -a gpurun_out/nb_sweep2.log
  FAER_B200_NB=$NB timeout 100 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee -a gpurun_out/nb_sweep2.log
done
for NB in 256 384 512 768; do
  echo "--- LU NB=$NB" | tee -a gpurun_out/nb_sweep2.log
  FAER_B200_NB=$NB timeout 100 python tools/time_factor.py lu 16384 2>&1 | tail -1 | tee -a gpurun_out/nb_sweep2.log
done
