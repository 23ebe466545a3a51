# Round 2, call X: tall LU panels (32768 / 16384 rows x 512) with narrower fused sub-panels (more levels on the host recursion).
mkdir -p gpurun_out
for W in 128 64 32; do
  echo "--- FAER_B200_LU_FUSED_W=$W (halved above 16384 rows)" | tee -a gpurun_out/r02_x_panel.log
  FAER_B200_LU_CLUSTER=16 FAER_B200_LU_FUSED_W=$W PANEL_ROWS=32768,16384,4096 timeout 100 python tools/time_lu_panel.py 512 2>&1 | tee -a gpurun_out/r02_x_panel.log
done
