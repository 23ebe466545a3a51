# Round 2, call W: phase cycles of the fused LU sub-panel kernel by panel height (one height per process).
mkdir -p gpurun_out
for M in 2048 8192 32768; do
  FAER_B200_LU_CLUSTER=16 FAER_B200_LU_SUBPANEL_PROF=1 PANEL_ROWS=$M timeout 100 python tools/time_lu_panel.py 128 2>&1 | tee -a gpurun_out/r02_w_phases.log
done
