# Round 2, call K: the LU panel chain in isolation (per-column cost by height, launch list of one 32768 x 512 panel).
mkdir -p gpurun_out
export FAER_B200_LU_CLUSTER=16
FAER_B200_LU_SUBPANEL_PROF=1 timeout 120 python tools/time_lu_panel.py 128 2>&1 | tee gpurun_out/r02_k_panel128.log
timeout 120 python tools/time_lu_panel.py 512 2>&1 | tee gpurun_out/r02_k_panel512.log
FAER_B200_LU_FUSED_W=0 timeout 120 python tools/time_lu_panel.py 512 2>&1 | tee gpurun_out/r02_k_panel512_nofused.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_k_panel_launches.csv python tools/time_lu_panel.py 512 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_k_panel_launches.csv 2>/dev/null | head -30
