# Round 2, call G: which ingredient of the ws GEMM makes Add-mode products go wrong now and then (bring-up variants).
mkdir -p gpurun_out
for cfg in "0 0" "0 20000" "1 0" "3 0" "4 0"; do
  set -- $cfg
  FAER_B200_WS_VAR=$1 FAER_B200_WS_EXTRA_SMEM=$2 timeout 120 python tools/debug_ws_variants.py 6 2>&1 | tail -3 | tee -a gpurun_out/r02_g_variants.log
done
