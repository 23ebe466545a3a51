# Round 2, call U: host-resident LLT n = 16384: default pipeline vs the opt-in left-looking first half, block widths.
mkdir -p gpurun_out
for HL in 0 1; do
  for NB in 256 512; do
    echo "--- FAER_B200_HOST_LEFT=$HL FAER_B200_NB=$NB" | tee -a gpurun_out/r02_u_e2e.log
    FAER_B200_HOST_LEFT=$HL FAER_B200_NB=$NB timeout 200 python tools/time_e2e.py 16384 2>&1 | tail -3 | tee -a gpurun_out/r02_u_e2e.log
  done
done
