mkdir -p gpurun_out
timeout 300 python tools/time_e2e.py 16384 2>&1 | tail -5 | tee gpurun_out/e2e.log
FAER_B200_LOOKAHEAD_MIN_N=0 FAER_B200_LU_CLUSTER=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "plu or lu or pivot" 2>&1 | tail -5 | tee gpurun_out/pytest_lu_cluster8.log
FAER_B200_LOOKAHEAD_MIN_N=0 FAER_B200_LU_CLUSTER=16 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "plu or lu or pivot" 2>&1 | tail -5 | tee gpurun_out/pytest_lu_cluster16.log
rm -f gpurun_out/cluster_lu.log
for C in 0 8 16; do
echo "--- recursive driver, cluster=$C" | tee -a gpurun_out/cluster_lu.log
FAER_B200_LOOKAHEAD_MIN_N=0 FAER_B200_LU_CLUSTER=$C timeout 200 python tools/time_factor.py lu 8192 16384 2>&1 | tail -2 | tee -a gpurun_out/cluster_lu.log
done
for C in 0 16; do
echo "--- GREEN_SMS=16 cluster=$C" | tee -a gpurun_out/cluster_lu.log
FAER_B200_LU_CLUSTER=$C FAER_B200_GREEN_SMS=16 FAER_B200_TRACE=1 timeout 200 python tools/time_factor.py lu 16384 > gpurun_out/cluster_lu_trace_$C.log 2>&1; tail -1 gpurun_out/cluster_lu_trace_$C.log | tee -a gpurun_out/cluster_lu.log
done
echo "--- GREEN_SMS=16 cluster=16 n=32768" | tee -a gpurun_out/cluster_lu.log
FAER_B200_GREEN_SMS=16 timeout 300 python tools/time_factor.py lu 32768 2>&1 | tail -1 | tee -a gpurun_out/cluster_lu.log
