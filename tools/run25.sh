mkdir -p gpurun_out
timeout 100 python tools/time_other.py qr f32 2>&1 | tee gpurun_out/time_qr32.log
FAER_B200_F32_TC=0 timeout 100 python tools/time_other.py qr f32 2>&1 | tee -a gpurun_out/time_qr32.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/qr32_launches.csv python tools/time_other.py qr f32 > gpurun_out/qr32_ncu.log 2>&1
