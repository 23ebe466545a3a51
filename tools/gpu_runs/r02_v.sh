# Round 2, call V: potf2 with rcp.rn (bit-identity tests), LLT timing, full ncu capture of one potf2 launch (source-level stalls).
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py tests/test_gpu_zz5_llt_f32.py tests/test_gpu_dist.py -k "llt or LLT or cholesky" 2>&1 | tail -4 | tee gpurun_out/r02_v_llt_tests.log
timeout 200 python tools/time_factor.py llt 16384 2>&1 | tail -1 | tee gpurun_out/r02_v_llt.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:potf2_kernel --launch-skip 40 --launch-count 1 -o gpurun_out/r02_potf2 -f python tools/time_factor.py llt 8192 > gpurun_out/r02_v_ncu.log 2>&1; tail -1 gpurun_out/r02_v_ncu.log
