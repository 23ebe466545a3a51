# Round 2, call H: production ws kernel (proxy fence) correctness, whole GPU suite, bench line, LU sub-panel phase profile,
# ncu capture of the LLT trailing update (SYRK shape) on the ws kernel.
mkdir -p gpurun_out
timeout 120 python tools/debug_ws_variants.py 4 2>&1 | tail -3 | tee gpurun_out/r02_h_ws_check.log
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 1500 $PYT tests > gpurun_out/r02_h_tests.log 2>&1; tail -30 gpurun_out/r02_h_tests.log
timeout 600 python bench.py > gpurun_out/r02_h_bench.log 2>&1; tail -1 gpurun_out/r02_h_bench.log
FAER_B200_LU_SUBPANEL_PROF=1 timeout 200 python tools/time_factor.py lu 16384 2>&1 | tail -3 | tee gpurun_out/r02_h_lu_prof.log
NCU="ncu --set full --clock-control none --import-source on -k regex:gemm_f64_ws --launch-skip 2 --launch-count 1"
timeout 300 $NCU -o gpurun_out/r02_syrk_ws -f python tools/run_syrk_once.py 2 16128 256 > gpurun_out/r02_h_ncu1.log 2>&1; tail -2 gpurun_out/r02_h_ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_subpanel --launch-skip 3 --launch-count 1 -o gpurun_out/r02_lu_subpanel -f python tools/time_factor.py lu 8192 > gpurun_out/r02_h_ncu2.log 2>&1; tail -2 gpurun_out/r02_h_ncu2.log
timeout 300 python tools/time_gemm_modes.py quick > gpurun_out/r02_h_gemm_modes.log 2>&1; tail -14 gpurun_out/r02_h_gemm_modes.log
