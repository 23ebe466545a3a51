# Round 2, call Y: c64 triangular solves, LLT and partial-pivoting LU (new entry points): parity tests, timing.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 600 $PYT tests/test_gpu_zzz_c64_llt_trsm_lu.py tests/test_gpu_zz13_lu_f32.py tests/test_gpu_parity.py -k "c64 or llt or lu" 2>&1 | tail -25 | tee gpurun_out/r02_y_tests.log
timeout 200 python tools/time_c64_llt.py 2>&1 | tail -2 | tee gpurun_out/r02_y_c64_llt.log
