# Round 2, call E (re-entry after the container was re-created): whole GPU suite with durations, bench line, GEMM mode timings, time_other.
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -v --durations=25 -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 1200 $PYT tests > gpurun_out/r02_e_tests.log 2>&1
grep -E "PASSED|FAILED|ERROR|passed|failed|error" gpurun_out/r02_e_tests.log | grep -v PASSED | tail -40
tail -30 gpurun_out/r02_e_tests.log
timeout 600 python bench.py > gpurun_out/r02_e_bench.log 2>&1; tail -2 gpurun_out/r02_e_bench.log
timeout 420 python tools/time_gemm_modes.py > gpurun_out/r02_e_gemm_modes.log 2>&1; tail -25 gpurun_out/r02_e_gemm_modes.log
timeout 300 python tools/time_other.py all > gpurun_out/r02_e_time_other.log 2>&1; tail -12 gpurun_out/r02_e_time_other.log
