#!/bin/bash
# round 2, call Z5: final tree: full GPU suite (-x) + timings of the complex factorizations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02_z5_tests.log
timeout 300 python tools/time_cplx.py 2>&1 | tail -14 | tee gpurun_out/r02_z5_time_cplx.log
