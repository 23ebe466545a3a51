#!/bin/bash
# round 2, call Z3: the round-end sequence on the final tree: full GPU suite (-x), smoke, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02_z3_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/r02_z3_smoke.log
timeout 900 python bench.py 2>&1 | tail -3 | tee gpurun_out/r02_z3_bench.log
