#!/bin/bash
# First hardware run of the entry points written after round 2's GPU budget was spent:
#   gpurun --timeout 900 -- 'bash tools/gpu_runs/r02_late_additions.sh'
# Runs their GPU tests on their own (each file in its own process, so that an abort in one does not hide the others) and writes the
# logs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/late_build.log 2>&1
for f in test_gpu_zzzzzzzzz_1_inverse_triangular test_gpu_zzzzzzzzz_2_reconstruct_types test_gpu_zzzzzzzzz_3_ldlt_types test_gpu_zzzzzzzzz_4_condensed_layouts test_gpu_zzzzzzzzz_5_hessenberg test_gpu_zzzzzzzzz_6_cplx_condensed_forms test_gpu_zzzzzzzzzz_cplx_svd_evd; do
  timeout 800 python -m pytest tests/$f.py -q -m gpu -x > gpurun_out/late_$f.log 2>&1
  echo "$f: exit $?" | tee -a gpurun_out/late_summary.log
  tail -3 gpurun_out/late_$f.log | tee -a gpurun_out/late_summary.log
done
