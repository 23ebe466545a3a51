# Round 2, call S: final state — whole GPU suite, smoke, bench line, ncu captures of the final LU sub-panel kernel and of the ws
# kernel on the LU update (k = 512).
mkdir -p gpurun_out
PYT="python -m pytest -m gpu -q --tb=short -o faulthandler_timeout=300 -p no:cacheprovider"
timeout 1500 $PYT tests > gpurun_out/r02_s_tests.log 2>&1; tail -6 gpurun_out/r02_s_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_s_bench.log 2>&1; tail -1 gpurun_out/r02_s_bench.log | cut -c1-300
FAER_B200_LU_CLUSTER=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_subpanel --launch-skip 8 --launch-count 1 -o gpurun_out/r02_lu_subpanel_final -f python tools/time_lu_panel.py 128 > gpurun_out/r02_s_ncu1.log 2>&1; tail -1 gpurun_out/r02_s_ncu1.log
cat > /tmp/run_upd.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import faer_b200
from faer_b200 import linalg as la
dev = torch.device("cuda:0"); lib = faer_b200.load(); lib.faer_b200_set_stream(torch.cuda.current_stream().cuda_stream)
m = n = 16384; k = 512
L = torch.randn((k, m), dtype=torch.float64, device=dev).T; U = torch.randn((n, k), dtype=torch.float64, device=dev).T
C = torch.randn((n, m), dtype=torch.float64, device=dev).T
for _ in range(4): la.matmul(C, 1, L, U, -1.0)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_ws --launch-skip 2 --launch-count 1 -o gpurun_out/r02_ws_update_k512 -f python /tmp/run_upd.py > gpurun_out/r02_s_ncu2.log 2>&1; tail -1 gpurun_out/r02_s_ncu2.log
