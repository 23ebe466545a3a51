// Host build of the second translation unit of the C ABI (csrc/ffi_types.cu) together with the real runtime (csrc/runtime.cu: entry
// lock, workspace pool, StagedMat staging) and the three flat-map drivers, all as they are, against hostcuda/cuda_runtime.h. The
// resulting library exports the very `libfaer_v0_23_*` symbols of ffi_types.cu, so the Python binding (capi structs, linalg wrappers)
// and the GPU tests' own functions run against it on the CPU (tests/test_ffi_types_host_cpu.py); only the building blocks underneath
// (structured products, triangular solves, Householder sequences, the real condensed solvers) are forwarded to the test's callback.
// TEST INFRASTRUCTURE: nothing in the product links this file.
#include "mock_blocks.hpp"

long long hostcuda_guard_errors = 0, hostcuda_live_allocs = 0;

#include "../../faer-rs_b200/csrc/runtime.cu"
#include "../../faer-rs_b200/csrc/cplx_condensed.cu"
#include "../../faer-rs_b200/csrc/ldlt_types.cu"
#include "../../faer-rs_b200/csrc/reconstruct_types.cu"
#include "../../faer-rs_b200/csrc/ffi_types.cu"

extern "C" {
void drivers_set_callback(mock_cb_t cb) { g_cb = cb; }
void drivers_set_reverse(int r) { fb::flat_map_host_reverse = r != 0; }
long long drivers_guard_errors() { return hostcuda_guard_errors; }
// blocks still owned by the pool after ws_release_all() would be a leak of the entry points
long long drivers_live_blocks() {
  fb::ws_release_all();
  return hostcuda_live_allocs;
}
}
