// Flat-map launches: body(i, j) for i < nx (rounded up to whole 256-thread blocks: every body checks its own bounds), j < ny, with
// no communication between the threads of one launch. The "functional, not tuned" drivers (cplx_condensed.cu, ldlt_types.cu) are
// written as sequences of such launches over plain functor structs, so that the same functors and sequences also compile for the
// host, where tools/emul/*_host.cpp runs them thread by thread (forward and in reverse thread order).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace fb {

template <class B>
__global__ void __launch_bounds__(256) flat_map_kernel(B body, long long y0) {
  body((long long)blockIdx.x * blockDim.x + threadIdx.x, y0 + (long long)blockIdx.y);
}

struct DevRun {
  cudaStream_t st;
  template <class B>
  void operator()(const B& body, i64 nx, i64 ny) const {
    if (nx <= 0 || ny <= 0) return;
    for (i64 y0 = 0; y0 < ny; y0 += 65535) {  // grid.y limit
      const unsigned nc = (unsigned)std::min<i64>(65535, ny - y0);
      flat_map_kernel<B><<<dim3((unsigned)((nx + 255) / 256), nc), 256, 0, st>>>(body, y0);
      FB_CUDA_CHECK(cudaGetLastError());
      note_launch();
    }
  }
};

}  // namespace fb
