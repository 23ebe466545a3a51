// Flat-map launches: body(i, j) for i < nx (rounded up to whole 256-thread blocks: every body checks its own bounds), j < ny, with
// no communication between the threads of one launch. The "functional, not tuned" drivers (cplx_condensed.cu, ldlt_types.cu,
// reconstruct_types.cu) are written as sequences of such launches over plain functor structs. Compiled by nvcc, DevRun launches a
// kernel; compiled by a plain C++ compiler (the host builds under tools/emul/, test infrastructure), the same DevRun visits the same
// index space with loops — forwards or, when `flat_map_host_reverse` is set, backwards, which shows that no body depends on
// another body of the same launch.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace fb {

#if defined(__CUDACC__)
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
#else
inline bool flat_map_host_reverse = false;

struct DevRun {
  cudaStream_t st;
  template <class B>
  void operator()(const B& body, i64 nx, i64 ny) const {
    if (nx <= 0 || ny <= 0) return;
    note_launch();
    const i64 gx = ((nx + 255) / 256) * 256;
    if (!flat_map_host_reverse) {
      for (i64 j = 0; j < ny; ++j)
        for (i64 i = 0; i < gx; ++i) body(i, j);
    } else {
      for (i64 j = ny - 1; j >= 0; --j)
        for (i64 i = gx - 1; i >= 0; --i) body(i, j);
    }
  }
};
#endif

}  // namespace fb
