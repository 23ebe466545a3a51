// Host stand-in for <cuda_runtime.h>, just enough for the flat-map drivers of faer-rs_b200/csrc (cplx_condensed.cu, ldlt_types.cu,
// reconstruct_types.cu) and the headers they include to compile with a plain C++ compiler: "device memory" is host memory, a stream
// is an opaque pointer, copies are memcpy. TEST INFRASTRUCTURE (tools/emul/drivers_host.cpp); never on the product's include path.
#pragma once
#include <cstdlib>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { memcpy(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t) { memset(dst, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorName(cudaError_t) { return "cudaSuccess"; }
inline const char* cudaGetErrorString(cudaError_t) { return "no error"; }
