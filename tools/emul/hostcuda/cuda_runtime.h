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

// ---- the rest of what runtime.cu and ffi_types.cu use (tools/emul/ffi_types_host.cpp) ----
// "device" allocations carry guard words checked on free; every user pointer is reported as host memory, so the staging layer of
// the C ABI always mirrors (which is the path the host build is there to exercise).
extern long long hostcuda_guard_errors, hostcuda_live_allocs;
inline cudaError_t cudaMalloc(void** p, size_t bytes) {
  unsigned char* raw = (unsigned char*)malloc(sizeof(size_t) + 64 + bytes + 64);
  if (!raw) return 2;
  *(size_t*)raw = bytes;
  memset(raw + sizeof(size_t), 0xA5, 64);
  memset(raw + sizeof(size_t) + 64, 0xFF, bytes);
  memset(raw + sizeof(size_t) + 64 + bytes, 0x5A, 64);
  ++hostcuda_live_allocs;
  *p = raw + sizeof(size_t) + 64;
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  unsigned char* user = (unsigned char*)p;
  unsigned char* raw = user - 64 - sizeof(size_t);
  const size_t bytes = *(size_t*)raw;
  for (size_t k = 0; k < 64; ++k) {
    if (raw[sizeof(size_t) + k] != 0xA5) ++hostcuda_guard_errors;
    if (user[bytes + k] != 0x5A) ++hostcuda_guard_errors;
  }
  --hostcuda_live_allocs;
  free(raw);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                     cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
  return cudaSuccess;
}
enum cudaMemoryType { cudaMemoryTypeUnregistered, cudaMemoryTypeHost, cudaMemoryTypeDevice, cudaMemoryTypeManaged };
struct cudaPointerAttributes {
  cudaMemoryType type;
};
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
typedef void* cudaEvent_t;
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
