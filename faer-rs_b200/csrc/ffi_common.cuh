// Helpers shared by the translation units of the extern "C" boundary (ffi.cu, ffi_types.cu): scalar / permutation arguments that may
// live on the host or on the device, and the end-of-call sequence of the staged matrices.
#pragma once
#include "../../include/faer_b200.h"
#include "runtime.cuh"

#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <vector>

namespace fb {
// (static, not an anonymous namespace: nvcc's kernel stubs of a unit that also has a global anonymous namespace must stay unambiguous)

static inline double read_scalar_f64(const FaerV0_24_Scalar* p) {
  FB_ASSERT(p != nullptr, "null scalar pointer");
  double v;
  if (is_device_pointer(p)) {
    FB_CUDA_CHECK(cudaMemcpy(&v, p, sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    memcpy(&v, p, sizeof(double));
  }
  return v;
}


static inline void finish_all(cudaStream_t st, std::initializer_list<StagedMat*> mats) {
  // the compute must be complete before input mirrors return to the pool; calls are synchronous anyway
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  for (auto* m : mats) m->finish();
}


static inline double read_real(const void* p, double) { return read_scalar_f64((const FaerV0_24_Scalar*)p); }
static inline float read_real(const void* p, float) {
  FB_ASSERT(p != nullptr, "null scalar pointer");
  float v;
  if (is_device_pointer(p)) FB_CUDA_CHECK(cudaMemcpy(&v, p, sizeof(float), cudaMemcpyDeviceToHost));
  else memcpy(&v, p, sizeof(float));
  return v;
}

// index slices (u32 / u64) <-> host int64
static inline std::vector<long long> read_perm(const void* p, size_t n, int idx_bytes) {
  std::vector<unsigned char> raw(n * (size_t)idx_bytes);
  if (n) {
    if (is_device_pointer(p)) FB_CUDA_CHECK(cudaMemcpy(raw.data(), p, raw.size(), cudaMemcpyDeviceToHost));
    else memcpy(raw.data(), p, raw.size());
  }
  std::vector<long long> out(n);
  for (size_t i = 0; i < n; ++i)
    out[i] = idx_bytes == 4 ? (long long)((const uint32_t*)raw.data())[i] : (long long)((const uint64_t*)raw.data())[i];
  return out;
}
static inline void write_perm(void* dst, const std::vector<long long>& v, int idx_bytes) {
  if (v.empty()) return;
  std::vector<unsigned char> buf(v.size() * (size_t)idx_bytes);
  for (size_t i = 0; i < v.size(); ++i) {
    if (idx_bytes == 4) ((uint32_t*)buf.data())[i] = (uint32_t)v[i];
    else ((uint64_t*)buf.data())[i] = (uint64_t)v[i];
  }
  if (is_device_pointer(dst)) FB_CUDA_CHECK(cudaMemcpy(dst, buf.data(), buf.size(), cudaMemcpyHostToDevice));
  else memcpy(dst, buf.data(), buf.size());
}

}  // namespace fb
