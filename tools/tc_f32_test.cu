// Dev aid: standalone bring-up test of the tcgen05 (kind::tf32) 3xTF32 GEMM kernel (csrc/gemm_f32_tc.cuh).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I faer-rs_b200/csrc -o tools/tc_f32_test tools/tc_f32_test.cu
// run:   tools/tc_f32_test            (checks 256x256x128 and odd shapes against an f64 host reference, then times large sizes)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_f32_tc.cuh"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

static double check(int m, int n, int k, bool accum, float alpha) {
  std::vector<float> hA((size_t)m * k), hB((size_t)k * n), hC((size_t)m * n), hC0((size_t)m * n);
  srand(1234 + m + 7 * n + 13 * k);
  for (auto& x : hA) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : hB) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto& x : hC0) x = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, hA.size() * 4));
  CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMalloc(&dC, hC.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, hC0.data(), hC.size() * 4, cudaMemcpyHostToDevice));
  // column-major A (m x k, ld m), B (k x n, ld k), C (m x n, ld m): the library's default layout
  fb::tc::Operand a{dA, m, k, 1, m}, b{dB, k, n, 1, k};
  fb::tc::Workspace ws;
  if (!fb::tc::gemm_f32_tc(0, dC, 1, m, m, n, k, accum ? 1 : 0, a, b, alpha, &ws)) {
    printf("gemm_f32_tc refused the problem %dx%dx%d\n", m, n, k);
    return -1;
  }
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int j = 0; j < n; j += (n > 512 ? 37 : 1))
    for (int i = 0; i < m; i += (m > 512 ? 41 : 1)) {
      double s = 0;
      for (int q = 0; q < k; ++q) s += (double)hA[(size_t)q * m + i] * (double)hB[(size_t)j * k + q];
      const double ref = alpha * s + (accum ? (double)hC0[(size_t)j * m + i] : 0.0);
      maxerr = fmax(maxerr, fabs(ref - (double)hC[(size_t)j * m + i]));
      maxref = fmax(maxref, fabs(ref));
    }
  fb::tc::release(&ws);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  printf("check %5d x %5d x %5d accum=%d alpha=%g: max abs err %.3e (max |ref| %.3e, rel %.2e)\n", m, n, k, (int)accum, alpha,
         maxerr, maxref, maxerr / maxref);
  return maxerr / maxref;
}

static void timeit(int n) {
  float *dA, *dB, *dC;
  CK(cudaMalloc(&dA, (size_t)n * n * 4));
  CK(cudaMalloc(&dB, (size_t)n * n * 4));
  CK(cudaMalloc(&dC, (size_t)n * n * 4));
  CK(cudaMemset(dA, 0, (size_t)n * n * 4));
  CK(cudaMemset(dB, 0, (size_t)n * n * 4));
  fb::tc::Operand a{dA, n, n, 1, n}, b{dB, n, n, 1, n};
  fb::tc::Workspace ws;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fb::tc::gemm_f32_tc(0, dC, 1, n, n, n, n, 0, a, b, 1.f, &ws);
  CK(cudaDeviceSynchronize());
  float best = 1e30f, best_mma = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    fb::tc::gemm_f32_tc(0, dC, 1, n, n, n, n, 0, a, b, 1.f, &ws);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = fminf(best, ms);
    best_mma = fminf(best_mma, ws.last_mma_ms);
  }
  printf("time n=%d: %.3f ms total (pack + mma) = %.1f TFLOP/s fp32-accurate\n", n, best, 2.0 * n * n * n / best / 1e9);
  fb::tc::release(&ws);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

int main() {
  double worst = 0;
  worst = fmax(worst, check(256, 256, 128, false, 1.f));
  worst = fmax(worst, check(128, 128, 32, false, 1.f));
  worst = fmax(worst, check(384, 256, 96, true, -0.5f));
  worst = fmax(worst, check(300, 200, 70, false, 2.f));
  worst = fmax(worst, check(1000, 333, 517, true, 1.f));
  worst = fmax(worst, check(4096, 2048, 1024, false, 1.f));
  worst = fmax(worst, check(256, 384, 16384, true, -1.f));  // split-K path (6 tiles)
  worst = fmax(worst, check(200, 100, 5000, false, 1.f));   // split-K, ragged
  printf("worst relative error %.3e (%s)\n", worst, worst < 2e-5 ? "OK" : "FAIL");
  timeit(4096);
  timeit(8192);
  return worst < 2e-5 ? 0 : 1;
}
