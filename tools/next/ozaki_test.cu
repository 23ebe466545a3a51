// Next-round bring-up tool (NOT RUN YET): stand-alone test of the int8-sliced f64 GEMM on tcgen05 (csrc/gemm_f64_ozaki.cuh).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I faer-rs_b200/csrc -o tools/next/ozaki_test tools/next/ozaki_test.cu
// run:   tools/next/ozaki_test     (errors are reported relative to (|A||B|)_ij, next to u = 2^-53; then timings)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_f64_ozaki.cuh"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

static double check(int m, int n, int k, bool accum, double alpha, int kind) {
  std::vector<double> hA((size_t)m * k), hB((size_t)k * n), hC((size_t)m * n), hC0((size_t)m * n);
  srand(99 + m + 3 * n + 7 * k);
  auto rnd = []() { return (double)rand() / RAND_MAX - 0.5; };
  for (size_t i = 0; i < hA.size(); ++i) hA[i] = rnd() * (kind == 1 ? std::exp(12.0 * rnd()) : 1.0);  // kind 1: wide range
  for (auto& x : hB) x = rnd();
  for (auto& x : hC0) x = rnd();
  double *dA, *dB, *dC;
  CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8)); CK(cudaMalloc(&dC, hC.size() * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, hC0.data(), hC.size() * 8, cudaMemcpyHostToDevice));
  fb::oz::Operand a{dA, m, k, 1, m}, b{dB, k, n, 1, k};  // column-major
  fb::oz::Workspace ws;
  if (!fb::oz::gemm_f64_ozaki(0, dC, 1, m, m, n, k, accum ? 1 : 0, a, b, alpha, &ws)) {
    printf("gemm_f64_ozaki refused %dx%dx%d\n", m, n, k);
    return -1;
  }
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hC.data(), dC, hC.size() * 8, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int j = 0; j < n; j += (n > 300 ? 29 : 1))
    for (int i = 0; i < m; i += (m > 300 ? 31 : 1)) {
      long double s = 0, sa = 0;
      for (int q = 0; q < k; ++q) {
        const long double t = (long double)hA[(size_t)q * m + i] * (long double)hB[(size_t)j * k + q];
        s += t;
        sa += fabsl(t);
      }
      const long double ref = (long double)alpha * s + (accum ? (long double)hC0[(size_t)j * m + i] : 0.0L);
      const double err = (double)(fabsl(ref - (long double)hC[(size_t)j * m + i]) / (fabsl((long double)alpha) * sa + 1e-300L));
      worst = fmax(worst, err);
    }
  fb::oz::release(&ws);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  printf("check %5d x %5d x %5d accum=%d alpha=%g kind=%d: max err / (|A||B|) = %.3e  (u = 1.1e-16)\n", m, n, k, (int)accum, alpha,
         kind, worst);
  return worst;
}

static void timeit(int n) {
  double *dA, *dB, *dC;
  CK(cudaMalloc(&dA, (size_t)n * n * 8)); CK(cudaMalloc(&dB, (size_t)n * n * 8)); CK(cudaMalloc(&dC, (size_t)n * n * 8));
  CK(cudaMemset(dA, 0, (size_t)n * n * 8)); CK(cudaMemset(dB, 0, (size_t)n * n * 8));
  fb::oz::Operand a{dA, n, n, 1, n}, b{dB, n, n, 1, n};
  fb::oz::Workspace ws;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fb::oz::gemm_f64_ozaki(0, dC, 1, n, n, n, n, 0, a, b, 1.0, &ws);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    fb::oz::gemm_f64_ozaki(0, dC, 1, n, n, n, n, 0, a, b, 1.0, &ws);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = fminf(best, ms);
  }
  printf("time n=%d: %.3f ms (exponents + slicing + 36 int8 products) = %.1f TFLOP/s f64-equivalent\n", n, best,
         2.0 * n * n * n / best / 1e9);
  fb::oz::release(&ws);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

int main() {
  double worst = 0;
  worst = fmax(worst, check(128, 64, 64, false, 1.0, 0));
  worst = fmax(worst, check(256, 128, 192, false, 1.0, 0));
  worst = fmax(worst, check(384, 200, 1000, true, -0.5, 0));
  worst = fmax(worst, check(300, 190, 70, false, 2.0, 1));
  worst = fmax(worst, check(1024, 512, 4096, false, 1.0, 0));
  printf("worst %.3e (%s: expected <= ~4e-16 for kind 0, ~2e-15 for the wide-range case)\n", worst, worst < 4e-15 ? "OK" : "FAIL");
  timeit(4096);
  timeit(8192);
  timeit(16384);
  return worst < 4e-15 ? 0 : 1;
}
