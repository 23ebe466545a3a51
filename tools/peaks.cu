// Measures the f64 roofline denominators on the box: DMMA.8x8x4 issue-bound peak, DFMA peak, and (as a
// sanity yardstick only, never on the product path) cuBLAS DGEMM at a few sizes.
// Build: make -C faer-rs_b200 tools/peaks     Run: tools/peaks [json_out]
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void dmma_peak(double* out, int iters) {
  double c[NACC][2];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void dfma_peak(double* out, int iters) {
  double c[NACC];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

template <class F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device: %s, %d SMs, %.0f MHz\n", prop.name, sms, prop.clockRate / 1e3);
  double* out; CK(cudaMalloc(&out, 1024));
  double best_dmma = 0, best_dfma = 0;
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { dmma_peak<16><<<sms, warps * 32>>>(out, iters); }, 5);
    double flops = 2.0 * 256 * 16.0 * iters * warps * sms;
    double tf = flops / (ms * 1e-3) / 1e12;
    printf("DMMA 16 acc, %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", warps, ms, tf);
    if (tf > best_dmma) best_dmma = tf;
  }
  for (int warps : {8, 16, 32}) {
    float ms = time_ms([&] { dfma_peak<16><<<sms, warps * 32>>>(out, iters); }, 5);
    double flops = 2.0 * 32 * 16.0 * iters * warps * sms;
    double tf = flops / (ms * 1e-3) / 1e12;
    printf("DFMA 16 acc, %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", warps, ms, tf);
    if (tf > best_dfma) best_dfma = tf;
  }
  // sustained DMMA (~2 s) to see the power-capped figure
  double sustained = 0;
  {
    int reps = 0;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    float total = 0;
    while (total < 2000.f && reps < 100000) {
      for (int i = 0; i < 20; ++i) dmma_peak<16><<<sms, 16 * 32>>>(out, iters);
      reps += 20;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&total, e0, e1));
    }
    sustained = 2.0 * 256 * 16.0 * iters * 16 * sms * reps / (total * 1e-3) / 1e12;
    printf("DMMA sustained over %.0f ms: %.2f TFLOP/s\n", total, sustained);
  }
  // cuBLAS DGEMM yardstick
  cublasHandle_t h; cublasCreate(&h);
  std::vector<std::pair<int, double>> cub;
  for (int n : {4096, 8192, 16384}) {
    double *A, *B, *C;
    size_t bytes = (size_t)n * n * 8;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes));
    double one = 1.0, zero = 0.0;
    float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, n >= 16384 ? 2 : 4);
    double tf = 2.0 * n * (double)n * n / (ms * 1e-3) / 1e12;
    printf("cuBLAS DGEMM n=%d: %.3f ms  %.2f TFLOP/s\n", n, ms, tf);
    cub.push_back({n, tf});
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  if (argc > 1) {
    FILE* f = fopen(argv[1], "w");
    fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"dmma_tflops_burst\": %.3f, \"dmma_tflops_sustained\": %.3f, \"dfma_tflops\": %.3f, \"cublas_dgemm_tflops\": {",
            prop.name, sms, best_dmma, sustained, best_dfma);
    for (size_t i = 0; i < cub.size(); ++i) fprintf(f, "%s\"%d\": %.3f", i ? ", " : "", cub[i].first, cub[i].second);
    fprintf(f, "}}\n");
    fclose(f);
  }
  return 0;
}
