// P1 + driver: in-place lower Cholesky (LLT), f64.
//
// Reference: faer/src/linalg/cholesky/llt/factor.rs:68-97 -> ldlt/factor.rs:367-498
// (`cholesky_recursion_right_looking`: for each block column: factor A00, A10 <- A10 * L00^-H,
//  A11(lower) -= A10 * A10^H), leaf recurrence ldlt/factor.rs:7-177 / 299-366:
//     a_ij <- fma(-conj(a_jk), a_ik, a_ij)  for k = 0..j-1 (in k order),
//     d = Re(a_jj); [regularise]; fail with Err(j) if !(d > 0); l_jj = sqrt(d); fail if l_jj == 0 or non-finite;
//     column j (INCLUDING the diagonal entry) is multiplied by recip(l_jj).
//
// B200 mapping: right-looking blocked loop on the host stream; the diagonal block is factored by ONE CTA in
// shared memory with a rank-1 right-looking sweep that performs exactly the reference's per-element FMA
// chain (same k order, same reciprocal-multiply), the panel solve is G3 and the trailing update is the
// lower-masked DMMA GEMM (G2). A device status word carries the first failing column / the
// regularisation count and is read back once per factorisation.
#include "linalg_f64.cuh"

namespace fb {

namespace {

constexpr int POTF2_THREADS = 512;
constexpr int POTF2_MAX = 128;

// info[0]: first failing global column (or -1), info[1]: regularisation count
__global__ void __launch_bounds__(POTF2_THREADS) potf2_kernel(double* __restrict__ A, i64 rs, i64 cs, int n, i64 j0,
                                                               int regularize, double eps, double delta,
                                                               long long* __restrict__ info) {
  extern __shared__ double s[];  // s[i * ld + c], c <= i
  const int ld = n + 1;
  const int tid = threadIdx.x;
  if (info[0] >= 0) return;  // an earlier block already failed (uniform across the CTA)
  for (int e = tid; e < n * n; e += POTF2_THREADS) {
    int i = e % n, c = e / n;
    if (c <= i) s[i * ld + c] = A[i * rs + c * cs];
  }
  __syncthreads();
  int count = 0;
  for (int j = 0; j < n; ++j) {
    double d = s[j * ld + j];
    if (regularize) {
      // LLT: sign == +1 (reference ldlt/factor.rs:122-144)
      if (d <= eps) {
        d = delta;
        ++count;
      }
    }
    if (!(d > 0.0)) {
      if (tid == 0) info[0] = j0 + j;
      return;
    }
    const double sd = sqrt(d);
    if (sd == 0.0 || !isfinite(sd)) {
      if (tid == 0) info[0] = j0 + j;
      return;
    }
    const double inv = 1.0 / sd;
    const int r = n - j - 1;
    // column j of L goes straight to global memory; shared column j stays unscaled for this step's readers
    // NB: like the reference, the stored diagonal is (unregularised a_jj) * recip(l_jj)
    // (ldlt/factor.rs:161-175 scales the whole column, diagonal included, and `diag` is a local copy).
    for (int i = j + tid; i < n; i += POTF2_THREADS) A[i * rs + j * cs] = s[i * ld + j] * inv;
    // trailing update, lower part only: a_ic <- fma(-l_cj, l_ij, a_ic)
    // (2-D thread map: 32 lanes along c, POTF2_THREADS/32 rows; no integer division in the hot loop)
    (void)r;
    for (int i = j + 1 + (tid >> 5); i < n; i += POTF2_THREADS / 32) {
      const double lij = s[i * ld + j] * inv;
      for (int c = j + 1 + (tid & 31); c <= i; c += 32) {
        const double lcj = s[c * ld + j] * inv;
        s[i * ld + c] = fma(-lcj, lij, s[i * ld + c]);
      }
    }
    __syncthreads();
  }
  if (tid == 0 && count) info[1] += count;
}

struct LltCtx {
  cudaStream_t stream;
  int regularize;
  double eps, delta;
  long long* d_info;
  i64 nb;  // leaf (diagonal block) size, <= POTF2_MAX
};

// Recursive blocked LLT. Same dataflow as the reference's right-looking recursion (factor A00, solve the panel,
// update the trailing lower triangle, continue) but split in HALVES instead of fixed 128-wide steps, so that
// almost all flops are DMMA GEMMs with a large contracted dimension (k = n/2, n/4, ...): the trailing matrix is
// read/written O(log n) times instead of n/128 times. Leaves (<= nb) are the single-CTA potf2 kernel.
void llt_rec(const LltCtx& ctx, VD A, i64 j0) {
  const i64 n = A.nrows;
  if (n <= ctx.nb) {
    potf2_kernel<<<1, POTF2_THREADS, (size_t)n * (n + 1) * sizeof(double), ctx.stream>>>(
        A.ptr, A.rs, A.cs, (int)n, j0, ctx.regularize, ctx.eps, ctx.delta, ctx.d_info);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  // split at a multiple of the leaf size closest to n/2
  i64 n1 = ((n / 2 + ctx.nb - 1) / ctx.nb) * ctx.nb;
  if (n1 >= n) n1 = ((n - 1) / ctx.nb) * ctx.nb;
  const i64 n2 = n - n1;
  VD A11 = A.sub(0, 0, n1, n1), A21 = A.sub(n1, 0, n2, n1), A22 = A.sub(n1, n1, n2, n2);
  llt_rec(ctx, A11, j0);
  // conj(L11) X = A21^T   (reference ldlt/factor.rs:421-426)
  solve_lower_triangular_in_place_f64(ctx.stream, cv(A11), false, A21.t());
  // A22(lower) += -1 * A21 * A21^H   (reference ldlt/factor.rs:435-446)
  gemm_f64(ctx.stream, A22, TRI_LOWER, 1, cv(A21), RECT, cv(A21).t(), RECT, -1.0);
  llt_rec(ctx, A22, j0 + n1);
}

}  // namespace

LltResult llt_cholesky_in_place_f64(cudaStream_t stream, VD A, double reg_delta, double reg_eps, LltParams params) {
  FB_ASSERT(A.nrows == A.ncols, "LLT needs a square matrix");
  const i64 n = A.nrows;
  LltResult res{true, 0, 0};
  if (n == 0) return res;
  const int regularize = (reg_delta > 0.0 && reg_eps > 0.0) ? 1 : 0;
  i64 nb = (i64)params.block_size;
  if (nb <= 0 || nb > POTF2_MAX) nb = POTF2_MAX;

  long long* d_info = (long long*)ws_alloc(2 * sizeof(long long));
  long long h_info[2] = {-1, 0};
  FB_CUDA_CHECK(cudaMemcpyAsync(d_info, h_info, sizeof(h_info), cudaMemcpyHostToDevice, stream));

  static bool configured = false;
  if (!configured) {
    FB_CUDA_CHECK(cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(POTF2_MAX * (POTF2_MAX + 1) * sizeof(double))));
    configured = true;
  }

  LltCtx ctx{stream, regularize, reg_eps, reg_delta, d_info, nb};
  llt_rec(ctx, A, 0);
  FB_CUDA_CHECK(cudaMemcpyAsync(h_info, d_info, sizeof(h_info), cudaMemcpyDeviceToHost, stream));
  FB_CUDA_CHECK(cudaStreamSynchronize(stream));
  ws_free(d_info);
  if (h_info[0] >= 0) {
    res.ok = false;
    res.non_positive_pivot_index = (size_t)h_info[0];
  } else {
    res.dynamic_regularization_count = (size_t)h_info[1];
  }
  return res;
}

}  // namespace fb
