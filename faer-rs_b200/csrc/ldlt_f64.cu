// LDLT without pivoting, f64: leaf kernel + recursive driver + the solve on the factors.
// STATUS: validated on hardware against the CPU checker's LDLT (tests/test_gpu_zz6_ldlt.py; the leaf recurrence is also emulated on
// the CPU in tests/test_ldlf2_emulation_cpu.py).
//
// Reference: faer/src/linalg/cholesky/ldlt/factor.rs
//   cholesky_in_place 725-767 -> cholesky_recursion_right_looking(is_llt = false) 367-498:
//       factor A00; conj(A00) X = A10^T with the UNIT-lower solve (X = L10 D0); L10 = X * recip(D0) column by column;
//       A11(lower) -= L10 X^H  (portable branch 471-492; the x86 branch gives the same product through `spicy_matmul`, 447-470).
//       D ends on the diagonal of A (757-765), the strict upper triangle is untouched.
//   leaf 7-177: a_ic <- fma(l_cj * (-d_j), l_ij, a_ic) for the earlier columns j in order; d_c = Re(a_cc); dynamic
//       regularisation 122-144 (expected signs); d == 0 or non-finite -> ZeroPivot(c); column c *= recip(d_c).
//   solve.rs:11-49: unit-lower solve with L, rows scaled by recip(d_i), unit-upper solve with L^H.
//
// B200 mapping: the same shape as the LLT path (llt.cu, left untouched): one CTA factors a <= 128-wide diagonal block
// with the block in registers (16 update warps) while 4 pivot warps keep a bit-identical copy of the diagonal and do the
// serial pivot arithmetic of column j + 1 during the update of column j; the host recursion splits in halves so that the
// flops are DMMA GEMMs with a large contracted dimension. The only extra traffic against LLT is the copy of the solved
// panel X (one read + one write of the panel) needed because L10 and X = L10 D0 both enter the trailing update.
#include "linalg_f64.cuh"

namespace fb {

namespace {

constexpr int LDL_MAX = 128;
constexpr int LDL_UPD_WARPS = 16;
constexpr int LDL_THREADS = LDL_UPD_WARPS * 32 + LDL_MAX;  // 512 update threads + 128 pivot threads
constexpr int LDL_CB = LDL_MAX / LDL_UPD_WARPS;            // column slots per update thread (8)

// info[0]: first failing global column (or -1), info[1]: regularisation count. signs: device int8, indexed by the
// GLOBAL column (j0 + local), or null.
__global__ void __launch_bounds__(LDL_THREADS) ldlf2_kernel(double* __restrict__ A, i64 rs, i64 cs, int n, i64 j0,
                                                             int regularize, double eps, double delta,
                                                             const signed char* __restrict__ signs,
                                                             long long* __restrict__ info) {
  __shared__ double colbuf[2][LDL_MAX];
  __shared__ double s_inv[2];
  __shared__ double s_d[2];
  __shared__ int s_fail[2];
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_upd = warp < LDL_UPD_WARPS;
  const int p = tid - LDL_UPD_WARPS * 32;  // pivot-thread index (diagonal entry p) when !is_upd
  if (info[0] >= 0) return;  // an earlier block already failed (uniform across the CTA)
  if (tid == 0) s_count = 0;

  // update threads: rows i = lane + 32a, columns c = warp + 16b, kept iff c <= i < n
  double a[4][LDL_CB];
  double dp = 0.0;  // pivot threads: diagonal entry p
  if (is_upd) {
#pragma unroll
    for (int ai = 0; ai < 4; ++ai)
#pragma unroll
      for (int bi = 0; bi < LDL_CB; ++bi) {
        const int i = lane + 32 * ai, c = warp + LDL_UPD_WARPS * bi;
        a[ai][bi] = (i < n && c <= i) ? A[(i64)i * rs + (i64)c * cs] : 0.0;
      }
  } else if (p < n) {
    dp = A[(i64)p * rs + (i64)p * cs];
  }
  __syncthreads();  // s_count initialised

  // serial pivot arithmetic of column jc (ldlt/factor.rs:122-160, is_llt = false), done by ONE pivot thread
  auto publish_pivot = [&](int jc, double d) {
    if (regularize) {
      const int sign = signs ? (int)signs[j0 + jc] : 0;
      const bool small_or_negative = d <= eps;
      const bool minus_small_or_positive = d >= -eps;
      if (sign == 1 && small_or_negative) {
        d = delta;
        s_count += 1;  // single writer per column, ordered by the per-column barrier; only this case is counted
      } else if (sign == -1 && minus_small_or_positive) {
        d = -delta;
      } else if (small_or_negative && minus_small_or_positive) {
        d = d < 0.0 ? -delta : delta;
      }
    }
    const int fail = (d == 0.0 || !isfinite(d)) ? 1 : 0;
    s_d[jc & 1] = d;
    s_inv[jc & 1] = fail ? 0.0 : 1.0 / d;
    s_fail[jc & 1] = fail;
  };

  if (is_upd) {
    if (warp == 0) {
#pragma unroll
      for (int ai = 0; ai < 4; ++ai) colbuf[0][lane + 32 * ai] = a[ai][0];
    }
  } else if (p == 0) {
    publish_pivot(0, dp);
  }
  __syncthreads();

  for (int j = 0; j < n; ++j) {
    const double* col = colbuf[j & 1];
    const double d = s_d[j & 1];
    if (s_fail[j & 1]) {
      if (tid == 0) {
        info[0] = j0 + j;
        A[(i64)j * rs + (i64)j * cs] = d;  // the diagonal is initialised up to and including the failing column (757-765)
      }
      return;
    }
    const double inv = s_inv[j & 1];
    const double nd = -d;
    if (!is_upd) {
      // pivot group: the SAME fma the update threads apply to a_pp, then the next column's pivot arithmetic
      if (p > j && p < n) {
        const double l = col[p] * inv;
        dp = fma(l * nd, l, dp);
        if (p == j + 1) publish_pivot(j + 1, dp);
      }
    } else {
      const int jw = j & (LDL_UPD_WARPS - 1);
      // column j goes to global memory (owners: warp jw): L below the diagonal, D on it
      if (warp == jw) {
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) {
          const int i = lane + 32 * ai;
          if (i > j && i < n) A[(i64)i * rs + (i64)j * cs] = col[i] * inv;
          else if (i == j) A[(i64)i * rs + (i64)j * cs] = d;
        }
      }
      // trailing update: a_ic <- fma(l_cj * (-d_j), l_ij, a_ic) for j < c <= i (warp-uniform column test, see potf2_kernel)
      if (warp + LDL_UPD_WARPS * (LDL_CB - 1) > j) {
        double li[4];
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) li[ai] = col[lane + 32 * ai] * inv;
#pragma unroll
        for (int bi = 0; bi < LDL_CB; ++bi) {
          const int c = warp + LDL_UPD_WARPS * bi;
          if (c > j && c < n) {
            const double lcd = (col[c] * inv) * nd;
#pragma unroll
            for (int ai = 0; ai < 4; ++ai) {
              const int i = lane + 32 * ai;
              if (32 * ai + 31 >= c) {  // warp-uniform: this row slot intersects i >= c
                if (i >= c && i < n) a[ai][bi] = fma(lcd, li[ai], a[ai][bi]);
              }
            }
          }
        }
      }
      // owners of column j+1 publish it (unscaled) into the other buffer
      if (j + 1 < n && warp == ((j + 1) & (LDL_UPD_WARPS - 1))) {
        const int nbk = (j + 1) / LDL_UPD_WARPS;
        double* nxt = colbuf[(j + 1) & 1];
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) {
          double v = 0.0;
#pragma unroll
          for (int bi = 0; bi < LDL_CB; ++bi)
            if (bi == nbk) v = a[ai][bi];
          nxt[lane + 32 * ai] = v;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0 && s_count) info[1] += s_count;
}

// A21(:, k) <- A21(:, k) * recip(d_k),  d_k = Dblk[k * dstride]
__global__ void ldlt_scale_kernel(double* __restrict__ A21, i64 rs, i64 cs, i64 n2, i64 n1, const double* __restrict__ Dblk,
                                  i64 dstride) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 k = blockIdx.y;
  if (i >= n2 || k >= n1) return;
  A21[i * rs + k * cs] *= 1.0 / Dblk[k * dstride];
}

// rhs(i, :) <- rhs(i, :) * recip(d_i)
__global__ void ldlt_row_scale_kernel(double* __restrict__ R, i64 rs, i64 cs, i64 n, i64 k, const double* __restrict__ D,
                                      i64 dstride) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 c = blockIdx.y;
  if (i >= n || c >= k) return;
  R[i * rs + c * cs] *= 1.0 / D[i * dstride];
}

struct LdltCtx {
  cudaStream_t stream;
  int regularize;
  double eps, delta;
  const signed char* d_signs;
  long long* d_info;
  i64 nb;  // leaf (diagonal block) size, <= LDL_MAX
};

void ldlt_rec(const LdltCtx& ctx, VD A, i64 j0) {
  const i64 n = A.nrows;
  if (n <= ctx.nb) {
    ldlf2_kernel<<<1, LDL_THREADS, 0, ctx.stream>>>(A.ptr, A.rs, A.cs, (int)n, j0, ctx.regularize, ctx.eps, ctx.delta,
                                                    ctx.d_signs, ctx.d_info);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  i64 n1 = ((n / 2 + ctx.nb - 1) / ctx.nb) * ctx.nb;
  if (n1 >= n) n1 = ((n - 1) / ctx.nb) * ctx.nb;
  const i64 n2 = n - n1;
  VD A11 = A.sub(0, 0, n1, n1), A21 = A.sub(n1, 0, n2, n1), A22 = A.sub(n1, n1, n2, n2);
  ldlt_rec(ctx, A11, j0);
  // conj(L11) X = A21^T with the unit-lower solve: X = L21 D1  (ldlt/factor.rs:427-433)
  solve_lower_triangular_in_place_f64(ctx.stream, cv(A11), true, A21.t());
  // A21 <- X * recip(D1) = L21 in place, then A22(lower) -= L21 D1 L21^H as ONE "spicy" product with the diagonal folded
  // into the lhs fragments (ldlt/factor.rs:447-470, the reference's has_spicy_matmul branch; the copy of X that its other
  // branch keeps, 471-492, is not needed)
  FB_ASSERT(n1 < 65536, "LDLT block too wide for one scale launch");
  dim3 grid((unsigned)((n2 + 255) / 256), (unsigned)n1);
  ldlt_scale_kernel<<<grid, 256, 0, ctx.stream>>>(A21.ptr, A21.rs, A21.cs, n2, n1, A11.ptr, A11.rs + A11.cs);
  FB_CUDA_CHECK(cudaGetLastError());
  note_launch();
  spicy_matmul_f64(ctx.stream, A22, TRI_LOWER, nullptr, nullptr, 1, cv(A21), cv(A21).t(), A11.ptr, A11.rs + A11.cs, -1.0);
  ldlt_rec(ctx, A22, j0 + n1);
}

}  // namespace

LdltResult ldlt_in_place_f64(cudaStream_t stream, VD A, double reg_delta, double reg_eps, const signed char* d_signs,
                             LltParams params) {
  FB_ASSERT(A.nrows == A.ncols, "LDLT needs a square matrix");
  const i64 n = A.nrows;
  LdltResult res{true, 0, 0};
  if (n == 0) return res;
  const int regularize = (reg_delta > 0.0 && reg_eps > 0.0) ? 1 : 0;  // ldlt/factor.rs:744-745
  i64 nb = (i64)params.block_size;
  if (nb <= 0 || nb > LDL_MAX) nb = LDL_MAX;
  long long* d_info = (long long*)ws_alloc(2 * sizeof(long long));
  long long h_info[2] = {-1, 0};
  FB_CUDA_CHECK(cudaMemcpyAsync(d_info, h_info, sizeof(h_info), cudaMemcpyHostToDevice, stream));
  LdltCtx ctx{stream, regularize, reg_eps, reg_delta, d_signs, d_info, nb};
  ldlt_rec(ctx, A, 0);
  FB_CUDA_CHECK(cudaMemcpyAsync(h_info, d_info, sizeof(h_info), cudaMemcpyDeviceToHost, stream));
  FB_CUDA_CHECK(cudaStreamSynchronize(stream));
  ws_free(d_info);
  if (h_info[0] >= 0) {
    res.ok = false;
    res.zero_pivot_index = (size_t)h_info[0];
  } else {
    res.dynamic_regularization_count = (size_t)h_info[1];
  }
  return res;
}

// ldlt/solve.rs:11-49 for real scalars. D: device pointer, `dstride` elements between consecutive entries.
void ldlt_solve_in_place_f64(cudaStream_t stream, VCD L, const double* D, i64 dstride, VD rhs) {
  const i64 n = L.nrows;
  FB_ASSERT(L.ncols == n && rhs.nrows == n, "LDLT solve shape mismatch");
  if (n == 0 || rhs.ncols == 0) return;
  solve_lower_triangular_in_place_f64(stream, L, true, rhs);
  FB_ASSERT(rhs.ncols < 65536, "too many right-hand sides for one scaling launch");
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)rhs.ncols);
  ldlt_row_scale_kernel<<<grid, 256, 0, stream>>>(rhs.ptr, rhs.rs, rhs.cs, n, rhs.ncols, D, dstride);
  FB_CUDA_CHECK(cudaGetLastError());
  note_launch();
  solve_upper_triangular_in_place_f64(stream, L.t(), true, rhs);
}

}  // namespace fb
