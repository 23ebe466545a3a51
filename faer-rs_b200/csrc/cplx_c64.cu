// c64 (complex<f64>) triangular solves and Cholesky LLT: faer's recursions on the complex GEMM (gemm_c64.cu) with two small
// complex leaf kernels.
//
// Reference:
//   triangular_solve::solve_[unit_]{lower,upper}_triangular_in_place_with_conj   faer/src/linalg/triangular_solve.rs:220-604
//       recursive split (block_size 200-211): top solve, rhs_bot -= conj?(T10) * rhs_top, bottom solve; leaves multiply by the
//       reciprocal of the diagonal (16-198); upper = lower on reversed views (577-604)
//   cholesky::llt::factor::cholesky_in_place for complex T   cholesky/llt/factor.rs:68-97 -> ldlt/factor.rs:367-498
//       A00 = L00 L00^H (recursion / leaf), conj(L00) X = A10^T, A11(lower) -= A10 A10^H; leaf recurrence
//       a_ij <- a_ij - conj(a_jk) a_ik, d = Re(a_jj), [regularise], fail if !(d > 0), l = sqrt(d), column j (diagonal included)
//       multiplied by recip(l)  (ldlt/factor.rs:7-177, 299-366)
//
// All views are in COMPLEX element units (pointer to the first complex element as double*, strides in complex elements), as in
// gemm_c64.cu. The leaves are plain scalar kernels (a 32 x 32 block per CTA / one thread per right-hand-side column): O(n^2 leaf)
// work next to the O(n^3) that runs on the DMMA GEMM.
#include <algorithm>
#include <vector>

#include "linalg_f64.cuh"
#include "runtime.cuh"

namespace fb {

namespace {

constexpr int CL = 32;  // leaf order of both recursions

struct Cx {
  double re, im;
};
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return Cx{fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)}; }
__device__ __forceinline__ Cx cld(const double* p, i64 off) { return Cx{p[2 * off], p[2 * off + 1]}; }
__device__ __forceinline__ void cst(double* p, i64 off, Cx v) {
  p[2 * off] = v.re;
  p[2 * off + 1] = v.im;
}

// conj?(T) X = B for a lower-triangular leaf T (n <= CL): one thread per column of B, the column in registers
__global__ void __launch_bounds__(64) trsm_leaf_lower_c64_kernel(const double* __restrict__ T, i64 t_rs, i64 t_cs, int n, int unit,
                                                                  int conj, double* __restrict__ B, i64 b_rs, i64 b_cs, i64 ncols) {
  __shared__ double nlr[CL][CL + 1], nli[CL][CL + 1];  // -(conj? l_ik) * inv_i below the diagonal, inv_i on it
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, k = e - i * n;
    if (k > i) continue;
    Cx inv{1.0, 0.0};
    if (!unit) {
      const Cx d = cld(T, (i64)i * t_rs + (i64)i * t_cs);
      const double s = 1.0 / fma(d.re, d.re, d.im * d.im);
      inv = Cx{d.re * s, -d.im * s};        // 1 / d
      if (conj) inv.im = -inv.im;           // conj(1 / d) = 1 / conj(d)
    }
    Cx v = inv;
    if (k < i) {
      Cx l = cld(T, (i64)i * t_rs + (i64)k * t_cs);
      if (conj) l.im = -l.im;
      v = cmul(Cx{-l.re, -l.im}, inv);
    }
    nlr[i][k] = v.re;
    nli[i][k] = v.im;
  }
  __syncthreads();
  const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double yr[CL], yi[CL];
#pragma unroll
  for (int i = 0; i < CL; ++i) {
    if (i < n) {
      Cx v = cld(B, (i64)i * b_rs + col * b_cs);
      v = cmul(v, Cx{nlr[i][i], nli[i][i]});  // * inv_i (1 for a unit diagonal)
#pragma unroll
      for (int k = 0; k < CL; ++k) {
        if (k < i) {
          const Cx t = cmul(Cx{nlr[i][k], nli[i][k]}, Cx{yr[k], yi[k]});
          v.re += t.re;
          v.im += t.im;
        }
      }
      yr[i] = v.re;
      yi[i] = v.im;
      cst(B, (i64)i * b_rs + col * b_cs, v);
    }
  }
}

// in-place lower Cholesky of a <= CL x CL Hermitian block (lower triangle read / written), one CTA, block in shared memory
__global__ void __launch_bounds__(CL * CL) potf2_c64_kernel(double* __restrict__ A, i64 rs, i64 cs, int n, i64 j0, int regularize,
                                                            double eps, double delta, long long* __restrict__ info) {
  __shared__ double Sr[CL][CL + 1], Si[CL][CL + 1];
  __shared__ double s_inv;
  __shared__ int s_fail, s_count;
  if (info[0] >= 0) return;  // an earlier block already failed
  const int i = threadIdx.x % CL, c = threadIdx.x / CL;
  const bool on = i < n && c <= i;
  if (on) {
    const Cx v = cld(A, (i64)i * rs + (i64)c * cs);
    Sr[i][c] = v.re;
    Si[i][c] = v.im;
  }
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      double d = Sr[j][j];
      int fail = 0;
      if (regularize && d <= eps) {
        d = delta;
        s_count += 1;
      }
      double inv = 0.0;
      if (!(d > 0.0)) fail = 1;
      else {
        const double sd = sqrt(d);
        if (sd == 0.0 || !isfinite(sd)) fail = 1;
        else inv = 1.0 / sd;
      }
      s_inv = inv;
      s_fail = fail;
    }
    __syncthreads();
    if (s_fail) {
      if (threadIdx.x == 0) info[0] = j0 + j;
      return;
    }
    const double inv = s_inv;
    // column j (diagonal included) * recip(l_jj)
    if (c == j && i >= j && i < n) {
      Sr[i][j] *= inv;
      Si[i][j] *= inv;
    }
    __syncthreads();
    // a_ic <- a_ic - conj(l_cj) l_ij for j < c <= i
    if (on && c > j) {
      const Cx lc{Sr[c][j], -Si[c][j]}, li{Sr[i][j], Si[i][j]};
      const Cx t = cmul(lc, li);
      Sr[i][c] -= t.re;
      Si[i][c] -= t.im;
    }
    __syncthreads();
  }
  if (on) cst(A, (i64)i * rs + (i64)c * cs, Cx{Sr[i][c], Si[i][c]});
  if (threadIdx.x == 0 && s_count) info[1] += s_count;
}

inline VD csub(VD v, i64 i, i64 j, i64 m, i64 n) { return VD{v.ptr + 2 * (i * v.rs + j * v.cs), m, n, v.rs, v.cs}; }
inline VCD csub(VCD v, i64 i, i64 j, i64 m, i64 n) { return VCD{v.ptr + 2 * (i * v.rs + j * v.cs), m, n, v.rs, v.cs}; }

void solve_lower_rec_c64(cudaStream_t st, VCD T, bool unit, bool conj, VD rhs) {
  const i64 n = T.nrows, k = rhs.ncols;
  if (n == 0 || k == 0) return;
  if (n <= CL) {
    trsm_leaf_lower_c64_kernel<<<(unsigned)((k + 63) / 64), 64, 0, st>>>(T.ptr, T.rs, T.cs, (int)n, unit ? 1 : 0, conj ? 1 : 0, rhs.ptr,
                                                                         rhs.rs, rhs.cs, k);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  const i64 n1 = ((n / 2 + CL - 1) / CL) * CL;
  solve_lower_rec_c64(st, csub(T, 0, 0, n1, n1), unit, conj, csub(rhs, 0, 0, n1, k));
  // rhs_bot -= conj?(T10) * rhs_top
  gemm_c64(st, csub(rhs, n1, 0, n - n1, k), RECT, 1, csub(T, n1, 0, n - n1, n1), RECT, conj, cv(csub(rhs, 0, 0, n1, k)), RECT, false,
           -1.0, 0.0);
  solve_lower_rec_c64(st, csub(T, n1, n1, n - n1, n - n1), unit, conj, csub(rhs, n1, 0, n - n1, k));
}

struct LltCtxC {
  cudaStream_t st;
  int regularize;
  double eps, delta;
  long long* d_info;
};

void llt_rec_c64(const LltCtxC& ctx, VD A, i64 j0) {
  const i64 n = A.nrows;
  if (n <= CL) {
    potf2_c64_kernel<<<1, CL * CL, 0, ctx.st>>>(A.ptr, A.rs, A.cs, (int)n, j0, ctx.regularize, ctx.eps, ctx.delta, ctx.d_info);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  const i64 n1 = ((n / 2 + CL - 1) / CL) * CL, n2 = n - n1;
  VD A11 = csub(A, 0, 0, n1, n1), A21 = csub(A, n1, 0, n2, n1), A22 = csub(A, n1, n1, n2, n2);
  llt_rec_c64(ctx, A11, j0);
  // conj(L11) X = A21^T   (ldlt/factor.rs:421-426)
  solve_lower_rec_c64(ctx.st, cv(A11), false, true, A21.t());
  // A22(lower) -= A21 A21^H   (435-446)
  gemm_c64(ctx.st, A22, TRI_LOWER, 1, cv(A21), RECT, false, cv(A21).t(), RECT, true, -1.0, 0.0);
  llt_rec_c64(ctx, A22, j0 + n1);
}


// ---- partial-pivoting LU (lu/partial_pivoting/factor.rs:19-295 for complex T) ---------------------------------------------
constexpr int LU_CW = 16;         // leaf width
constexpr int LU_CT = 1024;       // leaf threads (one CTA walks the panel's rows in global memory)

// Unblocked leaf on the view A (m rows, `ncols_view` columns; local row 0 = first diagonal row of the window [start, start + w)):
// pivot = first row attaining the largest abs1 = |re| + |im| (strict `>` from 0: an all-zero column keeps the diagonal row),
// the swap covers the whole row of the view (factor.rs:46), multipliers by reciprocal-multiply, rank-1 update of the window.
__global__ void __launch_bounds__(LU_CT) lu_leaf_c64_kernel(double* __restrict__ A, i64 rs, i64 cs, int m, int ncols_view, int start,
                                                            int w, int* __restrict__ trans) {
  __shared__ double red_v[LU_CT / 32];
  __shared__ int red_i[LU_CT / 32];
  __shared__ int s_piv;
  __shared__ double s_inv[2];
  __shared__ double s_row[LU_CW][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = 0; j < w && j < m; ++j) {
    const int col = start + j;
    // ---- pivot search over rows >= j ----
    double bv = 0.0;
    int bi = j;
    for (int i = j + tid; i < m; i += LU_CT) {
      const Cx a = cld(A, (i64)i * rs + (i64)col * cs);
      const double v = fabs(a.re) + fabs(a.im);
      if (v > bv) {  // rows ascend within a thread: the first maximum is kept
        bv = v;
        bi = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      red_v[warp] = bv;
      red_i[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
      bv = red_v[lane];
      bi = red_i[lane];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        s_piv = (bv > 0.0) ? bi : j;
        trans[j] = s_piv - j;
      }
    }
    __syncthreads();
    const int piv = s_piv;
    // ---- swap rows j <-> piv over the whole view ----
    if (piv != j) {
      for (int c = tid; c < ncols_view; c += LU_CT) {
        const Cx a = cld(A, (i64)j * rs + (i64)c * cs), b = cld(A, (i64)piv * rs + (i64)c * cs);
        cst(A, (i64)j * rs + (i64)c * cs, b);
        cst(A, (i64)piv * rs + (i64)c * cs, a);
      }
    }
    __syncthreads();
    if (tid == 0) {
      const Cx d = cld(A, (i64)j * rs + (i64)col * cs);
      const double sc = 1.0 / fma(d.re, d.re, d.im * d.im);
      s_inv[0] = d.re * sc;
      s_inv[1] = -d.im * sc;
    }
    if (tid < w - j - 1) {
      const Cx u = cld(A, (i64)j * rs + (i64)(col + 1 + tid) * cs);
      s_row[tid][0] = u.re;
      s_row[tid][1] = u.im;
    }
    __syncthreads();
    const Cx inv{s_inv[0], s_inv[1]};
    // ---- multipliers and rank-1 update: one row per thread and iteration ----
    for (int i = j + 1 + tid; i < m; i += LU_CT) {
      const Cx l = cmul(cld(A, (i64)i * rs + (i64)col * cs), inv);
      cst(A, (i64)i * rs + (i64)col * cs, l);
      for (int c = 0; c < w - j - 1; ++c) {
        const Cx t = cmul(l, Cx{s_row[c][0], s_row[c][1]});
        Cx a = cld(A, (i64)i * rs + (i64)(col + 1 + c) * cs);
        a.re -= t.re;
        a.im -= t.im;
        cst(A, (i64)i * rs + (i64)(col + 1 + c) * cs, a);
      }
    }
    __syncthreads();
  }
}

// apply n transpositions (row j <-> row j + trans[j], in order) to every column of the view: one thread per column
__global__ void laswp_c64_kernel(double* __restrict__ A, i64 rs, i64 cs, i64 ncols, const int* __restrict__ trans, int n) {
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  for (int j = 0; j < n; ++j) {
    const int t = trans[j];
    if (t == 0) continue;
    const Cx a = cld(A, (i64)j * rs + c * cs), b = cld(A, (i64)(j + t) * rs + c * cs);
    cst(A, (i64)j * rs + c * cs, b);
    cst(A, (i64)(j + t) * rs + c * cs, a);
  }
}

inline i64 next_pow2_c(i64 n) {
  i64 p = 1;
  while (p < n) p <<= 1;
  return p;
}

// the reference's recursion (factor.rs:68-187): A = current view (all m rows, ncols columns), window [start, end)
void lu_rec_c64(cudaStream_t st, VD A, i64 start, i64 end, int* trans) {
  const i64 m = A.nrows, ncols = A.ncols, n = end - start;
  if (n == 0) return;
  if (n <= LU_CW) {
    lu_leaf_c64_kernel<<<1, LU_CT, 0, st>>>(A.ptr, A.rs, A.cs, (int)m, (int)ncols, (int)start, (int)n, trans);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  const i64 half = n / 2;
  const i64 pw = std::min<i64>(16, next_pow2_c(half));
  const i64 bs = (half + pw - 1) / pw * pw;
  VD W = csub(A, 0, start, m, n);
  lu_rec_c64(st, W, 0, bs, trans);
  {
    VD A00 = csub(W, 0, 0, bs, bs), A01 = csub(W, 0, bs, bs, n - bs), A10 = csub(W, bs, 0, m - bs, bs), A11 = csub(W, bs, bs, m - bs, n - bs);
    solve_lower_rec_c64(st, cv(A00), true, false, A01);
    gemm_c64(st, A11, RECT, 1, cv(A10), RECT, false, cv(A01), RECT, false, -1.0, 0.0);
    lu_rec_c64(st, csub(W, bs, 0, m - bs, n), bs, n, trans + bs);
  }
  auto swap_cols = [&](VD M) {
    if (M.ncols == 0) return;
    laswp_c64_kernel<<<(unsigned)((M.ncols + 127) / 128), 128, 0, st>>>(M.ptr, M.rs, M.cs, M.ncols, trans, (int)n);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
  };
  swap_cols(csub(A, 0, 0, m, start));
  swap_cols(csub(A, 0, end, m, ncols - end));
}

// dst (compact column-major complex, ld = nrows) [i, c] = src[perm[i], c]
__global__ void gather_rows_c64_kernel(double* __restrict__ dst, const double* __restrict__ src, i64 rs, i64 cs, i64 nrows,
                                       const long long* __restrict__ perm) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 c = blockIdx.y;
  if (i < nrows) cst(dst, c * nrows + i, cld(src, perm[i] * rs + c * cs));
}
__global__ void scatter_rows_c64_kernel(double* __restrict__ dst, i64 rs, i64 cs, const double* __restrict__ src, i64 nrows) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 c = blockIdx.y;
  if (i < nrows) cst(dst, i * rs + c * cs, cld(src, c * nrows + i));
}

}  // namespace

// conj?(T) X = rhs, T lower / upper triangular (unit diagonal or not); views in complex units
void solve_lower_triangular_in_place_c64(cudaStream_t st, VCD tril, bool unit, bool conj, VD rhs) {
  FB_ASSERT(tril.nrows == tril.ncols && rhs.nrows == tril.nrows, "triangular solve shape mismatch");
  solve_lower_rec_c64(st, tril, unit, conj, rhs);
}
void solve_upper_triangular_in_place_c64(cudaStream_t st, VCD triu, bool unit, bool conj, VD rhs) {
  FB_ASSERT(triu.nrows == triu.ncols && rhs.nrows == triu.nrows, "triangular solve shape mismatch");
  const i64 n = triu.nrows;
  if (n == 0 || rhs.ncols == 0) return;
  // upper = lower on the views with rows and columns reversed (triangular_solve.rs:577-604)
  VCD t{triu.ptr + 2 * ((n - 1) * triu.rs + (n - 1) * triu.cs), n, n, -triu.rs, -triu.cs};
  VD r{rhs.ptr + 2 * ((n - 1) * rhs.rs), n, rhs.ncols, -rhs.rs, rhs.cs};
  solve_lower_rec_c64(st, t, unit, conj, r);
}

LltResult llt_cholesky_in_place_c64(cudaStream_t st, VD A, double reg_delta, double reg_eps) {
  FB_ASSERT(A.nrows == A.ncols, "LLT needs a square matrix");
  LltResult res{true, 0, 0};
  if (A.nrows == 0) return res;
  long long* d_info = (long long*)ws_alloc(2 * sizeof(long long));
  long long h_info[2] = {-1, 0};
  FB_CUDA_CHECK(cudaMemcpyAsync(d_info, h_info, sizeof(h_info), cudaMemcpyHostToDevice, st));
  LltCtxC ctx{st, (reg_delta > 0.0 && reg_eps > 0.0) ? 1 : 0, reg_eps, reg_delta, d_info};
  llt_rec_c64(ctx, A, 0);
  FB_CUDA_CHECK(cudaMemcpyAsync(h_info, d_info, sizeof(h_info), cudaMemcpyDeviceToHost, st));
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  ws_free(d_info);
  if (h_info[0] >= 0) {
    res.ok = false;
    res.non_positive_pivot_index = (size_t)h_info[0];
  } else {
    res.dynamic_regularization_count = (size_t)h_info[1];
  }
  return res;
}

// L y = b, then L^H x = y (cholesky/llt/solve.rs:12-35 for complex L)
void llt_solve_in_place_c64(cudaStream_t st, VCD L, bool conj, VD rhs) {
  FB_ASSERT(L.nrows == L.ncols && rhs.nrows == L.nrows, "LLT solve shape mismatch");
  solve_lower_triangular_in_place_c64(st, L, false, conj, rhs);
  solve_upper_triangular_in_place_c64(st, L.t(), false, !conj, rhs);
}

// In-place P A = L U of an m x n c64 matrix (views in complex units); perm arrays: HOST int64 of length m. Returns the
// transposition count (lu_in_place, factor.rs:234-295).
size_t lu_partial_piv_in_place_c64(cudaStream_t st, VD A, long long* perm_fwd, long long* perm_inv) {
  const i64 m = A.nrows, n = A.ncols, size = std::min(m, n);
  for (i64 i = 0; i < m; ++i) perm_fwd[i] = i;
  size_t n_trans = 0;
  if (size > 0) {
    int* d_trans = (int*)ws_alloc((size_t)size * sizeof(int));
    FB_CUDA_CHECK(cudaMemsetAsync(d_trans, 0, (size_t)size * sizeof(int), st));
    lu_rec_c64(st, A, 0, size, d_trans);
    std::vector<int> h_trans((size_t)size);
    FB_CUDA_CHECK(cudaMemcpyAsync(h_trans.data(), d_trans, (size_t)size * sizeof(int), cudaMemcpyDeviceToHost, st));
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    ws_free(d_trans);
    for (i64 i = 0; i < size; ++i) {
      const int t = h_trans[(size_t)i];
      if (t != 0) {
        std::swap(perm_fwd[i], perm_fwd[i + t]);
        ++n_trans;
      }
    }
    if (m < n) {  // factor.rs:278-285
      solve_lower_rec_c64(st, cv(csub(A, 0, 0, m, size)), true, false, csub(A, 0, size, m, n - size));
      FB_CUDA_CHECK(cudaStreamSynchronize(st));
    }
  }
  for (i64 i = 0; i < m; ++i) perm_inv[perm_fwd[i]] = i;
  return n_trans;
}

// rhs <- conj?(A)^-1 rhs from the factors (lu/partial_pivoting/solve.rs:21-54): permute rows, unit-lower solve, upper solve
void lu_solve_in_place_c64(cudaStream_t st, VCD L, VCD U, bool conj, const long long* perm_fwd, VD rhs) {
  const i64 n = L.nrows, k = rhs.ncols;
  FB_ASSERT(L.ncols == n && U.nrows == n && U.ncols == n && rhs.nrows == n, "LU solve shape mismatch");
  if (n == 0 || k == 0) return;
  FB_ASSERT(k < 65536, "too many right-hand sides for one permutation launch");
  long long* d_perm = (long long*)ws_alloc((size_t)n * 8);
  double* tmp = (double*)ws_alloc((size_t)n * (size_t)k * 16);
  FB_CUDA_CHECK(cudaMemcpyAsync(d_perm, perm_fwd, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)k);
  gather_rows_c64_kernel<<<grid, 256, 0, st>>>(tmp, rhs.ptr, rhs.rs, rhs.cs, n, d_perm);
  scatter_rows_c64_kernel<<<grid, 256, 0, st>>>(rhs.ptr, rhs.rs, rhs.cs, tmp, n);
  FB_CUDA_CHECK(cudaGetLastError());
  note_launch();
  note_launch();
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  ws_free(tmp);
  ws_free(d_perm);
  solve_lower_triangular_in_place_c64(st, L, true, conj, rhs);
  solve_upper_triangular_in_place_c64(st, U, false, conj, rhs);
}

}  // namespace fb
