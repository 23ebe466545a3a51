// LLT for f32: the f64 leaf kernel and recursive driver of llt_f64.cu with the scalar type changed (same thread layout,
// same recurrence: a_ic <- fmaf(-l_cj, l_ij, a_ic), regularise / test / sqrtf / reciprocal-multiply, ldlt/factor.rs:7-177),
// trailing updates on the f32 GEMM (3xTF32, tcgen05 for large products), panel solves on the f32 triangular solve.
// The look-ahead block-column drivers (dist.cu) are f64-only, so large f32 problems use the recursive driver.
// STATUS: a type-substituted copy of the validated f64 leaf and recursive driver (tests/test_gpu_zz5_llt_f32.py).
// Reference: faer/src/linalg/cholesky/llt/factor.rs:68-97 -> ldlt/factor.rs:367-498; solve: llt/solve.rs:12-35.
#include "gemm_f32.cuh"
#include "linalg_f64.cuh"

namespace fb {

namespace {

constexpr int POTF2F_MAX = 128;
constexpr int POTF2F_UPD_WARPS = 16;
constexpr int POTF2F_THREADS = POTF2F_UPD_WARPS * 32 + POTF2F_MAX;  // 512 update threads + 128 pivot threads
constexpr int POTF2F_CB = POTF2F_MAX / POTF2F_UPD_WARPS;           // column slots per update thread (8)

// info[0]: first failing global column (or -1), info[1]: regularisation count
__global__ void __launch_bounds__(POTF2F_THREADS) potf2_f32_kernel(float* __restrict__ A, i64 rs, i64 cs, int n, i64 j0,
                                                               int regularize, float eps, float delta,
                                                               long long* __restrict__ info) {
  __shared__ float colbuf[2][POTF2F_MAX];
  __shared__ float s_inv[2];
  __shared__ int s_fail[2];
  __shared__ int s_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_upd = warp < POTF2F_UPD_WARPS;
  const int p = tid - POTF2F_UPD_WARPS * 32;  // pivot-thread index (diagonal entry p) when !is_upd
  if (info[0] >= 0) return;  // an earlier block already failed (uniform across the CTA)
  if (tid == 0) s_count = 0;

  // update threads: rows i = lane + 32a, columns c = warp + 16b, kept iff c <= i < n
  float a[4][POTF2F_CB];
  float dp = 0.0f;  // pivot threads: diagonal entry p
  if (is_upd) {
#pragma unroll
    for (int ai = 0; ai < 4; ++ai)
#pragma unroll
      for (int bi = 0; bi < POTF2F_CB; ++bi) {
        const int i = lane + 32 * ai, c = warp + POTF2F_UPD_WARPS * bi;
        a[ai][bi] = (i < n && c <= i) ? A[(i64)i * rs + (i64)c * cs] : 0.0f;
      }
  } else if (p < n) {
    dp = A[(i64)p * rs + (i64)p * cs];
  }
  __syncthreads();  // s_count initialised

  // serial pivot arithmetic of column jc (reference ldlt/factor.rs:122-160), done by ONE pivot thread
  auto publish_pivot = [&](int jc, float d) {
    int fail = 0;
    if (regularize) {
      if (d <= eps) {  // LLT: sign == +1
        d = delta;
        s_count += 1;  // single writer per column, ordered by the per-column barrier
      }
    }
    float inv = 0.0f;
    if (!(d > 0.0f)) {
      fail = 1;
    } else {
      const float sd = sqrtf(d);
      if (sd == 0.0f || !isfinite(sd)) fail = 1;
      else inv = 1.0f / sd;
    }
    s_inv[jc & 1] = inv;
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
    const float* col = colbuf[j & 1];
    if (s_fail[j & 1]) {
      if (tid == 0) info[0] = j0 + j;
      return;
    }
    const float inv = s_inv[j & 1];
    if (!is_upd) {
      // pivot group: keep the diagonal current with the SAME fma the update threads apply to a_pp, then start the
      // next column's pivot arithmetic immediately
      if (p > j && p < n) {
        const float l = col[p] * inv;
        dp = fmaf(-l, l, dp);
        if (p == j + 1) publish_pivot(j + 1, dp);
      }
    } else {
      const int jw = j & (POTF2F_UPD_WARPS - 1);
      // column j of L goes to global memory (owners: warp jw).
      // NB: like the reference, the stored diagonal is (unregularised a_jj) * recip(l_jj)
      // (ldlt/factor.rs:161-175 scales the whole column, diagonal included, and `diag` is a local copy).
      if (warp == jw) {
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) {
          const int i = lane + 32 * ai;
          if (i >= j && i < n) A[(i64)i * rs + (i64)j * cs] = col[i] * inv;
        }
      }
      // trailing update: a_ic <- fmaf(-l_cj, l_ij, a_ic) for j < c <= i. The column test depends only on
      // (warp, bi, j): warp-uniform, so dead column slots are BRANCHED over (no predicated-off instruction issue).
      if (warp + POTF2F_UPD_WARPS * (POTF2F_CB - 1) > j) {
        float li[4];
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) li[ai] = col[lane + 32 * ai] * inv;
#pragma unroll
        for (int bi = 0; bi < POTF2F_CB; ++bi) {
          const int c = warp + POTF2F_UPD_WARPS * bi;
          if (c > j && c < n) {
            const float lc = col[c] * inv;
#pragma unroll
            for (int ai = 0; ai < 4; ++ai) {
              const int i = lane + 32 * ai;
              if (32 * ai + 31 >= c) {  // warp-uniform: this row slot intersects i >= c
                if (i >= c && i < n) a[ai][bi] = fmaf(-lc, li[ai], a[ai][bi]);
              }
            }
          }
        }
      }
      // owners of column j+1 publish it (unscaled) into the other buffer
      if (j + 1 < n && warp == ((j + 1) & (POTF2F_UPD_WARPS - 1))) {
        const int nbk = (j + 1) / POTF2F_UPD_WARPS;
        float* nxt = colbuf[(j + 1) & 1];
#pragma unroll
        for (int ai = 0; ai < 4; ++ai) {
          float v = 0.0f;
#pragma unroll
          for (int bi = 0; bi < POTF2F_CB; ++bi)
            if (bi == nbk) v = a[ai][bi];
          nxt[lane + 32 * ai] = v;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0 && s_count) info[1] += s_count;
}

struct LltCtxF {
  cudaStream_t stream;
  int regularize;
  float eps, delta;
  long long* d_info;
  i64 nb;
};

void llt_rec_f32(const LltCtxF& ctx, VF A, i64 j0) {
  const i64 n = A.nrows;
  if (n <= ctx.nb) {
    potf2_f32_kernel<<<1, POTF2F_THREADS, 0, ctx.stream>>>(A.ptr, A.rs, A.cs, (int)n, j0, ctx.regularize, ctx.eps, ctx.delta,
                                                          ctx.d_info);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  i64 n1 = ((n / 2 + ctx.nb - 1) / ctx.nb) * ctx.nb;
  if (n1 >= n) n1 = ((n - 1) / ctx.nb) * ctx.nb;
  const i64 n2 = n - n1;
  VF A11 = A.sub(0, 0, n1, n1), A21 = A.sub(n1, 0, n2, n1), A22 = A.sub(n1, n1, n2, n2);
  llt_rec_f32(ctx, A11, j0);
  VCF cA11{A11.ptr, A11.nrows, A11.ncols, A11.rs, A11.cs}, cA21{A21.ptr, A21.nrows, A21.ncols, A21.rs, A21.cs};
  solve_lower_triangular_in_place_f32(ctx.stream, cA11, false, A21.t());
  gemm_f32(ctx.stream, A22, TRI_LOWER, 1, cA21, RECT, cA21.t(), RECT, -1.0f);
  llt_rec_f32(ctx, A22, j0 + n1);
}

}  // namespace

LltResult llt_cholesky_in_place_f32(cudaStream_t stream, VF A, float reg_delta, float reg_eps, LltParams params) {
  FB_ASSERT(A.nrows == A.ncols, "LLT needs a square matrix");
  const i64 n = A.nrows;
  LltResult res{true, 0, 0};
  if (n == 0) return res;
  const int regularize = (reg_delta > 0.0f && reg_eps > 0.0f) ? 1 : 0;
  i64 nb = (i64)params.block_size;
  if (nb <= 0 || nb > POTF2F_MAX) nb = POTF2F_MAX;
  long long* d_info = (long long*)ws_alloc(2 * sizeof(long long));
  long long h_info[2] = {-1, 0};
  FB_CUDA_CHECK(cudaMemcpyAsync(d_info, h_info, sizeof(h_info), cudaMemcpyHostToDevice, stream));
  LltCtxF ctx{stream, regularize, reg_eps, reg_delta, d_info, nb};
  llt_rec_f32(ctx, A, 0);
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

void llt_solve_in_place_f32(cudaStream_t stream, VCF L, VF rhs) {
  FB_ASSERT(L.nrows == L.ncols && rhs.nrows == L.nrows, "LLT solve shape mismatch");
  solve_lower_triangular_in_place_f32(stream, L, false, rhs);
  solve_upper_triangular_in_place_f32(stream, L.t(), false, rhs);
}

}  // namespace fb
