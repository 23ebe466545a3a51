// G3: in-place triangular solve  rhs <- op(T)^-1 rhs  (f64).
//
// Reference semantics: faer/src/linalg/triangular_solve.rs:220-419 (public entry points),
// 420-576 (recursive split: solve top, rhs_bot -= T_bot_left * rhs_top, solve bottom),
// 200-211 (`block_size` split rule), 577-604 (upper = lower on reversed views),
// 16-198 (leaves: reciprocal of the diagonal, then multiply).
//
// B200 mapping: the recursion stays on the host (it is O(n/64) launches); every off-diagonal update is a
// DMMA GEMM launch (gemm_f64); the <=64-wide diagonal leaves run as one thread per right-hand-side column:
// the thread pulls its whole column (<=64 values) into REGISTERS with independent loads (all in flight at once;
// coalesced across the warp when rhs columns are contiguous, full 128-B lines per thread when rhs rows are),
// substitutes against T broadcast from shared memory with two interleaved FMA chains, and writes the column back.
#include "linalg_f64.cuh"

namespace fb {

namespace {

constexpr int LEAF = 64;        // diagonal leaf size (x lives in registers: LEAF doubles per thread)
constexpr int LEAF_COLS = 128;  // rhs columns per CTA (= threads)

__global__ void __launch_bounds__(LEAF_COLS) trsm_leaf_lower_kernel(const double* __restrict__ T, i64 t_rs, i64 t_cs,
                                                                     int n, int unit, double* __restrict__ R, i64 r_rs,
                                                                     i64 r_cs, i64 ncols) {
  __shared__ double Ts[LEAF][LEAF + 1];
  __shared__ double Tinv[LEAF];
  const int tid = threadIdx.x;
  const i64 c = (i64)blockIdx.x * LEAF_COLS + tid;
  const bool active = c < ncols;

  // T tile: fixed trip count + predicates, unrolled => 16 independent loads per thread in flight
  {
    const int i = tid & (LEAF - 1);
#pragma unroll 16
    for (int jj = 0; jj < LEAF / 2; ++jj) {
      const int j = (tid >> 6) + 2 * jj;
      double v = 0.0;
      if (i < n && j <= i) v = T[i * t_rs + j * t_cs];
      Ts[i][j] = v;
    }
  }
  // this thread's rhs column -> registers (independent loads)
  double x[LEAF];
  double* col = R + c * r_cs;
#pragma unroll
  for (int i = 0; i < LEAF; ++i) x[i] = (active && i < n) ? col[i * r_rs] : 0.0;
  __syncthreads();
  if (tid < LEAF) Tinv[tid] = (unit || tid >= n) ? 1.0 : 1.0 / Ts[tid][tid];
  __syncthreads();

#pragma unroll
  for (int i = 0; i < LEAF; ++i) {
    // two interleaved partial sums halve the dependent-FMA chain; rows >= n see T == 0 and are never stored
    double s0 = x[i], s1 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < i; k += 2) {
      s0 = fma(-Ts[i][k], x[k], s0);
      s1 = fma(-Ts[i][k + 1], x[k + 1], s1);
    }
    if (i & 1) s0 = fma(-Ts[i][i - 1], x[i - 1], s0);
    x[i] = (s0 + s1) * Tinv[i];
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < LEAF; ++i)
      if (i < n) col[i * r_rs] = x[i];
  }
}

// Fused leaf for 64 < n <= 128: one launch does  x_top = T11^-1 b_top;  b_bot -= T21 x_top;  x_bot = T22^-1 b_bot
// (the reference's recursion step, triangular_solve.rs:541-576, for one split) with T11/T21/T22 in shared memory.
// x_top stays in registers while the bottom rows are streamed through in chunks (updated in place in global memory),
// then the bottom column is pulled into registers for the second substitution.
constexpr int LEAF2 = 2 * LEAF;
constexpr size_t LEAF2_SMEM = sizeof(double) * (3 * LEAF * (LEAF + 1) + LEAF2);

__global__ void __launch_bounds__(LEAF_COLS) trsm_leaf128_lower_kernel(const double* __restrict__ T, i64 t_rs, i64 t_cs,
                                                                        int n, int unit, double* __restrict__ R, i64 r_rs,
                                                                        i64 r_cs, i64 ncols) {
  extern __shared__ double sm2[];
  double(*T11)[LEAF + 1] = reinterpret_cast<double(*)[LEAF + 1]>(sm2);
  double(*T21)[LEAF + 1] = reinterpret_cast<double(*)[LEAF + 1]>(sm2 + LEAF * (LEAF + 1));
  double(*T22)[LEAF + 1] = reinterpret_cast<double(*)[LEAF + 1]>(sm2 + 2 * LEAF * (LEAF + 1));
  double* Tinv = sm2 + 3 * LEAF * (LEAF + 1);
  const int tid = threadIdx.x;
  const i64 c = (i64)blockIdx.x * LEAF_COLS + tid;
  const bool active = c < ncols;
  const int nbot = n - LEAF;  // 1..64

  {
    const int i = tid & (LEAF - 1);
#pragma unroll 16
    for (int jj = 0; jj < LEAF / 2; ++jj) {
      const int j = (tid >> 6) + 2 * jj;
      T11[i][j] = (j <= i) ? T[i * t_rs + j * t_cs] : 0.0;
    }
#pragma unroll 16
    for (int jj = 0; jj < LEAF / 2; ++jj) {
      const int j = (tid >> 6) + 2 * jj;
      T21[i][j] = (i < nbot) ? T[(LEAF + i) * t_rs + j * t_cs] : 0.0;
    }
#pragma unroll 16
    for (int jj = 0; jj < LEAF / 2; ++jj) {
      const int j = (tid >> 6) + 2 * jj;
      T22[i][j] = (i < nbot && j <= i) ? T[(LEAF + i) * t_rs + (LEAF + j) * t_cs] : 0.0;
    }
  }
  double x[LEAF];
  double* col = R + c * r_cs;
#pragma unroll
  for (int i = 0; i < LEAF; ++i) x[i] = active ? col[i * r_rs] : 0.0;
  __syncthreads();
  if (tid < LEAF) {
    Tinv[tid] = unit ? 1.0 : 1.0 / T11[tid][tid];
    Tinv[LEAF + tid] = (unit || tid >= nbot) ? 1.0 : 1.0 / T22[tid][tid];
  }
  __syncthreads();

  // ---- top solve ----
#pragma unroll
  for (int i = 0; i < LEAF; ++i) {
    double s0 = x[i], s1 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < i; k += 2) {
      s0 = fma(-T11[i][k], x[k], s0);
      s1 = fma(-T11[i][k + 1], x[k + 1], s1);
    }
    if (i & 1) s0 = fma(-T11[i][i - 1], x[i - 1], s0);
    x[i] = (s0 + s1) * Tinv[i];
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < LEAF; ++i) col[i * r_rs] = x[i];
  }
  // ---- bottom rows: b_bot -= T21 x_top, streamed 8 rows at a time ----
  double* colb = col + (i64)LEAF * r_rs;
  for (int i0 = 0; i0 < LEAF; i0 += 8) {
    double y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) y[u] = (active && i0 + u < nbot) ? colb[(i0 + u) * r_rs] : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double s0 = y[u], s1 = 0.0;
#pragma unroll
      for (int k = 0; k < LEAF; k += 2) {
        s0 = fma(-T21[i0 + u][k], x[k], s0);
        s1 = fma(-T21[i0 + u][k + 1], x[k + 1], s1);
      }
      y[u] = s0 + s1;
    }
    if (active) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u < nbot) colb[(i0 + u) * r_rs] = y[u];
    }
  }
  // ---- bottom solve (the thread re-reads its own just-written values) ----
#pragma unroll
  for (int i = 0; i < LEAF; ++i) x[i] = (active && i < nbot) ? colb[i * r_rs] : 0.0;
#pragma unroll
  for (int i = 0; i < LEAF; ++i) {
    double s0 = x[i], s1 = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < i; k += 2) {
      s0 = fma(-T22[i][k], x[k], s0);
      s1 = fma(-T22[i][k + 1], x[k + 1], s1);
    }
    if (i & 1) s0 = fma(-T22[i][i - 1], x[i - 1], s0);
    x[i] = (s0 + s1) * Tinv[LEAF + i];
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < LEAF; ++i)
      if (i < nbot) colb[i * r_rs] = x[i];
  }
}

// faer's split rule (reference: triangular_solve.rs:200-211)
inline i64 split_size(i64 n) {
  i64 base_rem = n / 2;
  i64 sub;
  if (n >= 32) sub = (base_rem + 15) / 16 * 16;
  else if (n >= 16) sub = (base_rem + 7) / 8 * 8;
  else if (n >= 8) sub = (base_rem + 3) / 4 * 4;
  else sub = base_rem;
  return n - sub;
}

void solve_lower_rec(cudaStream_t stream, VCD T, bool unit, VD rhs) {
  const i64 n = T.nrows;
  if (n == 0 || rhs.ncols == 0) return;
  if (n <= LEAF) {
    unsigned blocks = (unsigned)((rhs.ncols + LEAF_COLS - 1) / LEAF_COLS);
    trsm_leaf_lower_kernel<<<blocks, LEAF_COLS, 0, stream>>>(T.ptr, T.rs, T.cs, (int)n, unit ? 1 : 0, rhs.ptr, rhs.rs,
                                                              rhs.cs, rhs.ncols);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  // The fused 65..128 leaf is kept for reference but disabled: its 8128 dependent FMAs per thread measured 71 us per
  // launch (profiles/r01_llt16384_launches_v3.txt) against ~38 us for two 64-leaves + one small DMMA GEMM.
  if (false && n <= LEAF2) {
    static bool configured = false;
    if (!configured) {
      FB_CUDA_CHECK(cudaFuncSetAttribute(trsm_leaf128_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)LEAF2_SMEM));
      configured = true;
    }
    unsigned blocks = (unsigned)((rhs.ncols + LEAF_COLS - 1) / LEAF_COLS);
    trsm_leaf128_lower_kernel<<<blocks, LEAF_COLS, LEAF2_SMEM, stream>>>(T.ptr, T.rs, T.cs, (int)n, unit ? 1 : 0, rhs.ptr,
                                                                         rhs.rs, rhs.cs, rhs.ncols);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  const i64 bs = split_size(n);
  VCD T00 = T.sub(0, 0, bs, bs), T10 = T.sub(bs, 0, n - bs, bs), T11 = T.sub(bs, bs, n - bs, n - bs);
  VD top = rhs.sub(0, 0, bs, rhs.ncols), bot = rhs.sub(bs, 0, n - bs, rhs.ncols);
  solve_lower_rec(stream, T00, unit, top);
  gemm_f64(stream, bot, 1, T10, cv(top), -1.0);
  solve_lower_rec(stream, T11, unit, bot);
}

}  // namespace

void solve_lower_triangular_in_place_f64(cudaStream_t stream, VCD tril, bool unit, VD rhs) {
  FB_ASSERT(tril.nrows == tril.ncols && rhs.nrows == tril.ncols, "triangular solve shape mismatch");
  solve_lower_rec(stream, tril, unit, rhs);
}

void solve_upper_triangular_in_place_f64(cudaStream_t stream, VCD triu, bool unit, VD rhs) {
  FB_ASSERT(triu.nrows == triu.ncols && rhs.nrows == triu.ncols, "triangular solve shape mismatch");
  if (triu.nrows == 0 || rhs.ncols == 0) return;
  solve_lower_rec(stream, triu.rev_rows_cols(), unit, rhs.rev_rows());
}

}  // namespace fb
