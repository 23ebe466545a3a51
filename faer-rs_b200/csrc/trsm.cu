// G3: in-place triangular solve  rhs <- op(T)^-1 rhs  (f64 and f32).
//
// Reference semantics: faer/src/linalg/triangular_solve.rs:220-419 (public entry points),
// 420-576 (recursive split: solve top, rhs_bot -= T_bot_left * rhs_top, solve bottom),
// 200-211 (`block_size` split rule), 577-604 (upper = lower on reversed views),
// 16-198 (leaves: reciprocal of the diagonal, then multiply).
//
// B200 mapping: the recursion stays on the host (it is O(n/64) launches); every off-diagonal update is a tensor-core
// GEMM launch (DMMA for f64, 3xTF32 for f32); the <=64-wide diagonal leaves run as one thread per right-hand-side
// column: the thread pulls its whole column (<=64 values) into REGISTERS with independent loads (all in flight at once;
// coalesced across the warp when rhs columns are contiguous, full lines per thread when rhs rows are), substitutes
// against T broadcast from shared memory with two interleaved FMA chains, and writes the column back.
// (A fused 128-wide leaf was tried and measured slower — 71 us vs ~38 us for two 64-leaves + one small GEMM at
// n = 16384 — because its 8128 dependent FMAs per thread are latency-bound; see profiles/r01_llt16384_launches_v3.txt.)
#include "gemm_f32.cuh"
#include "linalg_f64.cuh"

namespace fb {

namespace {

constexpr int LEAF = 64;        // diagonal leaf size (x lives in registers: LEAF values per thread)
constexpr int LEAF_COLS = 128;  // rhs columns per CTA (= threads)

template <class T>
__global__ void __launch_bounds__(LEAF_COLS) trsm_leaf_lower_kernel(const T* __restrict__ Tm, i64 t_rs, i64 t_cs, int n,
                                                                     int unit, T* __restrict__ R, i64 r_rs, i64 r_cs,
                                                                     i64 ncols) {
  __shared__ T Ts[LEAF][LEAF + 1];
  __shared__ T Tinv[LEAF];
  const int tid = threadIdx.x;
  const i64 c = (i64)blockIdx.x * LEAF_COLS + tid;
  const bool active = c < ncols;

  // T tile: fixed trip count + predicates, unrolled => 16 independent loads per thread in flight
  {
    const int i = tid & (LEAF - 1);
#pragma unroll 16
    for (int jj = 0; jj < LEAF / 2; ++jj) {
      const int j = (tid >> 6) + 2 * jj;
      T v = T(0);
      if (i < n && j <= i) v = Tm[i * t_rs + j * t_cs];
      Ts[i][j] = v;
    }
  }
  // this thread's rhs column -> registers (independent loads)
  T x[LEAF];
  T* col = R + c * r_cs;
#pragma unroll
  for (int i = 0; i < LEAF; ++i) x[i] = (active && i < n) ? col[i * r_rs] : T(0);
  __syncthreads();
  if (tid < LEAF) Tinv[tid] = (unit || tid >= n) ? T(1) : T(1) / Ts[tid][tid];
  __syncthreads();

#pragma unroll
  for (int i = 0; i < LEAF; ++i) {
    // two interleaved partial sums halve the dependent-FMA chain; rows >= n see T == 0 and are never stored
    T s0 = x[i], s1 = T(0);
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

inline void gemm_update(cudaStream_t st, VD dst, VCD lhs, VCD rhs) { gemm_f64(st, dst, 1, lhs, rhs, -1.0); }
inline void gemm_update(cudaStream_t st, VF dst, VCF lhs, VCF rhs) { gemm_f32(st, dst, 1, lhs, rhs, -1.0f); }

template <class T>
void solve_lower_rec(cudaStream_t stream, View<const T> Tm, bool unit, View<T> rhs) {
  const i64 n = Tm.nrows;
  if (n == 0 || rhs.ncols == 0) return;
  if (n <= LEAF) {
    unsigned blocks = (unsigned)((rhs.ncols + LEAF_COLS - 1) / LEAF_COLS);
    trsm_leaf_lower_kernel<T><<<blocks, LEAF_COLS, 0, stream>>>(Tm.ptr, Tm.rs, Tm.cs, (int)n, unit ? 1 : 0, rhs.ptr, rhs.rs,
                                                                rhs.cs, rhs.ncols);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    return;
  }
  const i64 bs = split_size(n);
  View<const T> T00 = Tm.sub(0, 0, bs, bs), T10 = Tm.sub(bs, 0, n - bs, bs), T11 = Tm.sub(bs, bs, n - bs, n - bs);
  View<T> top = rhs.sub(0, 0, bs, rhs.ncols), bot = rhs.sub(bs, 0, n - bs, rhs.ncols);
  solve_lower_rec<T>(stream, T00, unit, top);
  gemm_update(stream, bot, T10, View<const T>{top.ptr, top.nrows, top.ncols, top.rs, top.cs});
  solve_lower_rec<T>(stream, T11, unit, bot);
}

}  // namespace

void solve_lower_triangular_in_place_f64(cudaStream_t stream, VCD tril, bool unit, VD rhs) {
  FB_ASSERT(tril.nrows == tril.ncols && rhs.nrows == tril.ncols, "triangular solve shape mismatch");
  solve_lower_rec<double>(stream, tril, unit, rhs);
}
void solve_upper_triangular_in_place_f64(cudaStream_t stream, VCD triu, bool unit, VD rhs) {
  FB_ASSERT(triu.nrows == triu.ncols && rhs.nrows == triu.ncols, "triangular solve shape mismatch");
  if (triu.nrows == 0 || rhs.ncols == 0) return;
  solve_lower_rec<double>(stream, triu.rev_rows_cols(), unit, rhs.rev_rows());
}
void solve_lower_triangular_in_place_f32(cudaStream_t stream, VCF tril, bool unit, VF rhs) {
  FB_ASSERT(tril.nrows == tril.ncols && rhs.nrows == tril.ncols, "triangular solve shape mismatch");
  solve_lower_rec<float>(stream, tril, unit, rhs);
}
void solve_upper_triangular_in_place_f32(cudaStream_t stream, VCF triu, bool unit, VF rhs) {
  FB_ASSERT(triu.nrows == triu.ncols && rhs.nrows == triu.ncols, "triangular solve shape mismatch");
  if (triu.nrows == 0 || rhs.ncols == 0) return;
  solve_lower_rec<float>(stream, triu.rev_rows_cols(), unit, rhs.rev_rows());
}

}  // namespace fb
