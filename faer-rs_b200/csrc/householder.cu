// Block Householder machinery (f32 / f64): T-factor build and block-reflector application, as GEMM compositions.
//
// Reference: faer/src/linalg/householder.rs
//   conventions 10-23: H = I - v v^H / tau, v_0 = 1; H_0 ... H_{b-1} = I - V T^-1 V^H, T = striu(V^H V) + diag(tau)
//   upgrade_householder_factor 132-272: T(strict upper) = V_top^H(unit upper) V_top(unit lower) + V_bot^H V_bot
//   apply_block_householder_on_the_left_in_place_generic 370-620:
//       W = V_top^H(unit upper) M_top + V_bot^H M_bot;  W <- T^-H W (forward) or T^-1 W;  M_top -= V_top(unit lower) W;
//       M_bot -= V_bot W
// V shares storage with R (the part on/above the diagonal is NOT data of V): the structured-GEMM masks treat it as
// the implicit unit diagonal / zeros, exactly like the reference's triangular matmul.
// Tall-skinny products (V^H M with k = m) go through the GEMM's split-K path automatically.
#include "runtime.cuh"
#include "tensor_ops.cuh"

namespace fb {

template <class T>
void householder_build_t(cudaStream_t st, View<const T> V, View<T> Tf) {
  const i64 N = V.ncols, m = V.nrows;
  FB_ASSERT(Tf.nrows == N && Tf.ncols == N && m >= N, "householder T-factor shape mismatch");
  if (N <= 1) return;
  View<const T> Vt = V.sub(0, 0, N, N), Vb = V.sub(N, 0, m - N, N);
  gemm(st, Tf, UNIT_UPPER, 0, Vt.t(), UNIT_UPPER, Vt, UNIT_LOWER, T(1));
  if (m > N) gemm(st, Tf, UNIT_UPPER, 1, Vb.t(), RECT, Vb, RECT, T(1));
}

template <class T>
void apply_block_householder_on_the_left(cudaStream_t st, View<const T> V, View<const T> Tf, View<T> M, bool forward,
                                         T* tmp_buf) {
  const i64 N = V.ncols, m = V.nrows, K = M.ncols;
  FB_ASSERT(Tf.nrows == N && Tf.ncols == N && M.nrows == m && m >= N, "block Householder shape mismatch");
  if (N == 0 || K == 0) return;
  // `tmp_buf` (>= N * K elements, owned by the caller) avoids the pool allocation and the host synchronisation that
  // returning it to the stream-agnostic pool needs
  T* buf = tmp_buf ? tmp_buf : (T*)ws_alloc((size_t)N * K * sizeof(T));
  View<T> tmp{buf, N, K, 1, N};
  View<const T> Vt = V.sub(0, 0, N, N), Vb = V.sub(N, 0, m - N, N);
  View<T> top = M.sub(0, 0, N, K), bot = M.sub(N, 0, m - N, K);
  gemm(st, tmp, RECT, 0, Vt.t(), UNIT_UPPER, cview(top), RECT, T(1));
  if (m > N) gemm(st, tmp, RECT, 1, Vb.t(), RECT, cview(bot), RECT, T(1));
  if (forward) solve_lower(st, Tf.t(), false, tmp);
  else solve_upper(st, Tf, false, tmp);
  gemm(st, top, RECT, 1, Vt, UNIT_LOWER, cview(tmp), RECT, T(-1));
  if (m > N) gemm(st, bot, RECT, 1, Vb, RECT, cview(tmp), RECT, T(-1));
  if (!tmp_buf) {
    FB_CUDA_CHECK(cudaStreamSynchronize(st));  // the pool is stream-agnostic
    ws_free(buf);
  }
}

// householder.rs:724-765: M <- H_0 H_1 ... H_{k-1} M  (blocks applied last to first)
template <class T>
void apply_block_householder_sequence_on_the_left(cudaStream_t st, View<const T> basis, View<const T> factor, View<T> M) {
  const i64 m = basis.nrows, n = basis.ncols, bs = factor.nrows, size = factor.ncols;
  FB_ASSERT(bs > 0 && size == std::min(m, n) && M.nrows == m, "Householder sequence shape mismatch");
  i64 j = size;
  i64 b = size % bs;
  if (b == 0) b = bs;
  while (j > 0) {
    const i64 jp = j - b;
    apply_block_householder_on_the_left<T>(st, basis.sub(jp, jp, m - jp, j - jp), factor.sub(0, jp, j - jp, j - jp),
                                           M.sub(jp, 0, m - jp, M.ncols), false);
    j = jp;
    b = bs;
  }
}
// householder.rs:768-808: M <- (H_0 ... H_{k-1})^H M  (blocks applied first to last)
template <class T>
void apply_block_householder_sequence_transpose_on_the_left(cudaStream_t st, View<const T> basis, View<const T> factor,
                                                            View<T> M) {
  const i64 m = basis.nrows, n = basis.ncols, bs = factor.nrows, size = factor.ncols;
  FB_ASSERT(bs > 0 && size == std::min(m, n) && M.nrows == m, "Householder sequence shape mismatch");
  for (i64 j = 0; j < size;) {
    const i64 b = std::min(bs, size - j);
    apply_block_householder_on_the_left<T>(st, basis.sub(j, j, m - j, b), factor.sub(0, j, b, b),
                                           M.sub(j, 0, m - j, M.ncols), true);
    j += b;
  }
}
template void apply_block_householder_sequence_on_the_left<double>(cudaStream_t, View<const double>, View<const double>, View<double>);
template void apply_block_householder_sequence_on_the_left<float>(cudaStream_t, View<const float>, View<const float>, View<float>);
template void apply_block_householder_sequence_transpose_on_the_left<double>(cudaStream_t, View<const double>, View<const double>, View<double>);
template void apply_block_householder_sequence_transpose_on_the_left<float>(cudaStream_t, View<const float>, View<const float>, View<float>);

template void householder_build_t<double>(cudaStream_t, View<const double>, View<double>);
template void householder_build_t<float>(cudaStream_t, View<const float>, View<float>);
template void apply_block_householder_on_the_left<double>(cudaStream_t, View<const double>, View<const double>, View<double>, bool, double*);
template void apply_block_householder_on_the_left<float>(cudaStream_t, View<const float>, View<const float>, View<float>, bool, float*);

}  // namespace fb
