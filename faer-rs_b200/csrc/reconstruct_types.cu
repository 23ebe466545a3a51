// `*_reconstruct` / `*_inverse` on the factors for f32, c64 and c32 ("next" row; the f64 versions are in
// reconstruct.cu and stay as they are). Same compositions as the f64 file, written once over the scalar kind <R, CX>
// (R = real type, CX = interleaved complex): the structured product, the triangular solves, the LU solve, the row permutation
// and the block-Householder sequence of that scalar kind — every one of them already on the hot path and validated on its own.
// Complex views are in COMPLEX element units on an R* base (the convention of gemm_c64 / cplx.cu).
// Reference:
//   cholesky/llt/reconstruct.rs:12-33   out(lower) = L(lower) * L^H(upper)            (only the lower triangle is written)
//   cholesky/llt/inverse.rs:10-39       L_inv = L^-1 (lower), out(lower) = L_inv^H(upper) * L_inv(lower)
//   lu/partial_pivoting/reconstruct.rs:12-80   tmp = L U by structured products (square / tall / wide parts), out = P^-1 tmp
//   lu/partial_pivoting/inverse.rs:11-49       A^-1 from the factors; here as the LU solve applied to the identity
//   qr/no_pivoting/reconstruct.rs:13-39 out = [R; 0], then out <- Q out (sequence on the left, Conj::No)
//   qr/no_pivoting/inverse.rs:11-43     A^-1 = R^-1 Q^H; here as the QR solve applied to the identity (Q^H first, then R^-1)
// Tests: tests/test_gpu_zzzzzzzzz_2_reconstruct_types.py.
#include <algorithm>
#include <type_traits>

#include "flat_map.cuh"
#include "runtime.cuh"
#include "tensor_ops.cuh"

namespace fb {

namespace {

// Flat-map bodies (flat_map.cuh). W = 1 (real) or 2 (interleaved complex) R values per element; strides in elements.
#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif
template <class R, int W>
struct RtSetIdentity {
  R* A; i64 rs, cs, n;
  RT_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n) return;
    R* p = A + W * (i * rs + j * cs);
    p[0] = i == j ? R(1) : R(0);
    if (W == 2) p[W - 1] = R(0);
  }
};
// out (m x n) <- upper trapezoid of Rm (size x n) in its first `size` rows, zero elsewhere
template <class R, int W>
struct RtSetUpperTrapezoid {
  R* out; i64 o_rs, o_cs, m, n; const R* Rm; i64 r_rs, r_cs, size;
  RT_HD void operator()(i64 i, i64 j) const {
    if (i >= m || j >= n) return;
    R* p = out + W * (i * o_rs + j * o_cs);
    const bool take = i < size && i <= j;
    const R* q = Rm + W * (i * r_rs + j * r_cs);
    p[0] = take ? q[0] : R(0);
    if (W == 2) p[W - 1] = take ? q[W - 1] : R(0);
  }
};
// tmp (compact column-major, ld = nrows) [i, c] = src[perm[i], c]
template <class R, int W>
struct RtGatherRows {
  R* tmp; const R* src; i64 rs, cs, nrows, ncols; const long long* perm;
  RT_HD void operator()(i64 i, i64 c) const {
    if (i >= nrows || c >= ncols) return;
    const R* q = src + W * (perm[i] * rs + c * cs);
    R* p = tmp + W * (c * nrows + i);
    p[0] = q[0];
    if (W == 2) p[W - 1] = q[W - 1];
  }
};
template <class R, int W>
struct RtScatterBack {
  R* dst; i64 rs, cs; const R* tmp; i64 nrows, ncols;
  RT_HD void operator()(i64 i, i64 c) const {
    if (i >= nrows || c >= ncols) return;
    const R* q = tmp + W * (c * nrows + i);
    R* p = dst + W * (i * rs + c * cs);
    p[0] = q[0];
    if (W == 2) p[W - 1] = q[W - 1];
  }
};

// dst(i, j) = src(i, j) on the triangle `lower` selects — diagonal included unless `unit` — and nothing elsewhere
template <class R, int W>
struct RtCopyTriangle {
  R* dst; i64 d_rs, d_cs; const R* src; i64 s_rs, s_cs, n; int lower, unit;
  RT_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n) return;
    const bool in = lower ? (unit ? i > j : i >= j) : (unit ? i < j : i <= j);
    if (!in) return;
    const R* q = src + W * (i * s_rs + j * s_cs);
    R* p = dst + W * (i * d_rs + j * d_cs);
    p[0] = q[0];
    if (W == 2) p[W - 1] = q[W - 1];
  }
};

template <class R, bool CX>
struct Kind {
  static constexpr int W = CX ? 2 : 1;
  typedef View<R> V;
  typedef View<const R> VC;

  static V sub(V v, i64 i, i64 j, i64 m, i64 n) { return V{v.ptr + W * (i * v.rs + j * v.cs), m, n, v.rs, v.cs}; }
  static VC sub(VC v, i64 i, i64 j, i64 m, i64 n) { return VC{v.ptr + W * (i * v.rs + j * v.cs), m, n, v.rs, v.cs}; }
  static VC c(V v) { return VC{v.ptr, v.nrows, v.ncols, v.rs, v.cs}; }

  static void set_identity(cudaStream_t st, V A) {
    const i64 n = A.nrows;
    DevRun run{st};
    run(RtSetIdentity<R, W>{A.ptr, A.rs, A.cs, n}, n, n);
  }
  // rhs[i, :] <- rhs[perm[i], :]; perm: HOST int64 of rhs.nrows entries
  static void permute_rows(cudaStream_t st, V rhs, const long long* perm) {
    const i64 n = rhs.nrows, k = rhs.ncols;
    if (n == 0 || k == 0) return;
    long long* d_perm = (long long*)ws_alloc((size_t)n * 8);
    R* tmp = (R*)ws_alloc((size_t)n * (size_t)k * W * sizeof(R));
    FB_CUDA_CHECK(cudaMemcpyAsync(d_perm, perm, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    DevRun run{st};
    run(RtGatherRows<R, W>{tmp, rhs.ptr, rhs.rs, rhs.cs, n, k, d_perm}, n, k);
    run(RtScatterBack<R, W>{rhs.ptr, rhs.rs, rhs.cs, tmp, n, k}, n, k);
    FB_CUDA_CHECK(cudaStreamSynchronize(st));  // perm (host, pageable) and the pool buffers are released below
    ws_free(tmp);
    ws_free(d_perm);
  }
};

// ---- the building blocks by scalar kind (alpha = 1, Replace) ----
inline void rt_gemm(cudaStream_t st, VF dst, int ds, VCF a, int as, bool ca, VCF b, int bs, bool cb, std::false_type) {
  (void)ca; (void)cb;  // real: conjugation is the identity
  gemm_f32(st, dst, ds, 0, a, as, b, bs, 1.0f);
}
inline void rt_gemm(cudaStream_t st, VF dst, int ds, VCF a, int as, bool ca, VCF b, int bs, bool cb, std::true_type) {
  gemm_c32(st, dst, ds, 0, a, as, ca, b, bs, cb, 1.0f, 0.0f);
}
inline void rt_gemm(cudaStream_t st, VD dst, int ds, VCD a, int as, bool ca, VCD b, int bs, bool cb, std::true_type) {
  gemm_c64(st, dst, ds, 0, a, as, ca, b, bs, cb, 1.0, 0.0);
}
inline void rt_solve_lower(cudaStream_t st, VCF t, bool unit, VF rhs, std::false_type) { solve_lower_triangular_in_place_f32(st, t, unit, rhs); }
inline void rt_solve_lower(cudaStream_t st, VCF t, bool unit, VF rhs, std::true_type) { solve_lower_triangular_in_place_c32(st, t, unit, false, rhs); }
inline void rt_solve_lower(cudaStream_t st, VCD t, bool unit, VD rhs, std::true_type) { solve_lower_triangular_in_place_c64(st, t, unit, false, rhs); }
inline void rt_solve_upper(cudaStream_t st, VCF t, bool unit, VF rhs, std::false_type) { solve_upper_triangular_in_place_f32(st, t, unit, rhs); }
inline void rt_solve_upper(cudaStream_t st, VCF t, bool unit, VF rhs, std::true_type) { solve_upper_triangular_in_place_c32(st, t, unit, false, rhs); }
inline void rt_solve_upper(cudaStream_t st, VCD t, bool unit, VD rhs, std::true_type) { solve_upper_triangular_in_place_c64(st, t, unit, false, rhs); }
// rhs <- Q rhs (adjoint = false) or Q^H rhs (adjoint = true); Q^H is the transposed sequence with Conj::Yes composed in
// (householder.rs:768-808 as qr/no_pivoting/solve.rs:60-67 calls it)
inline void rt_hh_seq(cudaStream_t st, VCF b, VCF f, VF rhs, bool adjoint, std::false_type) {
  if (adjoint) apply_block_householder_sequence_transpose_on_the_left<float>(st, b, f, rhs);
  else apply_block_householder_sequence_on_the_left<float>(st, b, f, rhs);
}
inline void rt_hh_seq(cudaStream_t st, VCF b, VCF f, VF rhs, bool adjoint, std::true_type) {
  apply_householder_sequence_left_c32(st, b, f, adjoint, rhs, adjoint);
}
inline void rt_hh_seq(cudaStream_t st, VCD b, VCD f, VD rhs, bool adjoint, std::true_type) {
  apply_householder_sequence_left_c64(st, b, f, adjoint, rhs, adjoint);
}

}  // namespace

template <class R, bool CX>
void llt_reconstruct_t(cudaStream_t st, View<R> out, View<const R> L) {
  typedef std::integral_constant<bool, CX> Tag;
  const i64 n = out.nrows;
  FB_ASSERT(out.ncols == n && L.nrows == n && L.ncols == n, "llt_reconstruct shape mismatch");
  if (n == 0) return;
  rt_gemm(st, out, TRI_LOWER, L, TRI_LOWER, false, L.t(), TRI_UPPER, true, Tag());  // L * adjoint(L)
}

template <class R, bool CX>
void llt_inverse_t(cudaStream_t st, View<R> out, View<const R> L) {
  typedef Kind<R, CX> K;
  typedef std::integral_constant<bool, CX> Tag;
  const i64 n = out.nrows;
  FB_ASSERT(out.ncols == n && L.nrows == n && L.ncols == n, "llt_inverse shape mismatch");
  if (n == 0) return;
  R* buf = (R*)ws_alloc((size_t)n * (size_t)n * K::W * sizeof(R));
  View<R> Li{buf, n, n, 1, n};
  K::set_identity(st, Li);
  rt_solve_lower(st, L, false, Li, Tag());  // L_inv: lower triangular, exact zeros above the diagonal
  rt_gemm(st, out, TRI_LOWER, K::c(Li).t(), TRI_UPPER, true, K::c(Li), TRI_LOWER, false, Tag());  // adjoint(L_inv) * L_inv
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  ws_free(buf);
}

// perm_bwd: HOST int64[m], the inverse row permutation
template <class R, bool CX>
void lu_reconstruct_t(cudaStream_t st, View<R> out, View<const R> L, View<const R> U, const long long* perm_bwd) {
  typedef Kind<R, CX> K;
  typedef std::integral_constant<bool, CX> Tag;
  const i64 m = L.nrows, n = U.ncols, size = std::min(m, n);
  FB_ASSERT(out.nrows == m && out.ncols == n && L.ncols >= size && U.nrows >= size, "lu_reconstruct shape mismatch");
  if (m == 0 || n == 0) return;
  rt_gemm(st, K::sub(out, 0, 0, size, size), RECT, K::sub(L, 0, 0, size, size), UNIT_LOWER, false, K::sub(U, 0, 0, size, size),
          TRI_UPPER, false, Tag());
  if (m > n)
    rt_gemm(st, K::sub(out, size, 0, m - size, size), RECT, K::sub(L, size, 0, m - size, size), RECT, false,
            K::sub(U, 0, 0, size, size), TRI_UPPER, false, Tag());
  if (m < n)
    rt_gemm(st, K::sub(out, 0, size, size, n - size), RECT, K::sub(L, 0, 0, size, size), UNIT_LOWER, false,
            K::sub(U, 0, size, size, n - size), RECT, false, Tag());
  // (P A)[i, :] = A[perm_fwd[i], :]  =>  A[j, :] = (L U)[perm_bwd[j], :]
  K::permute_rows(st, out, perm_bwd);
}

template <class R, bool CX>
void lu_inverse_t(cudaStream_t st, View<R> out, View<const R> L, View<const R> U, const long long* perm_fwd) {
  typedef Kind<R, CX> K;
  typedef std::integral_constant<bool, CX> Tag;
  const i64 n = out.nrows;
  FB_ASSERT(out.ncols == n && L.nrows == n && L.ncols == n && U.nrows == n && U.ncols == n, "lu_inverse shape mismatch");
  if (n == 0) return;
  // lu/partial_pivoting/solve.rs:21-54 on the identity: rows permuted, unit-lower solve, upper solve
  K::set_identity(st, out);
  K::permute_rows(st, out, perm_fwd);
  rt_solve_lower(st, L, true, out, Tag());
  rt_solve_upper(st, U, false, out, Tag());
}

template <class R, bool CX>
void qr_reconstruct_t(cudaStream_t st, View<R> out, View<const R> Qb, View<const R> Qc, View<const R> Rm) {
  typedef Kind<R, CX> K;
  typedef std::integral_constant<bool, CX> Tag;
  const i64 m = Qb.nrows, n = Rm.ncols, size = std::min(m, n);
  FB_ASSERT(out.nrows == m && out.ncols == n && Qb.ncols == size && Qc.nrows > 0 && Qc.ncols == size && Rm.nrows == size,
            "qr_reconstruct shape mismatch");
  if (m == 0 || n == 0) return;
  DevRun run{st};
  run(RtSetUpperTrapezoid<R, K::W>{out.ptr, out.rs, out.cs, m, n, Rm.ptr, Rm.rs, Rm.cs, size}, m, n);
  if (size > 0) rt_hh_seq(st, Qb, Qc, out, false, Tag());
}

template <class R, bool CX>
void qr_inverse_t(cudaStream_t st, View<R> out, View<const R> Qb, View<const R> Qc, View<const R> Rm) {
  typedef Kind<R, CX> K;
  typedef std::integral_constant<bool, CX> Tag;
  const i64 n = out.nrows;
  FB_ASSERT(out.ncols == n && Qb.nrows == n && Qb.ncols == n && Qc.nrows > 0 && Qc.ncols == n && Rm.nrows == n && Rm.ncols == n,
            "qr_inverse shape mismatch");
  if (n == 0) return;
  K::set_identity(st, out);
  rt_hh_seq(st, Qb, Qc, out, true, Tag());     // Q^H
  rt_solve_upper(st, Rm, false, out, Tag());  // R^-1 Q^H
}

// linalg::triangular_inverse::invert_[unit_]{lower,upper}_triangular (triangular_inverse.rs): dst(triangle) <- src(triangle)^-1;
// only the triangle is written (the diagonal too unless `unit`). Here as the triangular solve applied to the identity.
inline void rt_solve_lower(cudaStream_t st, VCD t, bool unit, VD rhs, std::false_type) { solve_lower_triangular_in_place_f64(st, t, unit, rhs); }
inline void rt_solve_upper(cudaStream_t st, VCD t, bool unit, VD rhs, std::false_type) { solve_upper_triangular_in_place_f64(st, t, unit, rhs); }
template <class R, bool CX>
void inverse_triangular_t(cudaStream_t st, View<R> dst, View<const R> src, bool lower, bool unit) {
  typedef Kind<R, CX> K;
  typedef std::integral_constant<bool, CX> Tag;
  const i64 n = dst.nrows;
  FB_ASSERT(dst.ncols == n && src.nrows == n && src.ncols == n, "inverse_triangular shape mismatch");
  if (n == 0) return;
  R* buf = (R*)ws_alloc((size_t)n * (size_t)n * K::W * sizeof(R));
  View<R> X{buf, n, n, 1, n};
  K::set_identity(st, X);
  if (lower) rt_solve_lower(st, src, unit, X, Tag());
  else rt_solve_upper(st, src, unit, X, Tag());
  DevRun run{st};
  run(RtCopyTriangle<R, K::W>{dst.ptr, dst.rs, dst.cs, buf, 1, n, n, lower ? 1 : 0, unit ? 1 : 0}, n, n);
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  ws_free(buf);
}
template void inverse_triangular_t<double, false>(cudaStream_t, View<double>, View<const double>, bool, bool);
template void inverse_triangular_t<float, false>(cudaStream_t, View<float>, View<const float>, bool, bool);
template void inverse_triangular_t<double, true>(cudaStream_t, View<double>, View<const double>, bool, bool);
template void inverse_triangular_t<float, true>(cudaStream_t, View<float>, View<const float>, bool, bool);

#define FB_RECON_INST(R, CX)                                                                                                     \
  template void llt_reconstruct_t<R, CX>(cudaStream_t, View<R>, View<const R>);                                                  \
  template void llt_inverse_t<R, CX>(cudaStream_t, View<R>, View<const R>);                                                      \
  template void lu_reconstruct_t<R, CX>(cudaStream_t, View<R>, View<const R>, View<const R>, const long long*);                  \
  template void lu_inverse_t<R, CX>(cudaStream_t, View<R>, View<const R>, View<const R>, const long long*);                      \
  template void qr_reconstruct_t<R, CX>(cudaStream_t, View<R>, View<const R>, View<const R>, View<const R>);                     \
  template void qr_inverse_t<R, CX>(cudaStream_t, View<R>, View<const R>, View<const R>, View<const R>);
FB_RECON_INST(float, false)
FB_RECON_INST(double, true)
FB_RECON_INST(float, true)
#undef FB_RECON_INST

}  // namespace fb
