// Type-generic spellings (f32 / f64) of the tensor-core building blocks, so that the Householder / QR drivers are written
// once. f64 -> DMMA GEMM (gemm_f64.cu), f32 -> 3xTF32 GEMM (gemm_f32.cu); triangular solves from trsm.cu.
#pragma once
#include "gemm_f32.cuh"
#include "linalg_f64.cuh"

namespace fb {

inline void gemm(cudaStream_t st, VD dst, int ds, int accum, VCD lhs, int ls, VCD rhs, int rs, double alpha) {
  gemm_f64(st, dst, ds, accum, lhs, ls, rhs, rs, alpha);
}
inline void gemm(cudaStream_t st, VF dst, int ds, int accum, VCF lhs, int ls, VCF rhs, int rs, float alpha) {
  gemm_f32(st, dst, ds, accum, lhs, ls, rhs, rs, alpha);
}
inline void solve_lower(cudaStream_t st, VCD t, bool unit, VD rhs) { solve_lower_triangular_in_place_f64(st, t, unit, rhs); }
inline void solve_lower(cudaStream_t st, VCF t, bool unit, VF rhs) { solve_lower_triangular_in_place_f32(st, t, unit, rhs); }
inline void solve_upper(cudaStream_t st, VCD t, bool unit, VD rhs) { solve_upper_triangular_in_place_f64(st, t, unit, rhs); }
inline void solve_upper(cudaStream_t st, VCF t, bool unit, VF rhs) { solve_upper_triangular_in_place_f32(st, t, unit, rhs); }

template <class T>
inline View<const T> cview(const View<T>& v) { return View<const T>{v.ptr, v.nrows, v.ncols, v.rs, v.cs}; }

// ---- householder.cu ----
// T(strict upper) <- V^H V for the N reflectors stored in V (m x N, implicit unit diagonal, R above it is ignored);
// the diagonal of Tf (the taus) is left untouched.  Reference: householder.rs:132-272 (upgrade_householder_factor)
template <class T>
void householder_build_t(cudaStream_t st, View<const T> V, View<T> Tf);
// M <- (I - V T^-1 V^H) M  (forward = false)   or   M <- (I - V T^-H V^H) M  (forward = true)
// Reference: householder.rs:370-620 (apply_block_householder_on_the_left_in_place_generic)
template <class T>
void apply_block_householder_on_the_left(cudaStream_t st, View<const T> V, View<const T> Tf, View<T> M, bool forward,
                                         T* tmp_buf = nullptr);

// sequences of block reflectors stored as (basis = V below the diagonal, factor = block_size x size T blocks)
// Reference: householder.rs:724-808
template <class T>
void apply_block_householder_sequence_on_the_left(cudaStream_t st, View<const T> basis, View<const T> factor, View<T> M);
template <class T>
void apply_block_householder_sequence_transpose_on_the_left(cudaStream_t st, View<const T> basis, View<const T> factor,
                                                            View<T> M);

// ---- qr.cu ----
// Householder QR without pivoting, reference qr/no_pivoting/factor.rs:258-301. H: block_size x min(m, n).
// Returns the rank (dependent columns are skipped and the reflectors compacted as factor.rs:40-83 does).
template <class T>
i64 qr_in_place(cudaStream_t st, View<T> A, View<T> H);
i64 qr_recommended_block_size(i64 nrows, i64 ncols);

// ---- bidiag.cu ----
// A = U B V^H, m >= n, column-major A. Reference: svd/bidiag.rs:47-256. Hl: bl x n, Hr: br x (n-1) (T blocks).
template <class T>
void bidiag_in_place(cudaStream_t st, View<T> A, View<T> Hl, View<T> Hr);
// ---- reconstruct.cu ----
// out = Q [R; 0]  (qr/no_pivoting/reconstruct.rs:13-39)
template <class T>
void qr_reconstruct(cudaStream_t st, View<T> out, View<const T> Q_basis, View<const T> Q_coeff, View<const T> R);
// ---- evd.cu ----
// S (device, compact, n entries) <- eigenvalues (nondecreasing) of the self-adjoint matrix whose lower triangle is in A
// (evd/mod.rs:270-353 with u = None)
// returns false if the tridiagonal form is not finite (EvdError::NoConvergence in the reference)
template <class T>
bool self_adjoint_eigenvalues(cudaStream_t st, View<const T> A, T* S);
// ---- svd.cu ----
// S (device, compact, min(m, n) entries) <- the singular values of A, non-increasing (svd/mod.rs:530-648 with u = v = None)
// qr_ratio_threshold: SvdParams::qr_ratio_threshold (11/6): taller inputs go through QR first
// returns false if the bidiagonal form is not finite (SvdError::NoConvergence, svd/mod.rs:282-286)
template <class T>
bool singular_values(cudaStream_t st, View<const T> A, T* S, double qr_ratio_threshold);
// ---- svd_vectors.cu: decompositions WITH vectors (f64 arithmetic), false on non-finite input ----
template <class TA>
bool svd_with_vectors(cudaStream_t st, View<const TA> A, View<TA> U, TA* S, i64 sstride, View<TA> V, double qr_ratio_threshold);
template <class TA>
bool self_adjoint_evd_with_vectors(cudaStream_t st, View<const TA> A, View<TA> U, TA* S, i64 sstride);
template <class T>
bool device_all_finite(cudaStream_t st, const T* x, i64 n);
// SVD of the real upper-bidiagonal (d, e) of order n (device arrays; e has n entries, the last one unused): S_sorted
// non-increasing, UB / VB n x n column-major (ld n) with B = UB diag(S) VB^T. False on non-finite input.
bool bidiag_svd_vectors_f64(cudaStream_t st, const double* d, const double* e, i64 n, double* S_sorted, double* UB, double* VB);
// ---- cplx_condensed.cu: `svd` / `self_adjoint_evd` for complex T (TO = double: c64, float: c32 computed in c64); views in
// COMPLEX element units on a TO* base; U / V with ptr == nullptr or ncols == 0: not computed; S: (value, 0) pairs `sstride`
// complex elements apart (device). False on non-finite input. ----
template <class TO>
bool svd_cx(cudaStream_t st, View<const TO> A, View<TO> U, TO* S, i64 sstride, View<TO> V);
template <class TO>
bool self_adjoint_evd_cx(cudaStream_t st, View<const TO> A, View<TO> U, TO* S, i64 sstride);
// complex bidiagonalization / tridiagonalization with T blocks (TO = double: c64, float: c32 computed in c64; complex-unit views)
template <class TO>
void bidiag_in_place_cx(cudaStream_t st, View<TO> A, View<TO> Hl, View<TO> Hr);
template <class TO>
void tridiag_in_place_cx(cudaStream_t st, View<TO> A, View<TO> H);
// A <- upper Hessenberg form + reflectors, Hf (bs x (n - 1)) <- T blocks (evd/hessenberg.rs:549-567); <TO, complex?>: <double, false> f64,
// <float, false> f32, <double, true> c64, <float, true> c32 (complex views in complex units)
template <class TO, bool CX>
void hessenberg_in_place_t(cudaStream_t st, View<TO> A, View<TO> Hf);
// ---- tridiag_dc.cu: divide-and-conquer eigensolver of a symmetric tridiagonal matrix (device arrays) ----
bool tridiag_dc_f64(cudaStream_t st, const double* d, const double* e, i64 n, double* lam, double* Q, i64 ldq);
// ---- tridiag.cu ----
// A = Q T Q^H, self-adjoint A (lower triangle), column-major. Reference: evd/tridiag.rs:274-529. H: b x (n-1).
template <class T>
void tridiag_in_place(cudaStream_t st, View<T> A, View<T> H);

// ---- reconstruct_types.cu: `*_reconstruct` / `*_inverse` on the factors for the scalar kinds <float, false> (f32),
// <double, true> (c64), <float, true> (c32); complex views in COMPLEX element units on an R* base; perm arrays HOST int64 ----
template <class R, bool CX>
void llt_reconstruct_t(cudaStream_t st, View<R> out, View<const R> L);
template <class R, bool CX>
void llt_inverse_t(cudaStream_t st, View<R> out, View<const R> L);
template <class R, bool CX>
void lu_reconstruct_t(cudaStream_t st, View<R> out, View<const R> L, View<const R> U, const long long* perm_bwd_host);
template <class R, bool CX>
void lu_inverse_t(cudaStream_t st, View<R> out, View<const R> L, View<const R> U, const long long* perm_fwd_host);
template <class R, bool CX>
void qr_reconstruct_t(cudaStream_t st, View<R> out, View<const R> Q_basis, View<const R> Q_coeff, View<const R> Rm);
template <class R, bool CX>
void qr_inverse_t(cudaStream_t st, View<R> out, View<const R> Q_basis, View<const R> Q_coeff, View<const R> Rm);

// dst(triangle) <- src(triangle)^-1 (triangular_inverse.rs), every scalar kind; only the triangle is written (not the diagonal if unit)
template <class R, bool CX>
void inverse_triangular_t(cudaStream_t st, View<R> dst, View<const R> src, bool lower, bool unit);
// ---- ldlt_types.cu: LDLT beyond the f64 factorization, over the scalar kind <R, complex?> (complex views in COMPLEX element units
// on an R* base). Factor / solve: <float, false>, <double, true>, <float, true>; reconstruct / inverse: also <double, false>.
// D: DEVICE pointer to T-typed entries `dstride` elements apart; d_signs: device int8[n] or null ----
template <class R, bool CX>
LdltResult ldlt_in_place_t(cudaStream_t st, View<R> A, R delta, R eps, const signed char* d_signs);
template <class R, bool CX>
void ldlt_solve_in_place_t(cudaStream_t st, View<const R> L, const R* D, i64 dstride, bool conj, View<R> rhs);
template <class R, bool CX>
void ldlt_reconstruct_t(cudaStream_t st, View<R> out, View<const R> L, const R* D, i64 dstride);
template <class R, bool CX>
void ldlt_inverse_t(cudaStream_t st, View<R> out, View<const R> L, const R* D, i64 dstride);

}  // namespace fb
