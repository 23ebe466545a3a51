// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU restatement ("oracle") of the reference algorithms on the hot path, used only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker.
// Nothing under faer-rs_b200/ may include, link or call this.
//
// PARITY PINNING: faer (Rust) cannot be built in this environment (no rustc/cargo, GEMM arithmetic lives in
// un-vendored crates private-gemm-x86 0.1.20 / nano-gemm 0.2.2 / gemm 0.19.0, Cargo.lock:2129, 1639, 980), and
// the reference tree holds NO bit-level golden outputs for GEMM/LLT/LU/QR. The oracle is pinned by
//   (1) the reference's own test identities and tolerances (tests/test_oracle_*.py restate them), and
//   (2) the few known-answer fixtures the reference has (qr/mod.rs:118-146 lstsq; see tests/golden/).
// Bitwise parity with faer's AVX GEMM is therefore "unpinned"; tolerance parity is pinned.
//
// Every function cites the reference file:line it follows. Views are faer's 5-field strided views.
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace oracle {

typedef long long i64;

template <class T>
struct Mat {
  T* p;
  i64 m, n, rs, cs;
  T& operator()(i64 i, i64 j) const { return p[i * rs + j * cs]; }
  Mat sub(i64 i, i64 j, i64 mm, i64 nn) const { return Mat{p + i * rs + j * cs, mm, nn, rs, cs}; }
  Mat t() const { return Mat{p, n, m, cs, rs}; }
  Mat rev_rows() const { return Mat{p + (m - 1) * rs, m, n, -rs, cs}; }
  Mat rev_rows_cols() const { return Mat{p + (m - 1) * rs + (n - 1) * cs, m, n, -rs, -cs}; }
};

template <class T> struct real_of { typedef T type; };
template <class R> struct real_of<std::complex<R>> { typedef R type; };

template <class T> inline T conj_if(bool c, T x) { return x; }
template <class R> inline std::complex<R> conj_if(bool c, std::complex<R> x) { return c ? std::conj(x) : x; }
template <class T> inline typename real_of<T>::type real_part(T x) { return x; }
template <class R> inline R real_part(std::complex<R> x) { return x.real(); }
// abs1: |x| for reals, |re|+|im| for complex (faer-traits/src/lib.rs:2151-2153, 2643-2646)
template <class T> inline T abs1(T x) { return std::fabs(x); }
template <class R> inline R abs1(std::complex<R> x) { return std::fabs(x.real()) + std::fabs(x.imag()); }

enum Structure { RECT = 0, TRI_LOWER = 1, TRI_UPPER = 2, STRICT_LOWER = 3, STRICT_UPPER = 4, UNIT_LOWER = 5, UNIT_UPPER = 6 };

// matmul: faer/src/linalg/matmul/mod.rs:1617-1660; scalar definition 1505-1523 / 1909-1947
template <class T>
void matmul(Mat<T> dst, bool add, Mat<const T> lhs, bool conj_lhs, Mat<const T> rhs, bool conj_rhs, T alpha);

// triangular matmul: faer/src/linalg/matmul/triangular.rs:1193-1245 (+ semantics 26-52, 906-977)
template <class T>
void matmul_triangular(Mat<T> dst, int dst_s, bool add, Mat<const T> lhs, int lhs_s, bool conj_lhs, Mat<const T> rhs,
                       int rhs_s, bool conj_rhs, T alpha);

// triangular solves: faer/src/linalg/triangular_solve.rs:420-604
template <class T> void solve_lower(Mat<const T> tril, bool conj, bool unit, Mat<T> rhs);
template <class T> void solve_upper(Mat<const T> triu, bool conj, bool unit, Mat<T> rhs);

// LLT: faer/src/linalg/cholesky/llt/factor.rs:68-97 -> ldlt/factor.rs:367-498 (+ leaf 7-177)
// returns -1 on success (count in *reg_count) or the failing column index.
template <class T>
i64 llt_in_place(Mat<T> A, typename real_of<T>::type delta, typename real_of<T>::type eps, i64 recursion_threshold,
                 i64 block_size, i64* reg_count);

// LDLT: faer/src/linalg/cholesky/ldlt/factor.rs:725-767 (cholesky_in_place), 367-498 (recursion, is_llt = false), 7-177 (leaf).
// Returns -1 on success, else the ZeroPivot index. D ends on the diagonal of A, unit-lower L strictly below it.
// signs: optional expected signs of the pivots (+1 / -1 / 0), used by the dynamic regularisation only.
template <class T>
i64 ldlt_in_place(Mat<T> A, typename real_of<T>::type delta, typename real_of<T>::type eps, const signed char* signs,
                  i64 recursion_threshold, i64 block_size, i64* reg_count);

// LU: faer/src/linalg/lu/partial_pivoting/factor.rs:234-295 (recursion 68-187, leaf 19-67)
// perm / perm_inv: i64 arrays of length nrows. returns transposition count.
template <class T>
i64 lu_in_place(Mat<T> A, i64* perm, i64* perm_inv, i64 recursion_threshold);

// ---- oracle_qr.cpp ----
// norm_l2: faer/src/linalg/reductions/norm_l2.rs:6-172
template <class T> typename real_of<T>::type norm_l2(const T* p, i64 n, i64 stride);
// householder::make_householder_imp: faer/src/linalg/householder.rs:59-107 (`in == nullptr` -> in place)
template <class T>
struct HouseholderInfo {
  typename real_of<T>::type tau;
  T head_with_beta_inv;
  typename real_of<T>::type norm;
};
template <class T> HouseholderInfo<T> make_householder(T* head, T* out, i64 out_stride, const T* in, i64 in_stride, i64 len);
// householder::upgrade_householder_factor: householder.rs:132-272
template <class T> void upgrade_householder_factor(Mat<T> Tf, Mat<const T> V, i64 block_size, i64 prev_block_size);
// Householder QR without pivoting: faer/src/linalg/qr/no_pivoting/factor.rs:258-301; Q_coeff is block_size x min(m,n);
// returns the numerical rank.
template <class T> i64 qr_in_place(Mat<T> A, Mat<T> Q_coeff, i64 blocking_threshold);
i64 qr_recommended_block_size(i64 nrows, i64 ncols);
// block reflector application: faer/src/linalg/householder.rs:370-620
template <class T>
void apply_block_householder_on_the_left(Mat<const T> V, Mat<const T> Tf, bool conj_lhs, Mat<T> M, bool forward);

// ---- oracle_condensed.cpp ----
// bidiagonalization A = U B V^H (m >= n): faer/src/linalg/svd/bidiag.rs:47-256. Hl: bl x n, Hr: br x (n-1).
template <class T> void bidiag_in_place(Mat<T> A, Mat<T> Hl, Mat<T> Hr);
// tridiagonalization A = Q T Q^H (lower triangle): faer/src/linalg/evd/tridiag.rs:274-529. H: b x (n-1).
template <class T> void tridiag_in_place(Mat<T> A, Mat<T> H);

}  // namespace oracle
