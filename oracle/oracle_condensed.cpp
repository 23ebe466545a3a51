// TEST INFRASTRUCTURE — NOT PRODUCT CODE. CPU restatement of faer's reductions to condensed form (SURVEY.md §8a rows
// a8, a9). See oracle.hpp for the contract.
//
// Reference:
//   svd::bidiag::bidiag_in_place     faer/src/linalg/svd/bidiag.rs:47-256 (column loop 68-221, T-factor upgrade 222-255),
//                                    fused op 257-301 (scalar fallback: A22 -= up*y + z*vp ; y = u^H A22)
//   evd::tridiag::tridiag_in_place   faer/src/linalg/evd/tridiag.rs:274-529 (column loop 295-505, T upgrade 506-528),
//                                    fused op 36-160 (lower triangle only: rank-2 update, z += a*uc, y_j = f * sum_{i>j})
// Both keep ONE rank-2 update pending: the vectors (u, y, z, v) produced at step k are applied to the trailing matrix
// inside the fused pass of step k+1, so every column costs one read+write pass (plus one read pass for bidiag).
#include <algorithm>
#include <vector>

#include "oracle.hpp"

namespace oracle {

template <class T> static inline Mat<const T> cst(Mat<T> a) { return Mat<const T>{a.p, a.m, a.n, a.rs, a.cs}; }
template <class T> static inline T cj(T x) { return x; }
template <class R> static inline std::complex<R> cj(std::complex<R> x) { return std::conj(x); }
template <class T> static inline bool is_inf(T x) { return std::isinf(x); }
template <class R> static inline bool is_inf(std::complex<R> x) { return std::isinf(x.real()) || std::isinf(x.imag()); }

// svd/bidiag.rs:47-256. A: m x n with m >= n. Hl: bl x size, Hr: br x (size-1).
template <class T>
void bidiag_in_place(Mat<T> A, Mat<T> Hl, Mat<T> Hr) {
  typedef typename real_of<T>::type R;
  const i64 m = A.m, n = A.n, size = std::min(m, n);
  const i64 bl = Hl.m, br = Hr.m;
  std::vector<T> y(std::max<i64>(n, 1)), z(std::max<i64>(m, 1));
  for (i64 k = 0; k < size; ++k) {
    const i64 mr = m - k - 1, nr = n - k - 1;  // A22 is mr x nr, top-left (k+1, k+1)
    T& a11 = A(k, k);
    if (k > 0) {  // bidiag.rs:86-103: pending update applied to column k and row k
      const T y1 = y[k], z1 = z[k];
      const i64 k1 = k - 1;
      const T up0 = A(k, k1);
      a11 = a11 - (up0 * y1 + z1);
      for (i64 i = 0; i < mr; ++i) A(k + 1 + i, k) = A(k + 1 + i, k) - (A(k + 1 + i, k1) * y1 + z[k + 1 + i]);
      for (i64 j = 0; j < nr; ++j) A(k, k + 1 + j) = A(k, k + 1 + j) - (up0 * y[k + 1 + j] + z1 * A(k1, k + 1 + j));
    }
    const HouseholderInfo<T> hl = make_householder<T>(&a11, mr ? &A(k + 1, k) : &a11, A.rs, nullptr, 0, mr);
    const R tl_inv = R(1) / hl.tau;
    Hl(0, k) = T(hl.tau);
    if (k > 0) {  // fused op, bidiag.rs:282-301
      const i64 k1 = k - 1;
      for (i64 j = 0; j < nr; ++j) {
        const T yj = y[k + 1 + j], vj = A(k1, k + 1 + j);
        T acc = T(0);
        for (i64 i = 0; i < mr; ++i) {
          T a = A(k + 1 + i, k + 1 + j);
          a = a - A(k + 1 + i, k1) * yj;
          a = a - z[k + 1 + i] * vj;
          A(k + 1 + i, k + 1 + j) = a;
          acc = acc + cj(A(k + 1 + i, k)) * a;
        }
        y[k + 1 + j] = acc;
      }
    } else {
      for (i64 j = 0; j < nr; ++j) {
        T acc = T(0);
        for (i64 i = 0; i < mr; ++i) acc = acc + cj(A(k + 1 + i, k)) * A(k + 1 + i, k + 1 + j);
        y[k + 1 + j] = acc;
      }
    }
    for (i64 j = 0; j < nr; ++j) {  // bidiag.rs:151-154
      T& a = A(k, k + 1 + j);
      y[k + 1 + j] = (y[k + 1 + j] + a) * tl_inv;
      a = a - y[k + 1 + j];
    }
    const R norm = nr ? norm_l2<T>(&A(k, k + 1), nr, A.cs) : R(0);
    const R norm_inv = R(1) / norm;
    if (norm != R(0))
      for (i64 j = 0; j < nr; ++j) A(k, k + 1 + j) = A(k, k + 1 + j) * norm_inv;
    for (i64 i = 0; i < mr; ++i) {  // z2 = A22 * A12^H, bidiag.rs:160-167
      T acc = T(0);
      for (i64 j = 0; j < nr; ++j) acc = acc + A(k + 1 + i, k + 1 + j) * cj(A(k, k + 1 + j));
      z[k + 1 + i] = acc;
    }
    if (k + 1 == size) break;
    T& a12a = A(k, k + 1);
    const HouseholderInfo<T> hr = make_householder<T>(&a12a, nr > 1 ? &A(k, k + 2) : &a12a, A.cs, nullptr, 0, nr - 1);
    const R tr_inv = R(1) / hr.tau;
    const T mm = hr.head_with_beta_inv;
    Hr(0, k) = T(hr.tau);
    const T beta = a12a;
    a12a = a12a * norm;
    T b = y[k + 1];
    for (i64 j = 1; j < nr; ++j) b = b + y[k + 1 + j] * cj(A(k, k + 1 + j));
    if (!is_inf(mm)) {
      for (i64 i = 0; i < mr; ++i) {
        T w = z[k + 1 + i] - A(k + 1 + i, k + 1) * cj(beta);
        w = w * cj(mm);
        w = w - A(k + 1 + i, k) * b;
        z[k + 1 + i] = w * tr_inv;
      }
    } else {
      for (i64 i = 0; i < mr; ++i) z[k + 1 + i] = (A(k + 1 + i, k + 1) - A(k + 1 + i, k) * b) * tr_inv;
    }
  }
  // T factors, bidiag.rs:222-255
  for (i64 j = 0; j < size;) {
    const i64 b = std::min(bl, size - j);
    if (b <= 0) break;
    Mat<T> H = Hl.sub(0, j, b, b);
    for (i64 k = 0; k < b; ++k) H(k, k) = H(0, k);
    upgrade_householder_factor<T>(H, cst(A.sub(j, j, m - j, b)), b, 1);
    j += b;
  }
  if (size > 0) {
    const i64 s1 = size - 1;
    Mat<T> At = A.sub(0, 1, s1, n - 1).t();  // (n-1) x s1: reflector k is column k, starting at row k
    for (i64 j = 0; j < s1;) {
      const i64 b = std::min(br, s1 - j);
      if (b <= 0) break;
      Mat<T> H = Hr.sub(0, j, b, b);
      for (i64 k = 0; k < b; ++k) H(k, k) = H(0, k);
      upgrade_householder_factor<T>(H, cst(At.sub(j, j, At.m - j, b)), b, 1);
      j += b;
    }
  }
}

// evd/tridiag.rs:274-529. A: n x n self-adjoint, lower triangle read/written. H: b x (n-1).
template <class T>
void tridiag_in_place(Mat<T> A, Mat<T> H) {
  typedef typename real_of<T>::type R;
  const i64 n = A.m, bs = H.m;
  if (n == 0) return;
  std::vector<T> y(n), w(n), z(n);
  for (i64 k = 0; k < n; ++k) {
    if (k > 0) {  // tridiag.rs:307-317
      const T y1 = y[k];
      const i64 p = k - 1;
      A(k, k) = A(k, k) - (y1 + cj(y1));
      for (i64 i = k + 1; i < n; ++i) A(i, k) = A(i, k) - (cj(y1) * A(i, p) + y[i]);
    }
    if (k + 1 == n) break;
    const i64 k1 = k + 1;
    const i64 len = n - k - 2;  // x2 = A[k+2.., k]
    T* x2 = len ? &A(k + 2, k) : &A(k1, k);
    const HouseholderInfo<T> hi = make_householder<T>(&A(k1, k), x2, A.rs, nullptr, 0, len);
    const R tau_inv = R(1) / hi.tau;
    H(0, k) = T(hi.tau);
    T& y1 = y[k1];
    T& a11 = A(k1, k1);
    auto X = [&](i64 i) -> T { return A(k + 2 + i, k); };  // x2[i]
    if (k > 0) {
      const i64 p = k - 1;
      const T u1 = A(k1, p);
      a11 = a11 - (u1 * cj(y1) + y1 * cj(u1));
      for (i64 i = 0; i < len; ++i) A(k + 2 + i, k1) = A(k + 2 + i, k1) - (A(k + 2 + i, p) * cj(y1) + y[k + 2 + i] * cj(u1));
      for (i64 i = 0; i < len; ++i) w[i] = y[k + 2 + i];
      // fused op on the lower triangle of A22' = A[k+2.., k+2..], tridiag.rs:84-118 / 161-…
      const T f = T(tau_inv);
      for (i64 i = 0; i < len; ++i) z[i] = T(0);
      for (i64 j = 0; j < len; ++j) {
        const T ry = -w[j], ub = -A(k + 2 + j, p), uc = f * X(j);
        T acc = T(0);
        for (i64 i = j; i < len; ++i) {
          T a = A(k + 2 + i, k + 2 + j);
          a = a + cj(ry) * A(k + 2 + i, p);
          a = a + cj(ub) * w[i];
          A(k + 2 + i, k + 2 + j) = a;
          z[i] = z[i] + a * uc;
          acc = acc + cj(a) * X(i);
        }
        y[k + 2 + j] = f * (acc - A(k + 2 + j, k + 2 + j) * X(j));
      }
      for (i64 i = 0; i < len; ++i) y[k + 2 + i] = y[k + 2 + i] + z[i];
    } else {
      // y2 = tau_inv * (tril(A22) x2 + striu(A22^H) x2), tridiag.rs:440-461
      for (i64 i = 0; i < len; ++i) {
        T acc = T(0);
        for (i64 j = 0; j <= i; ++j) acc = acc + A(k + 2 + i, k + 2 + j) * X(j);
        T acc2 = T(0);
        for (i64 j = i + 1; j < len; ++j) acc2 = acc2 + cj(A(k + 2 + j, k + 2 + i)) * X(j);
        y[k + 2 + i] = T(tau_inv) * acc + T(tau_inv) * acc2;
      }
    }
    for (i64 i = 0; i < len; ++i) y[k + 2 + i] = y[k + 2 + i] + A(k + 2 + i, k1) * tau_inv;
    T d = T(0);
    for (i64 i = 0; i < len; ++i) d = d + cj(A(k + 2 + i, k1)) * X(i);
    y1 = (a11 + d) * tau_inv;
    T d2 = T(0);
    for (i64 i = 0; i < len; ++i) d2 = d2 + cj(X(i)) * y[k + 2 + i];
    const T b = (y1 + d2) * R(0.5) * tau_inv;
    y1 = y1 - b;
    for (i64 i = 0; i < len; ++i) y[k + 2 + i] = y[k + 2 + i] - b * X(i);
  }
  {
    const i64 n1 = n - 1;
    Mat<T> As = A.sub(1, 0, n1, n1);
    for (i64 j = 0; j < n1;) {
      const i64 b = std::min(bs, n1 - j);
      if (b <= 0) break;
      Mat<T> Hb = H.sub(0, j, b, b);
      for (i64 k = 0; k < b; ++k) Hb(k, k) = Hb(0, k);
      upgrade_householder_factor<T>(Hb, cst(As.sub(j, j, n1 - j, b)), b, 1);
      j += b;
    }
  }
}

#define ORACLE_COND_INST(T)                              \
  template void bidiag_in_place<T>(Mat<T>, Mat<T>, Mat<T>); \
  template void tridiag_in_place<T>(Mat<T>, Mat<T>);
ORACLE_COND_INST(double)
ORACLE_COND_INST(float)
ORACLE_COND_INST(std::complex<double>)
ORACLE_COND_INST(std::complex<float>)

}  // namespace oracle
