// Singular values of a real upper-bidiagonal matrix (diagonal d[0..n), superdiagonal e[0..n-1)) by bisection on the
// Golub-Kahan tridiagonal form: T_GK = P [0 B^T; B 0] P^T has a zero diagonal, off-diagonals b = (d_0, e_0, d_1, e_1, ..., d_{n-1})
// and eigenvalues +-sigma_i, and Sturm counts on it determine every singular value to high relative accuracy
// (Demmel & Kahan 1990). One independent scalar problem per singular value: the GPU kernel (svd.cu) runs one thread per value.
//
// Written as plain host/device functions so that the very same code is checked on the CPU (tests/test_bidiag_sv_cpu.py
// compiles this header with g++) — the reference reaches the same values through bidiag_svd (svd/bidiag_svd.rs: QR iteration /
// divide and conquer); parity is to n * eps * sigma_max, the tolerance of the reference's own SVD tests.
#pragma once
#include <cfloat>
#include <cmath>

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

namespace fb {
namespace bsv {

template <class T> struct Lim;
template <> struct Lim<double> {
  static FB_HD double eps() { return DBL_EPSILON; }
  static FB_HD double safmin() { return DBL_MIN; }
};
template <> struct Lim<float> {
  static FB_HD float eps() { return FLT_EPSILON; }
  static FB_HD float safmin() { return FLT_MIN; }
};

// number of eigenvalues of T_GK that are < x (x > 0): n + #{sigma_i < x}. LAPACK dstebz-style recurrence with a pivmin guard.
template <class T>
FB_HD int gk_negcount(const T* d, const T* e, int n, long long dstride, long long estride, T x, T pivmin) {
  T q = -x;
  if (fabs(q) < pivmin) q = -pivmin;
  int cnt = q < T(0) ? 1 : 0;
  for (int i = 0; i < n; ++i) {
    const T bd = d[(long long)i * dstride];
    q = -x - (bd * bd) / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < T(0) ? 1 : 0;
    if (i + 1 < n) {
      const T be = e[(long long)i * estride];
      q = -x - (be * be) / q;
      if (fabs(q) < pivmin) q = -pivmin;
      cnt += q < T(0) ? 1 : 0;
    }
  }
  return cnt;
}

// k-th LARGEST singular value (k = 0 .. n-1). `bound` >= sigma_max (gk_bound), `bmax2` = max b_j^2.
template <class T>
FB_HD T gk_kth_largest(const T* d, const T* e, int n, long long dstride, long long estride, int k, T bound, T bmax2) {
  const T eps = Lim<T>::eps();
  T pivmin = Lim<T>::safmin() * (bmax2 > T(1) ? bmax2 : T(1));
  const int idx = n - 1 - k;  // ascending index: sigma_(idx) is the smallest x with #{sigma < x'} > idx for all x' > x
  T lo = T(0), hi = bound;
  for (int it = 0; it < 1200; ++it) {
    const T mid = lo + (hi - lo) * T(0.5);
    if (mid <= lo || mid >= hi) break;
    if (hi - lo <= T(2) * eps * hi + pivmin) break;
    const int below = gk_negcount<T>(d, e, n, dstride, estride, mid, pivmin) - n;  // #{sigma_i < mid}
    if (below <= idx) lo = mid;
    else hi = mid;
  }
  return lo + (hi - lo) * T(0.5);
}

}  // namespace bsv
}  // namespace fb
