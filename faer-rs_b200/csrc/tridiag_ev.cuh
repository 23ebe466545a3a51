// Eigenvalues of a real symmetric tridiagonal matrix (diagonal d[0..n), off-diagonal e[0..n-1)) by Sturm-count bisection,
// one independent scalar problem per eigenvalue (the GPU kernel in evd.cu runs one thread per value). Plain host/device
// code: tests/test_tridiag_ev_cpu.py compiles this header with g++ and checks it against LAPACK. The reference reaches
// the same values through tridiag_evd::qr_algorithm (evd/mod.rs:345-353, values-only branch); parity is to n * eps * |T|.
#pragma once
#include <cfloat>
#include <cmath>

#ifndef FB_HD
#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif
#endif

namespace fb {
namespace tev {

template <class T> struct Lim;
template <> struct Lim<double> {
  static FB_HD double eps() { return DBL_EPSILON; }
  static FB_HD double safmin() { return DBL_MIN; }
};
template <> struct Lim<float> {
  static FB_HD float eps() { return FLT_EPSILON; }
  static FB_HD float safmin() { return FLT_MIN; }
};

// number of eigenvalues < x (LAPACK dstebz recurrence with a pivmin guard)
template <class T>
FB_HD int st_negcount(const T* d, const T* e, int n, T x, T pivmin) {
  T q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  int cnt = q < T(0) ? 1 : 0;
  for (int i = 1; i < n; ++i) {
    const T b = e[i - 1];
    q = d[i] - x - (b * b) / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < T(0) ? 1 : 0;
  }
  return cnt;
}

// k-th SMALLEST eigenvalue (k = 0 .. n-1); [glo, ghi] contains the spectrum (Gershgorin), emax2 = max e_i^2
template <class T>
FB_HD T st_kth_smallest(const T* d, const T* e, int n, int k, T glo, T ghi, T emax2) {
  const T eps = Lim<T>::eps();
  const T pivmin = Lim<T>::safmin() * (emax2 > T(1) ? emax2 : T(1));
  T lo = glo, hi = ghi;
  for (int it = 0; it < 1200; ++it) {
    const T mid = lo + (hi - lo) * T(0.5);
    if (mid <= lo || mid >= hi) break;
    const T mag = fabs(lo) > fabs(hi) ? fabs(lo) : fabs(hi);
    if (hi - lo <= T(2) * eps * mag + pivmin) break;
    if (st_negcount<T>(d, e, n, mid, pivmin) <= k) lo = mid;
    else hi = mid;
  }
  return lo + (hi - lo) * T(0.5);
}

}  // namespace tev
}  // namespace fb
