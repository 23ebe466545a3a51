// TEST INFRASTRUCTURE — NOT PRODUCT CODE. C entry points of the CPU oracle for ctypes (tests/, smoke(), bench.py
// cpu_baseline only). dtype: 0 = f32, 1 = f64, 2 = c32, 3 = c64. Scalars are passed by pointer and read by type.
#include <omp.h>

#include "oracle.hpp"

using namespace oracle;

struct OMat {
  void* p;
  long long m, n, rs, cs;
};
template <class T> static Mat<T> mm(OMat a) { return Mat<T>{(T*)a.p, a.m, a.n, a.rs, a.cs}; }
template <class T> static Mat<const T> mc(OMat a) { return Mat<const T>{(const T*)a.p, a.m, a.n, a.rs, a.cs}; }

extern "C" {

int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }

#define DISPATCH(dtype, CALL)                                         \
  switch (dtype) {                                                    \
    case 0: { typedef float T; CALL; } break;                         \
    case 1: { typedef double T; CALL; } break;                        \
    case 2: { typedef std::complex<float> T; CALL; } break;           \
    case 3: { typedef std::complex<double> T; CALL; } break;          \
    default: return -100;                                             \
  }


long long oracle_matmul(int dtype, OMat dst, int add, OMat lhs, int conj_lhs, OMat rhs, int conj_rhs, const void* alpha) {
  DISPATCH(dtype, matmul<T>(mm<T>(dst), add != 0, mc<T>(lhs), conj_lhs != 0, mc<T>(rhs), conj_rhs != 0, *(const T*)alpha));
  return 0;
}

long long oracle_matmul_triangular(int dtype, OMat dst, int dst_s, int add, OMat lhs, int lhs_s, int conj_lhs, OMat rhs,
                                   int rhs_s, int conj_rhs, const void* alpha) {
  DISPATCH(dtype, matmul_triangular<T>(mm<T>(dst), dst_s, add != 0, mc<T>(lhs), lhs_s, conj_lhs != 0, mc<T>(rhs), rhs_s,
                                       conj_rhs != 0, *(const T*)alpha));
  return 0;
}

long long oracle_solve_triangular(int dtype, int lower, int unit, OMat tri, int conj, OMat rhs) {
  if (lower) {
    DISPATCH(dtype, solve_lower<T>(mc<T>(tri), conj != 0, unit != 0, mm<T>(rhs)));
  } else {
    DISPATCH(dtype, solve_upper<T>(mc<T>(tri), conj != 0, unit != 0, mm<T>(rhs)));
  }
  return 0;
}

// returns -1 on success, else the failing column; *reg_count receives the regularisation count.
long long oracle_llt(int dtype, OMat A, double delta, double eps, long long recursion_threshold, long long block_size,
                     long long* reg_count) {
  long long r = -100;
  DISPATCH(dtype, r = llt_in_place<T>(mm<T>(A), (real_of<T>::type)delta, (real_of<T>::type)eps, recursion_threshold,
                                      block_size, reg_count));
  return r;
}

// LDLT: returns -1 on success, else the ZeroPivot index; signs may be null.
long long oracle_ldlt(int dtype, OMat A, double delta, double eps, const signed char* signs, long long recursion_threshold,
                      long long block_size, long long* reg_count) {
  long long r = -100;
  DISPATCH(dtype, r = ldlt_in_place<T>(mm<T>(A), (real_of<T>::type)delta, (real_of<T>::type)eps, signs, recursion_threshold,
                                       block_size, reg_count));
  return r;
}

// perm / perm_inv: int64[nrows]; returns the transposition count.
long long oracle_lu(int dtype, OMat A, long long* perm, long long* perm_inv, long long recursion_threshold) {
  long long r = -100;
  DISPATCH(dtype, r = lu_in_place<T>(mm<T>(A), perm, perm_inv, recursion_threshold));
  return r;
}

// Householder QR (no pivoting). Q_coeff: block_size x min(m,n). returns the rank.
long long oracle_qr(int dtype, OMat A, OMat Q_coeff, long long blocking_threshold) {
  long long r = -100;
  DISPATCH(dtype, r = qr_in_place<T>(mm<T>(A), mm<T>(Q_coeff), blocking_threshold));
  return r;
}
long long oracle_qr_recommended_block_size(long long nrows, long long ncols) {
  return qr_recommended_block_size(nrows, ncols);
}
// M <- (I - V T^-1 V^H) M (forward = 0) or (I - V T^-H V^H) M (forward = 1); conj_lhs conjugates V and T.
long long oracle_apply_block_householder_left(int dtype, OMat V, OMat Tf, int conj_lhs, OMat M, int forward) {
  DISPATCH(dtype, apply_block_householder_on_the_left<T>(mc<T>(V), mc<T>(Tf), conj_lhs != 0, mm<T>(M), forward != 0));
  return 0;
}
// reductions to condensed form (oracle_condensed.cpp)
long long oracle_bidiag(int dtype, OMat A, OMat Hl, OMat Hr) {
  DISPATCH(dtype, bidiag_in_place<T>(mm<T>(A), mm<T>(Hl), mm<T>(Hr)));
  return 0;
}
long long oracle_tridiag(int dtype, OMat A, OMat H) {
  DISPATCH(dtype, tridiag_in_place<T>(mm<T>(A), mm<T>(H)));
  return 0;
}
double oracle_norm_l2(int dtype, const void* p, long long n, long long stride) {
  switch (dtype) {
    case 0: return norm_l2<float>((const float*)p, n, stride);
    case 1: return norm_l2<double>((const double*)p, n, stride);
    case 2: return norm_l2<std::complex<float>>((const std::complex<float>*)p, n, stride);
    case 3: return norm_l2<std::complex<double>>((const std::complex<double>*)p, n, stride);
  }
  return -1.0;
}

}  // extern "C"
