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

// perm / perm_inv: int64[nrows]; returns the transposition count.
long long oracle_lu(int dtype, OMat A, long long* perm, long long* perm_inv, long long recursion_threshold) {
  long long r = -100;
  DISPATCH(dtype, r = lu_in_place<T>(mm<T>(A), perm, perm_inv, recursion_threshold));
  return r;
}

}  // extern "C"
