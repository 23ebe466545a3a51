// The building blocks of the flat-map drivers forwarded to a callback (see drivers_host.cpp / ffi_types_host.cpp): structured
// products, triangular solves, Householder sequences and the real condensed solvers are executed by the test with the oracle / LAPACK.
// TEST INFRASTRUCTURE.
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../faer-rs_b200/csrc/runtime.cuh"
#include "../../faer-rs_b200/csrc/tensor_ops.cuh"

// ---- the callback protocol -------------------------------------------------------------------------------------------------------
extern "C" {
struct MockMat {
  void* ptr;
  long long nrows, ncols, rs, cs;
  int is_double, is_complex;
};
struct MockCall {
  int op;  // 1 gemm, 2 trsm, 3 householder sequence, 4 tridiagonal eigensolver, 5 bidiagonal svd
  MockMat m[5];
  long long i[8];
  double d[4];
  long long ret;
};
typedef void (*mock_cb_t)(MockCall*);
}
static mock_cb_t g_cb = nullptr;


namespace fb {

// ---- building blocks -> callback ----
namespace {
template <class T>
MockMat mm(View<T> v, bool cx) {
  return MockMat{(void*)v.ptr, v.nrows, v.ncols, v.rs, v.cs, sizeof(T) == 8 ? 1 : 0, cx ? 1 : 0};
}
template <class R>
void gemm_cb(View<R> dst, int ds, int accum, View<const R> a, int as, bool ca, View<const R> b, int bs, bool cb, double ar, double ai, bool cx) {
  MockCall c{};
  c.op = 1;
  c.m[0] = mm(dst, cx); c.m[1] = mm(a, cx); c.m[2] = mm(b, cx);
  c.i[0] = ds; c.i[1] = accum; c.i[2] = as; c.i[3] = ca; c.i[4] = bs; c.i[5] = cb;
  c.d[0] = ar; c.d[1] = ai;
  g_cb(&c);
}
template <class R>
void trsm_cb(View<const R> t, bool lower, bool unit, bool conj, View<R> rhs, bool cx) {
  MockCall c{};
  c.op = 2;
  c.m[0] = mm(t, cx); c.m[1] = mm(rhs, cx);
  c.i[0] = lower; c.i[1] = unit; c.i[2] = conj;
  g_cb(&c);
}
template <class R>
void hhseq_cb(View<const R> basis, View<const R> factor, bool conj, View<R> rhs, bool transpose, bool cx) {
  MockCall c{};
  c.op = 3;
  c.m[0] = mm(basis, cx); c.m[1] = mm(factor, cx); c.m[2] = mm(rhs, cx);
  c.i[0] = conj; c.i[1] = transpose;
  g_cb(&c);
}
}  // namespace

void gemm_f64(cudaStream_t, VD dst, int ds, int accum, VCD a, int as, VCD b, int bs, double alpha) { gemm_cb<double>(dst, ds, accum, a, as, false, b, bs, false, alpha, 0, false); }
void gemm_f32(cudaStream_t, VF dst, int ds, int accum, VCF a, int as, VCF b, int bs, float alpha) { gemm_cb<float>(dst, ds, accum, a, as, false, b, bs, false, alpha, 0, false); }
void gemm_c64(cudaStream_t, VD dst, int ds, int accum, VCD a, int as, bool ca, VCD b, int bs, bool cb, double ar, double ai) { gemm_cb<double>(dst, ds, accum, a, as, ca, b, bs, cb, ar, ai, true); }
void gemm_c32(cudaStream_t, VF dst, int ds, int accum, VCF a, int as, bool ca, VCF b, int bs, bool cb, float ar, float ai) { gemm_cb<float>(dst, ds, accum, a, as, ca, b, bs, cb, ar, ai, true); }
void solve_lower_triangular_in_place_f64(cudaStream_t, VCD t, bool unit, VD rhs) { trsm_cb<double>(t, true, unit, false, rhs, false); }
void solve_upper_triangular_in_place_f64(cudaStream_t, VCD t, bool unit, VD rhs) { trsm_cb<double>(t, false, unit, false, rhs, false); }
void solve_lower_triangular_in_place_f32(cudaStream_t, VCF t, bool unit, VF rhs) { trsm_cb<float>(t, true, unit, false, rhs, false); }
void solve_upper_triangular_in_place_f32(cudaStream_t, VCF t, bool unit, VF rhs) { trsm_cb<float>(t, false, unit, false, rhs, false); }
void solve_lower_triangular_in_place_c64(cudaStream_t, VCD t, bool unit, bool conj, VD rhs) { trsm_cb<double>(t, true, unit, conj, rhs, true); }
void solve_upper_triangular_in_place_c64(cudaStream_t, VCD t, bool unit, bool conj, VD rhs) { trsm_cb<double>(t, false, unit, conj, rhs, true); }
void solve_lower_triangular_in_place_c32(cudaStream_t, VCF t, bool unit, bool conj, VF rhs) { trsm_cb<float>(t, true, unit, conj, rhs, true); }
void solve_upper_triangular_in_place_c32(cudaStream_t, VCF t, bool unit, bool conj, VF rhs) { trsm_cb<float>(t, false, unit, conj, rhs, true); }
template <class T>
void apply_block_householder_sequence_on_the_left(cudaStream_t, View<const T> basis, View<const T> factor, View<T> M) { hhseq_cb<T>(basis, factor, false, M, false, false); }
template <class T>
void apply_block_householder_sequence_transpose_on_the_left(cudaStream_t, View<const T> basis, View<const T> factor, View<T> M) { hhseq_cb<T>(basis, factor, false, M, true, false); }
template void apply_block_householder_sequence_on_the_left<float>(cudaStream_t, View<const float>, View<const float>, View<float>);
template void apply_block_householder_sequence_on_the_left<double>(cudaStream_t, View<const double>, View<const double>, View<double>);
template void apply_block_householder_sequence_transpose_on_the_left<float>(cudaStream_t, View<const float>, View<const float>, View<float>);
template void apply_block_householder_sequence_transpose_on_the_left<double>(cudaStream_t, View<const double>, View<const double>, View<double>);
void apply_householder_sequence_left_c64(cudaStream_t, VCD basis, VCD factor, bool conj, VD rhs, bool transpose) { hhseq_cb<double>(basis, factor, conj, rhs, transpose, true); }
void apply_householder_sequence_left_c32(cudaStream_t, VCF basis, VCF factor, bool conj, VF rhs, bool transpose) { hhseq_cb<float>(basis, factor, conj, rhs, transpose, true); }

bool tridiag_dc_f64(cudaStream_t, const double* d, const double* e, i64 n, double* lam, double* Q, i64 ldq) {
  MockCall c{};
  c.op = 4;
  c.m[0] = MockMat{(void*)d, n, 1, 1, n, 1, 0}; c.m[1] = MockMat{(void*)e, n, 1, 1, n, 1, 0};
  c.m[2] = MockMat{(void*)lam, n, 1, 1, n, 1, 0}; c.m[3] = MockMat{(void*)Q, n, n, 1, ldq, 1, 0};
  g_cb(&c);
  return c.ret != 0;
}
bool bidiag_svd_vectors_f64(cudaStream_t, const double* d, const double* e, i64 n, double* S, double* UB, double* VB) {
  MockCall c{};
  c.op = 5;
  c.m[0] = MockMat{(void*)d, n, 1, 1, n, 1, 0}; c.m[1] = MockMat{(void*)e, n, 1, 1, n, 1, 0};
  c.m[2] = MockMat{(void*)S, n, 1, 1, n, 1, 0}; c.m[3] = MockMat{(void*)UB, n, n, 1, n, 1, 0}; c.m[4] = MockMat{(void*)VB, n, n, 1, n, 1, 0};
  g_cb(&c);
  return c.ret != 0;
}
// the HBM-bound reductions of bidiag.cu / tridiag.cu: op 6 / 7
template <class T>
void bidiag_in_place(cudaStream_t, View<T> A, View<T> Hl, View<T> Hr) {
  FB_ASSERT(A.rs == 1, "bidiag_in_place: column-major (row stride 1) matrix required");  // the kernels' own precondition
  MockCall c{};
  c.op = 6;
  c.m[0] = mm(A, false); c.m[1] = mm(Hl, false); c.m[2] = mm(Hr, false);
  g_cb(&c);
}
template <class T>
void tridiag_in_place(cudaStream_t, View<T> A, View<T> H) {
  FB_ASSERT(A.rs == 1, "tridiag_in_place: column-major (row stride 1) matrix required");
  MockCall c{};
  c.op = 7;
  c.m[0] = mm(A, false); c.m[1] = mm(H, false);
  g_cb(&c);
}
template void bidiag_in_place<double>(cudaStream_t, View<double>, View<double>, View<double>);
template void bidiag_in_place<float>(cudaStream_t, View<float>, View<float>, View<float>);
template void tridiag_in_place<double>(cudaStream_t, View<double>, View<double>);
template void tridiag_in_place<float>(cudaStream_t, View<float>, View<float>);
template <class T>
bool device_all_finite(cudaStream_t, const T* x, i64 n) {
  for (i64 i = 0; i < n; ++i)
    if (!(x[i] - x[i] == T(0))) return false;
  return true;
}
template bool device_all_finite<double>(cudaStream_t, const double*, i64);

}  // namespace fb
