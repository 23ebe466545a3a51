// LDLT without pivoting for every scalar kind <R, CX> (R = real type, CX = interleaved complex): the per-thread bodies and the
// launch sequence of an unblocked right-looking factorization as flat maps (flat_map.cuh), shared by the CUDA build (ldlt_types.cu:
// f32 / c64 / c32; the f64 factorization stays on the tuned kernels of ldlt_f64.cu) and a host build that runs the same sequence
// thread by thread (tools/emul/ldlt_host.cpp, tests/test_ldlt_types_emul_cpu.py).
//
// Reference: cholesky/ldlt/factor.rs:725-767 (driver: D on the diagonal, unit-lower L strictly below, strict upper triangle
// untouched; ZeroPivot { index } with A(i, i) = D[i] for i <= index), the leaf recurrence 299-366 and the dynamic regularisation
// 122-144 (sign +1 and d <= eps -> delta, counted; sign -1 and d >= -eps -> -delta; no sign and |d| <= eps -> -delta for d < 0,
// else delta; d = 0 or non-finite -> ZeroPivot). The reference recurses (right-looking blocks on `spicy_matmul`); here one column
// at a time: same L and D up to rounding, same failure index and regularisation count on the same pivots.
// Also the small bodies of `ldlt::solve` (ldlt/solve.rs:11-49), `reconstruct` (reconstruct.rs:9-55) and `inverse` (inverse.rs:9-60).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define LD_HD __host__ __device__ __forceinline__
#else
#define LD_HD inline
#endif

namespace ldl {

typedef long long i64;

template <class R>
struct El {
  R re, im;
};
// element access by scalar kind: W = 1 (real) or 2 (interleaved complex) R values per element, strides in elements
template <class R, bool CX>
LD_HD El<R> ld(const R* p, i64 off) {
  El<R> e;
  if (CX) { e.re = p[2 * off]; e.im = p[2 * off + 1]; }
  else { e.re = p[off]; e.im = R(0); }
  return e;
}
template <class R, bool CX>
LD_HD void st(R* p, i64 off, El<R> e) {
  if (CX) { p[2 * off] = e.re; p[2 * off + 1] = e.im; }
  else p[off] = e.re;
}
LD_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
LD_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
template <class R>
LD_HD El<R> mul(El<R> a, El<R> b) {
  El<R> r;
  r.re = fma_(a.re, b.re, -a.im * b.im);
  r.im = fma_(a.re, b.im, a.im * b.re);
  return r;
}
template <class R>
LD_HD El<R> conj(El<R> a) { a.im = -a.im; return a; }
template <class R>
LD_HD El<R> scale(El<R> a, R s) { a.re *= s; a.im *= s; return a; }
template <class R>
LD_HD bool finite_r(R x) { return x - x == R(0); }

// info[0] = failure index (-1: none), info[1] = regularisation count; sc[0] = 1 / D[j] of the current column
enum { INFO_FAIL = 0, INFO_COUNT = 1 };

// the pivot of column j (one thread): regularisation, D[j], failure test, reciprocal
template <class R, bool CX>
struct Pivot {
  const R* A; i64 rs, cs, j; R* D; const signed char* signs; R delta, eps; int regularize; i64* info; R* sc;
  LD_HD void operator()(i64 t, i64) const {
    if (t != 0 || info[INFO_FAIL] >= 0) return;
    R d = ld<R, CX>(A, j * rs + j * cs).re;
    if (regularize) {
      const int sign = signs ? (int)signs[j] : 0;
      const bool small_or_negative = d <= eps, minus_small_or_positive = d >= -eps;
      if (sign == 1 && small_or_negative) { d = delta; info[INFO_COUNT] += 1; }
      else if (sign == -1 && minus_small_or_positive) d = -delta;
      else if (small_or_negative && minus_small_or_positive) d = d < R(0) ? -delta : delta;
    }
    D[j] = d;
    if (d == R(0) || !finite_r(d)) { info[INFO_FAIL] = j; return; }
    sc[0] = R(1) / d;
  }
};
// column j below the diagonal: w(i) = a(i, j) (= l(i, j) d(j)), a(i, j) <- l(i, j)
template <class R, bool CX>
struct ScaleCol {
  R* A; i64 rs, cs, j, n; const i64* info; const R* sc; R* w;
  LD_HD void operator()(i64 t, i64) const {
    const i64 i = j + 1 + t;
    if (i >= n || info[INFO_FAIL] >= 0) return;
    const El<R> a = ld<R, CX>(A, i * rs + j * cs);
    st<R, CX>(w, i, a);
    st<R, CX>(A, i * rs + j * cs, scale(a, sc[0]));
  }
};
// trailing lower triangle: a(i, k) -= l(i, j) conj(w(k)) for i >= k > j
template <class R, bool CX>
struct TrailingUpdate {
  R* A; i64 rs, cs, j, n; const i64* info; const R* w;
  LD_HD void operator()(i64 ti, i64 tk) const {
    const i64 i = j + 1 + ti, k = j + 1 + tk;
    if (i >= n || k >= n || i < k || info[INFO_FAIL] >= 0) return;
    const El<R> l = ld<R, CX>(A, i * rs + j * cs), wk = conj(ld<R, CX>(w, k));
    const El<R> p = mul(l, wk);
    El<R> a = ld<R, CX>(A, i * rs + k * cs);
    a.re -= p.re; a.im -= p.im;
    st<R, CX>(A, i * rs + k * cs, a);
  }
};
// factor.rs:757-765: a(i, i) = D[i] for i < n (success) or i <= index (failure)
template <class R, bool CX>
struct WriteDiag {
  R* A; i64 rs, cs, n; const R* D; const i64* info;
  LD_HD void operator()(i64 i, i64) const {
    const i64 init = info[INFO_FAIL] >= 0 ? info[INFO_FAIL] + 1 : n;
    if (i >= init || i >= n) return;
    El<R> e; e.re = D[i]; e.im = R(0);
    st<R, CX>(A, i * rs + i * cs, e);
  }
};

// the whole factorization of the lower triangle of A (n x n view, any strides). info: {-1, 0} on entry. w: n elements.
template <class R, bool CX, class L>
void factor_unblocked(L& run, R* A, i64 rs, i64 cs, i64 n, R* D, const signed char* signs, R delta, R eps, i64* info, R* sc, R* w) {
  const int regularize = delta > R(0) && eps > R(0);  // factor.rs:744-745
  for (i64 j = 0; j < n; ++j) {
    run(Pivot<R, CX>{A, rs, cs, j, D, signs, delta, eps, regularize, info, sc}, 1, 1);
    if (j + 1 < n) {
      run(ScaleCol<R, CX>{A, rs, cs, j, n, info, sc, w}, n - j - 1, 1);
      run(TrailingUpdate<R, CX>{A, rs, cs, j, n, info, w}, n - j - 1, n - j - 1);
    }
  }
  run(WriteDiag<R, CX>{A, rs, cs, n, D, info}, n, 1);
}

// ---- bodies of the solve / reconstruct / inverse compositions --------------------------------------------------------------------
// dinv(i) = 1 / Re D(i), D a strided vector of T-typed elements (ldlt/solve.rs:33-41: the real part only)
template <class R, bool CX>
struct RecipDiag {
  const R* Dv; i64 stride, n; R* dinv;
  LD_HD void operator()(i64 i, i64) const {
    if (i < n) dinv[i] = R(1) / ld<R, CX>(Dv, i * stride).re;
  }
};
// rhs(i, c) *= dinv(i)
template <class R, bool CX>
struct ScaleRows {
  R* X; i64 rs, cs, n, k; const R* dinv;
  LD_HD void operator()(i64 i, i64 c) const {
    if (i >= n || c >= k) return;
    st<R, CX>(X, i * rs + c * cs, scale(ld<R, CX>(X, i * rs + c * cs), dinv[i]));
  }
};
// LxD (compact n x n): (j, j) = d(j); (i, j) = l(i, j) d(j) below; the strict upper part is zero-filled (never read: TRI_LOWER)
template <class R, bool CX>
struct BuildLxD {
  const R* Lm; i64 rs, cs; const R* Dv; i64 dstride; R* out; i64 ld_, n;
  LD_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n) return;
    const R d = ld<R, CX>(Dv, j * dstride).re;
    El<R> e; e.re = R(0); e.im = R(0);
    if (i == j) e.re = d;
    else if (i > j) e = scale(ld<R, CX>(Lm, i * rs + j * cs), d);
    st<R, CX>(out, i + j * ld_, e);
  }
};
// compact n x n identity
template <class R, bool CX>
struct SetIdentity {
  R* X; i64 ld_, n;
  LD_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n) return;
    El<R> e; e.re = i == j ? R(1) : R(0); e.im = R(0);
    st<R, CX>(X, i + j * ld_, e);
  }
};
// inverse.rs:33-48 on M = L^-1 (compact, unit lower): (j, j) = 1 / d(j); (j, i) = conj(M(i, j)) / d(i) above the diagonal; the strict
// lower part (M itself) stays. Reads only the strict lower part, writes only the diagonal and the strict upper part.
template <class R, bool CX>
struct FillUpperAdjoint {
  R* M; i64 ld_, n; const R* dinv;
  LD_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n || i < j) return;
    if (i == j) { El<R> e; e.re = dinv[j]; e.im = R(0); st<R, CX>(M, j + j * ld_, e); return; }
    st<R, CX>(M, j + i * ld_, scale(conj(ld<R, CX>(M, i + j * ld_)), dinv[i]));
  }
};

}  // namespace ldl
