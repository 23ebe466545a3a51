// Complex (c64) reductions to condensed form, unblocked: the per-thread bodies and the launch sequences, shared by the CUDA
// build (cplx_condensed.cu) and a host build that runs the same sequences thread by thread (tools/emul/cplx_condensed_host.cpp,
// tests/test_cplx_condensed_emul_cpu.py). Every launch is a flat map: body(i, j) for i < nx, j < ny, with no communication
// between the threads of one launch (reductions are split into a partial-sum launch and a single-thread finishing launch), so
// the host build is an exact model of the device control flow; it also replays every launch in reverse thread order to show
// that no body reads what another body of the same launch writes.
//
// Algorithm (tests/cplx_condensed_model.py is the numpy statement of the same steps; numbers refer to it):
//   tridiagonalization of a self-adjoint matrix  evd/tridiag.rs:274-529   the reference keeps one rank-2 update pending and
//   bidiagonalization                            svd/bidiag.rs:47-256     fuses it into the next column's pass; here each
//   column applies its two-sided update at once. Same reflectors H_k = I - v_k v_k^H / tau_k (householder.rs:59-107: tau is
//   real, so H_k is Hermitian and unitary), same condensed entries up to rounding (test_cplx_condensed_model_cpu.py).
//   Functional, not tuned: one thread per row / element, a full Hermitian working copy so that every access is coalesced.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CC_HD __host__ __device__ __forceinline__
#else
#define CC_HD inline
#endif

namespace cc {

typedef long long i64;

struct Cx {
  double re, im;
};
CC_HD Cx cx(double re, double im) { Cx r; r.re = re; r.im = im; return r; }
CC_HD Cx cconj(Cx a) { return cx(a.re, -a.im); }
CC_HD Cx cadd(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
CC_HD Cx csub(Cx a, Cx b) { return cx(a.re - b.re, a.im - b.im); }
CC_HD Cx cmul(Cx a, Cx b) { return cx(fma(a.re, b.re, -a.im * b.im), fma(a.re, b.im, a.im * b.re)); }
CC_HD Cx cscale(Cx a, double s) { return cx(a.re * s, a.im * s); }
CC_HD double cabs(Cx a) { return hypot(a.re, a.im); }
// 1 / d without spurious overflow (faer-traits/src/lib.rs:337-401, the scaled reciprocal)
CC_HD Cx crecip(Cx d) {
  const double mx = fmax(fabs(d.re), fabs(d.im));
  if (mx == 0.0) return cx(INFINITY, 0.0);
  const double sc = 1.0 / mx;
  const double a = d.re * sc, b = d.im * sc;
  const double s = sc / fma(a, a, b * b);
  return cx(a * s, -b * s);
}
// z / |z|, 1 for z = 0 (svd/mod.rs:214-222 `normalized`: scaled first so that |z|^2 cannot overflow)
CC_HD Cx cphase(Cx z) {
  const double mx = fmax(fabs(z.re), fabs(z.im));
  if (mx == 0.0 || !(mx == mx)) return cx(1.0, 0.0);
  const Cx y = cscale(z, 1.0 / mx);
  return cscale(y, 1.0 / cabs(y));
}

constexpr double MIN_POS = 2.2250738585072014e-308;
constexpr double SML = 0x1p-511, BIG = 0x1p511;  // norm_l2.rs:6-172: the three scaled accumulators
constexpr int NP = 1024;                           // partial sums per reduction

// scalar block shared by the launches of one column step
enum Sc : int { SC_INV_RE = 0, SC_INV_IM, SC_TAU_INV, SC_SKIP, SC_K_RE, SC_K_IM, SC_COUNT };

// ---- the per-thread bodies -----------------------------------------------------------------------------------------------------
// W(i, j) = the self-adjoint matrix whose lower triangle is in A (any strides, complex units); diagonal: real part only
struct BuildHermitian {
  const double* A; i64 rs, cs; Cx* W; i64 ld, n;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n) return;
    Cx v;
    if (i > j) { const double* p = A + 2 * (i * rs + j * cs); v = cx(p[0], p[1]); }
    else if (i < j) { const double* p = A + 2 * (j * rs + i * cs); v = cx(p[0], -p[1]); }
    else { const double* p = A + 2 * (i * rs + i * cs); v = cx(p[0], 0.0); }
    W[i + j * ld] = v;
  }
};
// W(i, j) = A(i, j) or conj(A(j, i)) (adjoint = true: W is A^H), m x n
struct CopyIn {
  const double* A; i64 rs, cs; Cx* W; i64 ld, m, n; int adjoint;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= m || j >= n) return;
    if (adjoint) { const double* p = A + 2 * (j * rs + i * cs); W[i + j * ld] = cx(p[0], -p[1]); }
    else { const double* p = A + 2 * (i * rs + j * cs); W[i + j * ld] = cx(p[0], p[1]); }
  }
};
// partial sums of |x_i|^2 (three scalings) over the TAIL x[1..len) of the strided vector x; thread t takes i = 1 + t, 1 + t + NP, ...
struct NormPartial {
  const Cx* x; i64 stride, len; double* part;  // part[3 * NP]
  CC_HD void operator()(i64 t, i64) const {
    if (t >= NP) return;
    double a0 = 0, a1 = 0, a2 = 0;
    for (i64 i = 1 + t; i < len; i += NP) {
      const Cx v = x[i * stride];
      const double xs = v.re * SML, ys = v.im * SML, xb = v.re * BIG, yb = v.im * BIG;
      a0 += fma(xs, xs, ys * ys);
      a1 += fma(v.re, v.re, v.im * v.im);
      a2 += fma(xb, xb, yb * yb);
    }
    part[3 * t] = a0; part[3 * t + 1] = a1; part[3 * t + 2] = a2;
  }
};
// make_householder_imp (householder.rs:59-107) on x (head x[0]); one thread. Writes beta into x[0], tau, and the scalar block
struct HouseFinal {
  Cx* x; const double* part; double* tau_out; double* sc;
  CC_HD void operator()(i64 t, i64) const {
    if (t != 0) return;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int q = 0; q < NP; ++q) { a0 += part[3 * q]; a1 += part[3 * q + 1]; a2 += part[3 * q + 2]; }
    double tail_norm;
    if (a0 >= 1.0) tail_norm = sqrt(a0) * BIG;
    else if (a1 >= 1.0) tail_norm = sqrt(a1);
    else tail_norm = sqrt(a2) * SML;
    Cx head = x[0];
    double head_norm = cabs(head);
    if (head_norm < MIN_POS) { head = cx(0.0, 0.0); head_norm = 0.0; x[0] = head; }
    if (tail_norm < MIN_POS) {
      tau_out[0] = INFINITY;
      sc[SC_INV_RE] = 1.0; sc[SC_INV_IM] = 0.0; sc[SC_TAU_INV] = 0.0; sc[SC_SKIP] = 1.0;
      return;
    }
    const double norm = hypot(head_norm, tail_norm);
    Cx sign = cx(1.0, 0.0);
    if (head_norm != 0.0) sign = cscale(head, 1.0 / head_norm);
    const Cx signed_norm = cscale(sign, norm);
    const Cx inv = crecip(cadd(head, signed_norm));
    x[0] = cx(-signed_norm.re, -signed_norm.im);
    const double tn = tail_norm * cabs(inv);
    const double tau = 0.5 * (1.0 + tn * tn);
    tau_out[0] = tau;
    sc[SC_INV_RE] = inv.re; sc[SC_INV_IM] = inv.im; sc[SC_TAU_INV] = 1.0 / tau; sc[SC_SKIP] = 0.0;
  }
};
// tail of x scaled by inv (the essential part), and the dense copies of [1, essential]: v (conj_v: conjugated) and, when given,
// v2 (never conjugated)
struct ScaleTail {
  Cx* x; i64 stride, len; const double* sc; Cx* v; int conj_v; Cx* v2;
  CC_HD void operator()(i64 i, i64) const {
    if (i >= len) return;
    Cx e = cx(1.0, 0.0);
    if (i > 0) {
      e = x[i * stride];
      if (sc[SC_SKIP] == 0.0) { e = cmul(e, cx(sc[SC_INV_RE], sc[SC_INV_IM])); x[i * stride] = e; }
    }
    v[i] = conj_v ? cconj(e) : e;
    if (v2) v2[i] = e;
  }
};
// p(i) = tau_inv * sum_j M(i, j) v(j), M: rows x cols at (r0, c0) of W (thread per row: coalesced down the columns)
struct MatVec {
  const Cx* W; i64 ld, r0, c0, rows, cols; const Cx* v; const double* sc; Cx* p;
  CC_HD void operator()(i64 i, i64) const {
    if (i >= rows || sc[SC_SKIP] != 0.0) return;
    double ar = 0, ai = 0;
    const Cx* row = W + (r0 + i) + c0 * ld;
    for (i64 j = 0; j < cols; ++j) {
      const Cx a = row[j * ld], b = v[j];
      ar += fma(a.re, b.re, -a.im * b.im);
      ai += fma(a.re, b.im, a.im * b.re);
    }
    p[i] = cscale(cx(ar, ai), sc[SC_TAU_INV]);
  }
};
// y(j) = tau_inv * sum_i conj(v(i)) M(i, j)  (thread per column)
struct VecHMat {
  const Cx* W; i64 ld, r0, c0, rows, cols; const Cx* v; const double* sc; Cx* y;
  CC_HD void operator()(i64 j, i64) const {
    if (j >= cols || sc[SC_SKIP] != 0.0) return;
    double ar = 0, ai = 0;
    const Cx* col = W + r0 + (c0 + j) * ld;
    for (i64 i = 0; i < rows; ++i) {
      const Cx a = col[i], b = v[i];
      ar += fma(b.re, a.re, b.im * a.im);   // conj(b) * a
      ai += fma(b.re, a.im, -b.im * a.re);
    }
    y[j] = cscale(cx(ar, ai), sc[SC_TAU_INV]);
  }
};
// partial sums of conj(v(i)) p(i)
struct DotPartial {
  const Cx* v; const Cx* p; i64 len; const double* sc; double* part;  // part[2 * NP]
  CC_HD void operator()(i64 t, i64) const {
    if (t >= NP) return;
    double ar = 0, ai = 0;
    if (sc[SC_SKIP] == 0.0)
      for (i64 i = t; i < len; i += NP) {
        const Cx a = v[i], b = p[i];
        ar += fma(a.re, b.re, a.im * b.im);
        ai += fma(a.re, b.im, -a.im * b.re);
      }
    part[2 * t] = ar; part[2 * t + 1] = ai;
  }
};
// K = (v^H p) * tau_inv / 2; one thread
struct DotFinal {
  const double* part; double* sc;
  CC_HD void operator()(i64 t, i64) const {
    if (t != 0) return;
    double ar = 0, ai = 0;
    for (int q = 0; q < NP; ++q) { ar += part[2 * q]; ai += part[2 * q + 1]; }
    const double h = 0.5 * sc[SC_TAU_INV];
    sc[SC_K_RE] = ar * h; sc[SC_K_IM] = ai * h;
  }
};
// w = p - K v
struct WVec {
  const Cx* v; const Cx* p; i64 len; const double* sc; Cx* w;
  CC_HD void operator()(i64 i, i64) const {
    if (i >= len || sc[SC_SKIP] != 0.0) return;
    w[i] = csub(p[i], cmul(cx(sc[SC_K_RE], sc[SC_K_IM]), v[i]));
  }
};
// M(i, j) -= v(i) conj(w(j)) + w(i) conj(v(j)), M: len x len at (r0, r0)
struct Rank2 {
  Cx* W; i64 ld, r0, len; const Cx* v; const Cx* w; const double* sc;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= len || j >= len || sc[SC_SKIP] != 0.0) return;
    Cx* a = W + (r0 + i) + (r0 + j) * ld;
    Cx t = cadd(cmul(v[i], cconj(w[j])), cmul(w[i], cconj(v[j])));
    if (i == j) t.im = 0.0;  // v conj(w) + w conj(v) is real: the diagonal of a self-adjoint matrix stays exactly real
    *a = csub(*a, t);
  }
};
// M(i, j) -= a(i) b(j), M: rows x cols at (r0, c0)
struct Rank1 {
  Cx* W; i64 ld, r0, c0, rows, cols; const Cx* a; const Cx* b; const double* sc;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= rows || j >= cols || sc[SC_SKIP] != 0.0) return;
    Cx* e = W + (r0 + i) + (c0 + j) * ld;
    *e = csub(*e, cmul(a[i], b[j]));
  }
};
// tridiagonal -> real: d(k) = Re W(k, k), e(k) = |W(k+1, k)|, ph(0) = 1, ph(k+1) = ph(k) phase(W(k+1, k)); tauc(k) = (tau(k), 0).
// One thread (a running product).
struct TridiagPhases {
  const Cx* W; i64 ld, n; const double* tau; double* d; double* e; Cx* ph; Cx* tauc;
  CC_HD void operator()(i64 t, i64) const {
    if (t != 0) return;
    Cx run = cx(1.0, 0.0);
    for (i64 k = 0; k < n; ++k) {
      d[k] = W[k + k * ld].re;
      ph[k] = run;
      if (k + 1 < n) {
        const Cx s = W[(k + 1) + k * ld];
        e[k] = cabs(s);
        run = cmul(run, cphase(s));
        run = cscale(run, 1.0 / cabs(run));  // keep the running product on the unit circle
        tauc[k] = cx(tau[k], 0.0);
      }
    }
  }
};
// bidiagonal -> real (B = Dl B_real Dr^H): r(0) = 1, l(k) = phase(d(k) r(k)), r(k+1) = conj(phase(conj(l(k)) f(k))); one thread
struct BidiagPhases {
  const Cx* W; i64 ld, n; const double* tl; const double* tr; double* d; double* f; Cx* l; Cx* r; Cx* tlc; Cx* trc;
  CC_HD void operator()(i64 t, i64) const {
    if (t != 0) return;
    Cx rk = cx(1.0, 0.0);
    for (i64 k = 0; k < n; ++k) {
      const Cx dk = W[k + k * ld];
      r[k] = rk;
      const Cx lk = cphase(cmul(dk, rk));
      l[k] = lk;
      d[k] = cabs(dk);
      tlc[k] = cx(tl[k], 0.0);
      if (k + 1 < n) {
        const Cx fk = W[k + (k + 1) * ld];
        f[k] = cabs(fk);
        rk = cconj(cphase(cmul(cconj(lk), fk)));
        trc[k] = cx(tr[k], 0.0);
      }
    }
  }
};
// out(i, j) = ph(i) * Q(i, j) for i < nq, j < nq (Q real, column-major); identity on the rest of the rows x cols output
struct ScaleRowsEmbed {
  const double* Q; i64 ldq, nq; const Cx* ph; Cx* out; i64 ldo, rows, cols;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= rows || j >= cols) return;
    Cx v;
    if (i < nq && j < nq) v = cscale(ph[i], Q[i + j * ldq]);
    else v = cx(i == j ? 1.0 : 0.0, 0.0);
    out[i + j * ldo] = v;
  }
};
// T(i, j) = W(j, i) for an n x n corner (the right reflectors as columns: svd/mod.rs:415-419)
struct TransposeCorner {
  const Cx* W; i64 ld; Cx* T; i64 ldt, n;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n) return;
    T[i + j * ldt] = W[j + i * ld];
  }
};
// out view (any strides, complex units, element type TO = double or float) <- src (rows x cols, column-major)
template <class TO>
struct CopyOut {
  TO* out; i64 rs, cs; const Cx* src; i64 lds, rows, cols;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= rows || j >= cols) return;
    const Cx v = src[i + j * lds];
    TO* p = out + 2 * (i * rs + j * cs);
    p[0] = (TO)v.re; p[1] = (TO)v.im;
  }
};
// the LOWER triangle of the out view <- the lower triangle of src (n x n); nothing above the diagonal is written
template <class TO>
struct CopyOutLower {
  TO* out; i64 rs, cs; const Cx* src; i64 lds, n;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= n || j >= n || i < j) return;
    const Cx v = src[i + j * lds];
    TO* p = out + 2 * (i * rs + j * cs);
    p[0] = (TO)v.re; p[1] = (TO)v.im;
  }
};
// S(k) = (vals(k), 0) with stride (complex units)
template <class TO>
struct CopyValues {
  TO* S; i64 stride; const double* vals; i64 n;
  CC_HD void operator()(i64 k, i64) const {
    if (k >= n) return;
    S[2 * k * stride] = (TO)vals[k];
    S[2 * k * stride + 1] = (TO)0;
  }
};
// widening copy of a c32 view to a compact c64 matrix
struct WidenC32 {
  const float* A; i64 rs, cs; Cx* W; i64 ld, m, n;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i >= m || j >= n) return;
    const float* p = A + 2 * (i * rs + j * cs);
    W[i + j * ld] = cx((double)p[0], (double)p[1]);
  }
};

// T blocks of the reflectors stored in the columns of V (m x s, unit-lower trapezoid, leading dimension ld): for block j0 = (c / bs) bs
// the b x b upper-triangular T with diag = tau and T(a, c) = v_a^H v_c above it (upgrade_householder_factor, householder.rs:132-272),
// written to Tf(a - j0, c) of the bs x s factor (column-major, leading dimension ldt). One thread per (a, c) with a <= c in one block.
struct BuildTBlocks {
  const Cx* V; i64 ld, m, s; const double* tau; Cx* Tf; i64 ldt, bs;
  CC_HD void operator()(i64 r, i64 c) const {
    if (c >= s || r >= bs) return;
    const i64 j0 = (c / bs) * bs, a = j0 + r;
    if (a > c) return;
    if (a == c) { Tf[r + c * ldt] = cx(tau[c], 0.0); return; }
    // v_a = [0.., 1 (row a), V(a+1.., a)], v_c likewise; the product runs over rows >= c
    const Cx vac = V[c + a * ld];  // row c of v_a (c > a: an essential), times the implicit 1 of v_c
    double ar = vac.re, ai = -vac.im;
    for (i64 i = c + 1; i < m; ++i) {
      const Cx x = V[i + a * ld], y = V[i + c * ld];
      ar += fma(x.re, y.re, x.im * y.im);  // conj(x) * y
      ai += fma(x.re, y.im, -x.im * y.re);
    }
    Tf[r + c * ldt] = cx(ar, ai);
  }
};
// real view (any strides) -> compact complex, and back (the real dtypes run the complex sequences on (x, 0): every product with a
// zero imaginary part stays exactly real)
template <class TR>
struct RealToCx {
  const TR* A; i64 rs, cs; Cx* W; i64 ld, m, n;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i < m && j < n) W[i + j * ld] = cx((double)A[i * rs + j * cs], 0.0);
  }
};
template <class TR>
struct CxToReal {
  TR* A; i64 rs, cs; const Cx* W; i64 ld, m, n;
  CC_HD void operator()(i64 i, i64 j) const {
    if (i < m && j < n) A[i * rs + j * cs] = (TR)W[i + j * ld].re;
  }
};

// workspace of the reductions (device pointers on the GPU, plain arrays in the host build)
struct Work {
  Cx *v, *p, *w;       // max(m, n) entries each
  double* part;        // 3 * NP
  double* sc;          // SC_COUNT
};

// ---- launch sequences (L: the launcher; L.run(body, nx, ny)) --------------------------------------------------------------------
// W: n x n full Hermitian (ld), tau: n - 1 entries. On return T is on W's diagonal / subdiagonal, reflector k below the
// subdiagonal of column k (tests/cplx_condensed_model.py: tridiag_unblocked).
template <class L>
void tridiag_unblocked(L& run, Cx* W, i64 ld, i64 n, double* tau, const Work& ws) {
  for (i64 k = 0; k + 1 < n; ++k) {
    const i64 len = n - k - 1;
    Cx* x = W + (k + 1) + k * ld;
    run(NormPartial{x, 1, len, ws.part}, NP, 1);                                     // (1)
    run(HouseFinal{x, ws.part, tau + k, ws.sc}, 1, 1);
    run(ScaleTail{x, 1, len, ws.sc, ws.v, 0, nullptr}, len, 1);
    run(MatVec{W, ld, k + 1, k + 1, len, len, ws.v, ws.sc, ws.p}, len, 1);           // (2)
    run(DotPartial{ws.v, ws.p, len, ws.sc, ws.part}, NP, 1);                         // (3)
    run(DotFinal{ws.part, ws.sc}, 1, 1);
    run(WVec{ws.v, ws.p, len, ws.sc, ws.w}, len, 1);
    run(Rank2{W, ld, k + 1, len, ws.v, ws.w, ws.sc}, len, len);                      // (4)
  }
}

// W: m x n (m >= n, ld), tl: n entries, tr: n - 1 entries (tests/cplx_condensed_model.py: bidiag_unblocked)
template <class L>
void bidiag_unblocked(L& run, Cx* W, i64 ld, i64 m, i64 n, double* tl, double* tr, const Work& ws) {
  for (i64 k = 0; k < n; ++k) {
    const i64 rows = m - k, cols = n - k - 1;
    Cx* x = W + k + k * ld;
    run(NormPartial{x, 1, rows, ws.part}, NP, 1);                                    // (1)
    run(HouseFinal{x, ws.part, tl + k, ws.sc}, 1, 1);
    run(ScaleTail{x, 1, rows, ws.sc, ws.v, 0, nullptr}, rows, 1);
    if (cols == 0) break;
    run(VecHMat{W, ld, k, k + 1, rows, cols, ws.v, ws.sc, ws.p}, cols, 1);           // (2) y = v^H M / tau
    run(Rank1{W, ld, k, k + 1, rows, cols, ws.v, ws.p, ws.sc}, rows, cols);          // (3) M -= v y
    Cx* r = W + k + (k + 1) * ld;
    run(NormPartial{r, ld, cols, ws.part}, NP, 1);                                   // (4)
    run(HouseFinal{r, ws.part, tr + k, ws.sc}, 1, 1);
    run(ScaleTail{r, ld, cols, ws.sc, ws.v, 1, ws.w}, cols, 1);                      //     v = conj([1, essential]), w = [1, essential]
    run(MatVec{W, ld, k + 1, k + 1, rows - 1, cols, ws.v, ws.sc, ws.p}, rows - 1, 1);  // (5) z = M conj(v) / tau
    run(Rank1{W, ld, k + 1, k + 1, rows - 1, cols, ws.p, ws.w, ws.sc}, rows - 1, cols);  // (6) M -= z v^T
  }
}

// W: n x n general matrix (ld), tau: n - 1 entries. On return the upper Hessenberg form H = Q^H A Q is in the entries (i, j) with
// i <= j + 1, reflector k (Q = H_0 H_1 ... H_{n-2}) below the subdiagonal of column k (evd/hessenberg.rs:549-567; the reference's
// unblocked variant defers and fuses the updates, here every column applies its similarity transform at once).
template <class L>
void hessenberg_unblocked(L& run, Cx* W, i64 ld, i64 n, double* tau, const Work& ws) {
  for (i64 k = 0; k + 1 < n; ++k) {
    const i64 len = n - k - 1;
    Cx* x = W + (k + 1) + k * ld;
    run(NormPartial{x, 1, len, ws.part}, NP, 1);
    run(HouseFinal{x, ws.part, tau + k, ws.sc}, 1, 1);
    run(ScaleTail{x, 1, len, ws.sc, ws.w, 1, ws.v}, len, 1);                          // v = [1, essential], w = conj(v)
    run(VecHMat{W, ld, k + 1, k + 1, len, len, ws.v, ws.sc, ws.p}, len, 1);           // y = v^H M / tau, M = W[k+1.., k+1..]
    run(Rank1{W, ld, k + 1, k + 1, len, len, ws.v, ws.p, ws.sc}, len, len);           // M -= v y            (H M)
    run(MatVec{W, ld, 0, k + 1, n, len, ws.v, ws.sc, ws.p}, n, 1);                    // z = N v / tau, N = W[.., k+1..]
    run(Rank1{W, ld, 0, k + 1, n, len, ws.p, ws.w, ws.sc}, n, len);                   // N -= z v^H          (N H)
  }
}

}  // namespace cc
