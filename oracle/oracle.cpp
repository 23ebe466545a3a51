// TEST INFRASTRUCTURE — NOT PRODUCT CODE. See oracle.hpp for the contract and the parity-pinning statement.
#include "oracle.hpp"

#include <omp.h>

#include <algorithm>
#include <cstring>
#include <vector>

namespace oracle {

// ------------------------------------------------------------------------------------------------
// matmul — schoolbook definition: acc = sum_k conj_a(a_ik) * conj_b(b_kj) accumulated in k order, then
// dst = acc*alpha (Replace) or dst + acc*alpha (Add).
// Reference: faer/src/linalg/matmul/mod.rs:1909-1947 (`matmul_with_conj_fallback`, the comparator the
// reference's own test_matmul uses), 1505-1523 (generic path), 1190-1198 (M==0/N==0/K==0 edge cases),
// 1580-1582 (Replace must not read dst).
// The (i,j) loop is register-blocked for speed; each output element still sums in k order.
// ------------------------------------------------------------------------------------------------
// acc[b][a] = sum_k conj?(lhs(i0 + a, k)) * conj?(rhs(k, j0 + b)), k ascending (the scalar definition, matmul/mod.rs:1909-1947).
// Full tiles of a unit-row-stride lhs take the fixed-trip-count loop, which the compiler vectorises along `a`; the
// per-element operation order is the same on every path, so the results are bitwise independent of the path taken.
constexpr int OR_BI = 8, OR_BJ = 4;
template <class T>
static inline void tile_product(T (&acc)[OR_BJ][OR_BI], const Mat<const T>& lhs, bool conj_lhs, const Mat<const T>& rhs,
                                bool conj_rhs, i64 i0, i64 ni, i64 j0, i64 nj, i64 K) {
  for (int b = 0; b < OR_BJ; ++b)
    for (int a = 0; a < OR_BI; ++a) acc[b][a] = T(0);
  if (ni == OR_BI && nj == OR_BJ && lhs.rs == 1) {
    for (i64 k = 0; k < K; ++k) {
      const T* lp = &lhs(i0, k);
      T bv[OR_BJ];
      for (int b = 0; b < OR_BJ; ++b) bv[b] = conj_if(conj_rhs, rhs(k, j0 + b));
      for (int b = 0; b < OR_BJ; ++b)
        for (int a = 0; a < OR_BI; ++a) acc[b][a] = acc[b][a] + conj_if(conj_lhs, lp[a]) * bv[b];
    }
    return;
  }
  for (i64 k = 0; k < K; ++k) {
    T bv[OR_BJ];
    for (i64 b = 0; b < nj; ++b) bv[b] = conj_if(conj_rhs, rhs(k, j0 + b));
    for (i64 a = 0; a < ni; ++a) {
      const T av = conj_if(conj_lhs, lhs(i0 + a, k));
      for (i64 b = 0; b < nj; ++b) acc[b][a] = acc[b][a] + av * bv[b];
    }
  }
}

// ---- cache-blocked form of the same definition -----------------------------------------------------------------
// Large products run on packed panels (the classical MC x KC / KC x NC blocking): a task owns an MC x NC block of dst,
// packs the operands KC columns at a time into contiguous 8-row / 4-column micro-panels (conjugation applied while
// packing, zero padding at the edges) and keeps the raw accumulators of its block in a local buffer between K blocks.
// Every output element is still  acc = 0; acc = acc + a_ik * b_kj for k ascending; v = acc * alpha; dst (+)= v  with
// unfused multiply and add, so the results are bitwise those of the loops above and below; only the memory traffic
// changes (the unpacked loop touches one cache line of lhs per k). Tasks are independent: OpenMP over the 2-D grid.
constexpr i64 OR_MC = 128, OR_NC = 64, OR_KC = 256;

static inline bool s_lower(int s);
static inline bool s_upper(int s);
static inline bool s_nodiag(int s);

template <class T>
static inline void micro_kernel(T* __restrict c, const T* __restrict ap, const T* __restrict bp, i64 kc) {
  T acc[OR_BJ][OR_BI];
  for (int b = 0; b < OR_BJ; ++b)
    for (int a = 0; a < OR_BI; ++a) acc[b][a] = c[b * OR_BI + a];
  for (i64 k = 0; k < kc; ++k) {
    const T* __restrict av = ap + k * OR_BI;
    const T* __restrict bv = bp + k * OR_BJ;
    for (int b = 0; b < OR_BJ; ++b)
      for (int a = 0; a < OR_BI; ++a) acc[b][a] = acc[b][a] + av[a] * bv[b];
  }
  for (int b = 0; b < OR_BJ; ++b)
    for (int a = 0; a < OR_BI; ++a) c[b * OR_BI + a] = acc[b][a];
}

// dst_s selects the written part of dst (RECT: all of it); lhs and rhs are unstructured.
template <class T>
static void gemm_blocked(Mat<T> dst, int dst_s, bool add, Mat<const T> lhs, bool conj_lhs, Mat<const T> rhs, bool conj_rhs,
                         T alpha) {
  const i64 M = dst.m, N = dst.n, K = lhs.n;
  const i64 nic = (M + OR_MC - 1) / OR_MC, njc = (N + OR_NC - 1) / OR_NC;
  constexpr i64 TI = OR_MC / OR_BI, TJ = OR_NC / OR_BJ;
  // no more threads than blocks (hosts with hundreds of hardware threads), buffers only for threads that get a block
  const int nthr = (int)std::max<i64>(1, std::min<i64>(omp_get_max_threads(), nic * njc));
#pragma omp parallel num_threads(nthr)
  {
    std::vector<T> Ap, Bp, Cb;
#pragma omp for schedule(dynamic, 1) collapse(2)
    for (i64 jc = 0; jc < njc; ++jc) {
      for (i64 ic = 0; ic < nic; ++ic) {
        const i64 i0 = ic * OR_MC, mc = std::min<i64>(OR_MC, M - i0);
        const i64 j0 = jc * OR_NC, nc = std::min<i64>(OR_NC, N - j0);
        if (dst_s != RECT) {
          if (s_lower(dst_s) && i0 + mc - 1 < j0) continue;  // block strictly above the diagonal
          if (s_upper(dst_s) && i0 > j0 + nc - 1) continue;  // block strictly below the diagonal
        }
        if (Ap.empty()) {
          Ap.resize((size_t)OR_MC * OR_KC);
          Bp.resize((size_t)OR_KC * OR_NC);
          Cb.resize((size_t)OR_MC * OR_NC);
        }
        const i64 ti = (mc + OR_BI - 1) / OR_BI, tj = (nc + OR_BJ - 1) / OR_BJ;
        std::fill(Cb.begin(), Cb.begin() + (size_t)(ti * tj) * OR_BI * OR_BJ, T(0));
        for (i64 p0 = 0; p0 < K; p0 += OR_KC) {
          const i64 kc = std::min<i64>(OR_KC, K - p0);
          // Bp[t][k][4], Ap[t][k][8]
          for (i64 t = 0; t < tj; ++t) {
            T* bp = Bp.data() + (size_t)t * kc * OR_BJ;
            const i64 jb = j0 + t * OR_BJ, nj = std::min<i64>(OR_BJ, N - jb);
            for (i64 k = 0; k < kc; ++k)
              for (i64 b = 0; b < OR_BJ; ++b) bp[k * OR_BJ + b] = b < nj ? conj_if(conj_rhs, rhs(p0 + k, jb + b)) : T(0);
          }
          for (i64 t = 0; t < ti; ++t) {
            T* ap = Ap.data() + (size_t)t * kc * OR_BI;
            const i64 ib = i0 + t * OR_BI, ni = std::min<i64>(OR_BI, M - ib);
            if (ni == OR_BI && lhs.rs == 1 && !conj_lhs) {
              for (i64 k = 0; k < kc; ++k) std::memcpy(ap + k * OR_BI, &lhs(ib, p0 + k), sizeof(T) * OR_BI);
            } else {
              for (i64 k = 0; k < kc; ++k)
                for (i64 a = 0; a < OR_BI; ++a) ap[k * OR_BI + a] = a < ni ? conj_if(conj_lhs, lhs(ib + a, p0 + k)) : T(0);
            }
          }
          for (i64 t2 = 0; t2 < tj; ++t2) {
            const i64 jb = j0 + t2 * OR_BJ, nj = std::min<i64>(OR_BJ, N - jb);
            for (i64 t1 = 0; t1 < ti; ++t1) {
              const i64 ib = i0 + t1 * OR_BI, ni = std::min<i64>(OR_BI, M - ib);
              if (dst_s != RECT) {
                if (s_lower(dst_s) && ib + ni - 1 < jb) continue;
                if (s_upper(dst_s) && ib > jb + nj - 1) continue;
              }
              micro_kernel<T>(Cb.data() + (size_t)(t2 * ti + t1) * OR_BI * OR_BJ, Ap.data() + (size_t)t1 * kc * OR_BI,
                              Bp.data() + (size_t)t2 * kc * OR_BJ, kc);
            }
          }
        }
        for (i64 t2 = 0; t2 < tj; ++t2) {
          const i64 jb = j0 + t2 * OR_BJ, nj = std::min<i64>(OR_BJ, N - jb);
          for (i64 t1 = 0; t1 < ti; ++t1) {
            const i64 ib = i0 + t1 * OR_BI, ni = std::min<i64>(OR_BI, M - ib);
            if (dst_s != RECT) {
              if (s_lower(dst_s) && ib + ni - 1 < jb) continue;
              if (s_upper(dst_s) && ib > jb + nj - 1) continue;
            }
            const T* c = Cb.data() + (size_t)(t2 * ti + t1) * OR_BI * OR_BJ;
            for (i64 b = 0; b < nj; ++b)
              for (i64 a = 0; a < ni; ++a) {
                const i64 i = ib + a, j = jb + b;
                if (dst_s != RECT) {
                  if (i == j && s_nodiag(dst_s)) continue;
                  if (i != j && (s_lower(dst_s) ? (i < j) : (i > j))) continue;
                }
                T v = c[b * OR_BI + a] * alpha;
                if (add) v = dst(i, j) + v;
                dst(i, j) = v;
              }
          }
        }
      }
    }
  }
  (void)TI; (void)TJ;
}
// the blocked form pays off once the operands no longer sit in the first-level cache
static inline bool use_blocked(i64 M, i64 N, i64 K) { return K >= 16 && M * N >= 64 * 64 && M * N * K >= (i64)1 << 20; }

template <class T>
void matmul(Mat<T> dst, bool add, Mat<const T> lhs, bool conj_lhs, Mat<const T> rhs, bool conj_rhs, T alpha) {
  const i64 M = dst.m, N = dst.n, K = lhs.n;
  if (M == 0 || N == 0) return;
  if (K == 0) {
    if (!add)
      for (i64 j = 0; j < N; ++j)
        for (i64 i = 0; i < M; ++i) dst(i, j) = T(0);
    return;
  }
  if (use_blocked(M, N, K)) {
    gemm_blocked<T>(dst, RECT, add, lhs, conj_lhs, rhs, conj_rhs, alpha);
    return;
  }
  constexpr int BI = OR_BI, BJ = OR_BJ;
  const i64 nbj = (N + BJ - 1) / BJ;
#pragma omp parallel for schedule(dynamic, 1) if (M * N * K >= 32768)
  for (i64 jb = 0; jb < nbj; ++jb) {
    const i64 j0 = jb * BJ, nj = std::min<i64>(BJ, N - j0);
    for (i64 i0 = 0; i0 < M; i0 += BI) {
      const i64 ni = std::min<i64>(BI, M - i0);
      T acc[BJ][BI];
      tile_product<T>(acc, lhs, conj_lhs, rhs, conj_rhs, i0, ni, j0, nj, K);
      for (i64 b = 0; b < nj; ++b)
        for (i64 a = 0; a < ni; ++a) {
          T v = acc[b][a] * alpha;
          if (add) v = dst(i0 + a, j0 + b) + v;
          dst(i0 + a, j0 + b) = v;
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// triangular matmul. Reference: faer/src/linalg/matmul/triangular.rs:1193-1245, 1246-1495.
// Semantics restated (Appendix A of SURVEY.md, test `test_triangular` matmul/mod.rs:2106-2266):
//  * the half of a triangular INPUT excluded by its BlockStructure is never read as data: it is 0, and the
//    diagonal is 0 (strict) / 1 (unit)  (triangular.rs:26-52, 130-156);
//  * only the part of dst selected by its structure is written; strict/unit dst leaves the diagonal untouched.
// ------------------------------------------------------------------------------------------------
static inline bool s_lower(int s) { return s == TRI_LOWER || s == STRICT_LOWER || s == UNIT_LOWER; }
static inline bool s_upper(int s) { return s == TRI_UPPER || s == STRICT_UPPER || s == UNIT_UPPER; }
static inline bool s_nodiag(int s) { return s >= STRICT_LOWER; }

template <class T>
static inline T masked(const Mat<const T>& a, int s, i64 i, i64 j) {
  if (s == RECT) return a(i, j);
  if (i == j) {
    if (s == UNIT_LOWER || s == UNIT_UPPER) return T(1);
    if (s == STRICT_LOWER || s == STRICT_UPPER) return T(0);
    return a(i, j);
  }
  const bool keep = s_lower(s) ? (i > j) : (i < j);
  return keep ? a(i, j) : T(0);
}

template <class T>
void matmul_triangular(Mat<T> dst, int dst_s, bool add, Mat<const T> lhs, int lhs_s, bool conj_lhs, Mat<const T> rhs,
                       int rhs_s, bool conj_rhs, T alpha) {
  const i64 M = dst.m, N = dst.n, K = lhs.n;
  if (M == 0 || N == 0) return;
  if (lhs_s == RECT && rhs_s == RECT) {
    // Unstructured operands, structured destination (the SYRK-type trailing updates): same per-element arithmetic as
    // the loop below (acc over k ascending, then * alpha, then + dst), computed on 8 x 4 register tiles like matmul();
    // tiles entirely outside the selected triangle are skipped. Bitwise the same results, ~an order of magnitude faster.
    if (K > 0 && use_blocked(M, N, K)) {
      gemm_blocked<T>(dst, dst_s, add, lhs, conj_lhs, rhs, conj_rhs, alpha);
      return;
    }
    constexpr int BI = OR_BI, BJ = OR_BJ;
    const i64 nbj = (N + BJ - 1) / BJ;
#pragma omp parallel for schedule(dynamic, 1)
    for (i64 jb = 0; jb < nbj; ++jb) {
      const i64 j0 = jb * BJ, nj = std::min<i64>(BJ, N - j0);
      for (i64 i0 = 0; i0 < M; i0 += BI) {
        const i64 ni = std::min<i64>(BI, M - i0);
        if (dst_s != RECT) {
          if (s_lower(dst_s) && i0 + ni - 1 < j0) continue;  // tile strictly above the diagonal
          if (s_upper(dst_s) && i0 > j0 + nj - 1) continue;  // tile strictly below the diagonal
        }
        T acc[BJ][BI];
        tile_product<T>(acc, lhs, conj_lhs, rhs, conj_rhs, i0, ni, j0, nj, K);
        for (i64 b = 0; b < nj; ++b)
          for (i64 a = 0; a < ni; ++a) {
            const i64 i = i0 + a, j = j0 + b;
            if (dst_s != RECT) {
              if (i == j && s_nodiag(dst_s)) continue;
              if (i != j && (s_lower(dst_s) ? (i < j) : (i > j))) continue;
            }
            T v = acc[b][a] * alpha;
            if (add) v = dst(i, j) + v;
            dst(i, j) = v;
          }
      }
    }
    return;
  }
#pragma omp parallel for schedule(dynamic, 4)
  for (i64 j = 0; j < N; ++j) {
    for (i64 i = 0; i < M; ++i) {
      if (dst_s != RECT) {
        if (i == j && s_nodiag(dst_s)) continue;
        if (i != j && (s_lower(dst_s) ? (i < j) : (i > j))) continue;
      }
      T acc = T(0);
      i64 k0 = 0, k1 = K;
      if (s_lower(lhs_s)) k1 = std::min(k1, i + 1);
      if (s_upper(lhs_s)) k0 = std::max(k0, i);
      if (s_lower(rhs_s)) k0 = std::max(k0, j);
      if (s_upper(rhs_s)) k1 = std::min(k1, j + 1);
      for (i64 k = k0; k < k1; ++k)
        acc = acc + conj_if(conj_lhs, masked(lhs, lhs_s, i, k)) * conj_if(conj_rhs, masked(rhs, rhs_s, k, j));
      T v = acc * alpha;
      if (add) v = dst(i, j) + v;
      dst(i, j) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// triangular solve. Reference: faer/src/linalg/triangular_solve.rs
//   block_size 200-211, recursion_threshold 213-215 (= 4), recursive split 420-576 (top solve, GEMM update
//   with alpha = -1, bottom solve; rhs split by columns when k > 64 && n <= 128 — independent columns, so
//   the split does not change any value), leaves 16-198 (reciprocal of the diagonal, products pre-divided),
//   upper = lower on reversed views 577-604.
// ------------------------------------------------------------------------------------------------
static inline i64 ts_block_size(i64 n) {
  const i64 base_rem = n / 2;
  i64 sub;
  if (n >= 32) sub = (base_rem + 15) / 16 * 16;
  else if (n >= 16) sub = (base_rem + 7) / 8 * 8;
  else if (n >= 8) sub = (base_rem + 3) / 4 * 4;
  else sub = base_rem;
  return n - sub;
}

template <class T>
static void solve_lower_leaf(Mat<const T> L, bool conj, bool unit, Mat<T> rhs) {
  // triangular_solve.rs:16-198: y_i = y_i * (1/l_ii) + sum_{k<i} (-(l_ik) * (1/l_ii)) * y_k   (non-unit)
  //                             y_i = y_i + sum_{k<i} (-(l_ik)) * y_k                        (unit)
  const i64 n = L.m;
  T inv[4], nl[4][4];
  for (i64 i = 0; i < n; ++i) {
    inv[i] = unit ? T(1) : conj_if(conj, T(1) / L(i, i));
    for (i64 k = 0; k < i; ++k) nl[i][k] = unit ? conj_if(conj, -L(i, k)) : conj_if(conj, -L(i, k)) * inv[i];
  }
  for (i64 c = 0; c < rhs.n; ++c) {
    T y[4];
    for (i64 i = 0; i < n; ++i) {
      T v = rhs(i, c);
      if (!unit) v = v * inv[i];
      for (i64 k = 0; k < i; ++k) v = v + nl[i][k] * y[k];
      y[i] = v;
    }
    for (i64 i = 0; i < n; ++i) rhs(i, c) = y[i];
  }
}

template <class T>
static void solve_lower_rec(Mat<const T> L, bool conj, bool unit, Mat<T> rhs) {
  const i64 n = L.m;
  if (n == 0 || rhs.n == 0) return;
  if (n <= 4) {
    solve_lower_leaf(L, conj, unit, rhs);
    return;
  }
  const i64 bs = ts_block_size(n);
  Mat<const T> L00 = L.sub(0, 0, bs, bs), L10 = L.sub(bs, 0, n - bs, bs), L11 = L.sub(bs, bs, n - bs, n - bs);
  Mat<T> top = rhs.sub(0, 0, bs, rhs.n), bot = rhs.sub(bs, 0, n - bs, rhs.n);
  solve_lower_rec(L00, conj, unit, top);
  matmul<T>(bot, true, L10, conj, Mat<const T>{top.p, top.m, top.n, top.rs, top.cs}, false, T(-1));
  solve_lower_rec(L11, conj, unit, bot);
}

// The right-hand-side columns are independent and every product element is computed the same way whatever the blocking,
// so solving chunks of columns on different threads changes no value (the reference splits the same way,
// triangular_solve.rs:420-576); inside a chunk the recursion and its products run on the calling thread.
template <class T>
static void solve_lower_cols(Mat<const T> tril, bool conj, bool unit, Mat<T> rhs) {
  const i64 k = rhs.n;
  const int nt = omp_get_max_threads();
  if (k < 32 || nt <= 1 || omp_in_parallel() || tril.m < 8) {
    solve_lower_rec(tril, conj, unit, rhs);
    return;
  }
  const i64 chunk = std::max<i64>(8, (k + 4 * nt - 1) / (4 * nt));
  const i64 nchunks = (k + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1)
  for (i64 c = 0; c < nchunks; ++c) {
    const i64 c0 = c * chunk;
    solve_lower_rec(tril, conj, unit, rhs.sub(0, c0, rhs.m, std::min(chunk, k - c0)));
  }
}
template <class T>
void solve_lower(Mat<const T> tril, bool conj, bool unit, Mat<T> rhs) { solve_lower_cols(tril, conj, unit, rhs); }
template <class T>
void solve_upper(Mat<const T> triu, bool conj, bool unit, Mat<T> rhs) {
  if (triu.m == 0 || rhs.n == 0) return;
  solve_lower_cols(triu.rev_rows_cols(), conj, unit, rhs.rev_rows());
}

// ------------------------------------------------------------------------------------------------
// LLT. Reference: faer/src/linalg/cholesky/llt/factor.rs:68-97 calls
// ldlt::factor::cholesky_block_left_looking(is_llt = true) whose `if true ||` (ldlt/factor.rs:528) always
// forwards to cholesky_recursion_right_looking (367-498):
//   n <= recursion_threshold -> leaf (simd_cholesky 7-298; scalar statement 299-366)
//   else bs = min(next_pow2(n)/2, block_size); for j in steps of bs:
//        recurse on A00 with (recursion_threshold, block_size = bs)          405-419  -> Err(j + idx)
//        conj(A00) X = A10^T  (solve_lower_triangular_in_place)                421-426
//        A11(lower) += -1 * A10 * A10^H (triangular::matmul, dst lower)        435-446
// Leaf (SIMD col-major path, 7-177): a_ij <- fma(-conj(a_jk), a_ik, a_ij) for k < j in k order; d = Re(a_jj);
// regularise (122-144, sign = +1 for LLT); !(d > 0) -> Err(j); l = sqrt(d); l == 0 or non-finite -> Err(j);
// the WHOLE column j, diagonal included, is multiplied by recip(l)   (161-175).
// ------------------------------------------------------------------------------------------------
static inline i64 next_pow2(i64 n) {
  i64 p = 1;
  while (p < n) p <<= 1;
  return p;
}

template <class R> static inline R fma_(R a, R b, R c) { return std::fma(a, b, c); }
template <class R> static inline std::complex<R> fma_(std::complex<R> a, std::complex<R> b, std::complex<R> c) {
  // complex mul_add as four real FMAs (pulp's c64 mul_add lowers to fmaddsub pairs): re = fma(ar,br, fma(-ai,bi,cr))
  R re = std::fma(a.real(), b.real(), std::fma(-a.imag(), b.imag(), c.real()));
  R im = std::fma(a.real(), b.imag(), std::fma(a.imag(), b.real(), c.imag()));
  return std::complex<R>(re, im);
}

template <class T> static inline T neg_conj(T x) { return -x; }
template <class R> static inline std::complex<R> neg_conj(std::complex<R> x) { return -std::conj(x); }
template <class T> static inline T mul_real(T x, typename real_of<T>::type r) { return x * r; }
template <class R> static inline std::complex<R> mul_real(std::complex<R> x, R r) {
  return std::complex<R>(x.real() * r, x.imag() * r);
}

template <class T>
static i64 llt_leaf(Mat<T> A, typename real_of<T>::type delta, typename real_of<T>::type eps, bool regularize,
                    i64* count) {
  typedef typename real_of<T>::type R;
  const i64 n = A.m;
  for (i64 j = 0; j < n; ++j) {
    // a_ij <- fma(-conj(a_jk), a_ik, a_ij), k = 0..j-1 in order (ldlt/factor.rs:44-53)
    for (i64 i = j; i < n; ++i) {
      T a = A(i, j);
      for (i64 k = 0; k < j; ++k) a = fma_(neg_conj(A(j, k)), A(i, k), a);
      A(i, j) = a;
    }
    R diag = real_part(A(j, j));
    if (regularize) {
      // sign == +1 for LLT (ldlt/factor.rs:122-144)
      const bool small_or_negative = diag <= eps;
      if (small_or_negative) {
        diag = delta;
        *count += 1;
      }
    }
    if (!(diag > R(0))) return j;             // 147-150
    diag = std::sqrt(diag);
    if (diag == R(0) || !std::isfinite(diag)) return j;  // 156-158
    const R inv = R(1) / diag;                // recip, then multiply (161-175), diagonal included
    for (i64 i = j; i < n; ++i) A(i, j) = mul_real(A(i, j), inv);
  }
  return -1;
}

template <class T>
static i64 llt_rec(Mat<T> A, typename real_of<T>::type delta, typename real_of<T>::type eps, bool regularize,
                   i64 recursion_threshold, i64 block_size, i64* count) {
  const i64 n = A.n;
  if (n <= recursion_threshold) return llt_leaf(A, delta, eps, regularize, count);
  const i64 bs0 = std::min(next_pow2(n) / 2, block_size);
  for (i64 j = 0; j < n;) {
    const i64 bs = std::min(bs0, n - j);
    Mat<T> A00 = A.sub(j, j, bs, bs);
    const i64 fail = llt_rec(A00, delta, eps, regularize, recursion_threshold, bs, count);
    if (fail >= 0) return j + fail;
    const i64 rem = n - j - bs;
    if (rem > 0) {
      Mat<T> A10 = A.sub(j + bs, j, rem, bs);
      Mat<T> A11 = A.sub(j + bs, j + bs, rem, rem);
      Mat<const T> cA00{A00.p, A00.m, A00.n, A00.rs, A00.cs};
      Mat<const T> cA10{A10.p, A10.m, A10.n, A10.rs, A10.cs};
      solve_lower<T>(cA00, /*conj=*/true, /*unit=*/false, A10.t());
      matmul_triangular<T>(A11, TRI_LOWER, true, cA10, RECT, false, cA10.t(), RECT, true, T(-1));
    }
    j += bs;
  }
  return -1;
}

template <class T>
i64 llt_in_place(Mat<T> A, typename real_of<T>::type delta, typename real_of<T>::type eps, i64 recursion_threshold,
                 i64 block_size, i64* reg_count) {
  typedef typename real_of<T>::type R;
  *reg_count = 0;
  const bool regularize = delta > R(0) && eps > R(0);  // llt/factor.rs:85-87
  return llt_rec(A, delta, eps, regularize, recursion_threshold, block_size, reg_count);
}

// ------------------------------------------------------------------------------------------------
// LDLT. Reference: faer/src/linalg/cholesky/ldlt/factor.rs
//   cholesky_in_place 725-767: D lives in a scratch row during the factorization; afterwards A(i, i) = D[i] for
//        i < n (success) or i <= index (ZeroPivot { index }); the strict lower part holds the unit-lower L.
//   cholesky_block_left_looking 499-...: `if true ||` forwards to cholesky_recursion_right_looking(is_llt = false) 367-498:
//        recurse on A00; conj(A00) X = A10^T with the UNIT-lower solve (X = L10 D0); L10 = X * recip(D0) column by
//        column; A11(lower) -= L10 X^H. The x86 build forms the product through `spicy_matmul` (diagonal applied inside an
//        un-vendored GEMM, 447-470); the portable branch (471-492) multiplies L10 by the saved X, which is restated here
//        — X is kept in a temporary instead of the upper triangle, which the x86 path leaves untouched.
//   leaf (simd_cholesky 7-177 == cholesky_fallback 299-366 up to fusing): a_ij <- fma(conj(a_jk) * (-D_k), a_ik, a_ij),
//        k ascending; d = Re(a_jj); regularisation 122-144: sign +1 and d <= eps -> delta (counted); sign -1 and
//        d >= -eps -> -delta; no sign and |d| <= eps -> copysign-like (d < 0 ? -delta : delta); only the first case
//        increments the count; D_j = d; d == 0 or non-finite -> ZeroPivot(j); column j (diagonal included) *= recip(d).
// ------------------------------------------------------------------------------------------------
template <class T>
static i64 ldlt_leaf(Mat<T> A, typename real_of<T>::type* D, bool regularize, typename real_of<T>::type eps,
                     typename real_of<T>::type delta, const signed char* signs, i64* count) {
  typedef typename real_of<T>::type R;
  const i64 n = A.m;
  for (i64 j = 0; j < n; ++j) {
    for (i64 i = j; i < n; ++i) {
      T a = A(i, j);
      for (i64 k = 0; k < j; ++k) a = fma_(mul_real(conj_if(true, A(j, k)), -D[k]), A(i, k), a);
      A(i, j) = a;
    }
    R diag = real_part(A(j, j));
    if (regularize) {
      const int sign = signs ? (int)signs[j] : 0;
      const bool small_or_negative = diag <= eps;
      const bool minus_small_or_positive = diag >= -eps;
      if (sign == 1 && small_or_negative) {
        diag = delta;
        *count += 1;
      } else if (sign == -1 && minus_small_or_positive) {
        diag = -delta;
      } else if (small_or_negative && minus_small_or_positive) {
        diag = diag < R(0) ? -delta : delta;
      }
    }
    D[j] = diag;
    if (diag == R(0) || !std::isfinite(diag)) return j;
    const R inv = R(1) / diag;
    for (i64 i = j; i < n; ++i) A(i, j) = mul_real(A(i, j), inv);
  }
  return -1;
}

template <class T>
static i64 ldlt_rec(Mat<T> A, typename real_of<T>::type* D, bool regularize, typename real_of<T>::type eps,
                    typename real_of<T>::type delta, const signed char* signs, i64 recursion_threshold, i64 block_size,
                    i64* count) {
  typedef typename real_of<T>::type R;
  const i64 n = A.n;
  if (n <= recursion_threshold) return ldlt_leaf(A, D, regularize, eps, delta, signs, count);
  const i64 bs0 = std::min(next_pow2(n) / 2, block_size);
  for (i64 j = 0; j < n;) {
    const i64 bs = std::min(bs0, n - j);
    Mat<T> A00 = A.sub(j, j, bs, bs);
    const i64 fail = ldlt_rec(A00, D + j, regularize, eps, delta, signs ? signs + j : nullptr, recursion_threshold, bs, count);
    if (fail >= 0) return j + fail;
    const i64 rem = n - j - bs;
    if (rem > 0) {
      Mat<T> A10 = A.sub(j + bs, j, rem, bs);
      Mat<T> A11 = A.sub(j + bs, j + bs, rem, rem);
      Mat<const T> cA00{A00.p, A00.m, A00.n, A00.rs, A00.cs};
      solve_lower<T>(cA00, /*conj=*/true, /*unit=*/true, A10.t());
      std::vector<T> xbuf((size_t)rem * (size_t)bs);
      Mat<T> X{xbuf.data(), rem, bs, 1, rem};
      for (i64 k = 0; k < bs; ++k) {
        const R d = R(1) / D[j + k];
        for (i64 i = 0; i < rem; ++i) {
          const T a = A10(i, k);
          X(i, k) = a;
          A10(i, k) = mul_real(a, d);
        }
      }
      Mat<const T> cA10{A10.p, A10.m, A10.n, A10.rs, A10.cs};
      Mat<const T> cX{X.p, X.m, X.n, X.rs, X.cs};
      matmul_triangular<T>(A11, TRI_LOWER, true, cA10, RECT, false, cX.t(), RECT, true, T(-1));
    }
    j += bs;
  }
  return -1;
}

template <class T>
i64 ldlt_in_place(Mat<T> A, typename real_of<T>::type delta, typename real_of<T>::type eps, const signed char* signs,
                  i64 recursion_threshold, i64 block_size, i64* reg_count) {
  typedef typename real_of<T>::type R;
  *reg_count = 0;
  const i64 n = A.m;
  std::vector<R> D((size_t)std::max<i64>(n, 1), R(0));
  const bool regularize = delta > R(0) && eps > R(0);  // ldlt/factor.rs:744-745
  const i64 fail = ldlt_rec(A, D.data(), regularize, eps, delta, signs, recursion_threshold, block_size, reg_count);
  const i64 init = fail >= 0 ? fail + 1 : n;           // 757-765
  for (i64 i = 0; i < init; ++i) A(i, i) = T(D[(size_t)i]);
  return fail;
}

// ------------------------------------------------------------------------------------------------
// LU with partial pivoting. Reference: faer/src/linalg/lu/partial_pivoting/factor.rs
//   lu_in_place 234-295: perm <- identity; transpositions from the recursion; perm.swap(i, i + t_i) in order;
//                        m < n tail: unit-lower solve of the right block; perm_inv[perm[i]] = i.
//   lu_in_place_recursion 68-187: n <= recursion_threshold -> unblocked; bs = round_up(n/2, min(16, next_pow2(n/2)));
//        recurse left (all rows, cols [0, bs) of the [start, end) window); A01 <- unit_lower(A00)^-1 A01;
//        A11 -= A10 A01; recurse on rows bs.. for cols [bs, n); then apply ALL n transpositions of this level to
//        the columns left of `start` and right of `end` of the current view (127-185).
//   lu_in_place_unblocked 19-67: pivot = FIRST row attaining max abs1 (strict `>` from max = 0); the swap is applied
//        to the whole row of the current view (46); multipliers via recip-then-multiply (50-53); rank-1 update with
//        alpha = -1 through matmul (K = 1).
// ------------------------------------------------------------------------------------------------
template <class T>
static i64 lu_unblocked(Mat<T> A, i64 start, i64 end, i64* trans) {
  const i64 m = A.m;
  if (start == end) return 0;
  i64 n_trans = 0;
  for (i64 j = start; j < end; ++j) {
    const i64 col = j, row = j - start;
    i64 imax = row;
    typename real_of<T>::type mx = 0;
    for (i64 i = imax; i < m; ++i) {
      const auto a = abs1(A(i, col));
      if (a > mx) {
        mx = a;
        imax = i;
      }
    }
    trans[row] = imax - row;
    if (imax != row) {
      for (i64 c = 0; c < A.n; ++c) std::swap(A(row, c), A(imax, c));  // swap_rows_idx over the whole view
      n_trans += 1;
    }
    Mat<T> W = A.sub(0, start, m, end - start);
    const T inv = T(1) / W(row, row);
    for (i64 i = row + 1; i < m; ++i) W(i, row) = W(i, row) * inv;
    // A11 += -1 * A10 * A01 (rank-1)
    for (i64 c = row + 1; c < end - start; ++c) {
      const T u = W(row, c);
      for (i64 i = row + 1; i < m; ++i) W(i, c) = W(i, c) + (W(i, row) * u) * T(-1);
    }
  }
  return n_trans;
}

static inline i64 next_multiple_of(i64 x, i64 m) { return (x + m - 1) / m * m; }

template <class T>
static i64 lu_rec(Mat<T> A, i64 start, i64 end, i64* trans, i64 recursion_threshold) {
  const i64 m = A.m, ncols = A.n, n = end - start;
  if (n <= recursion_threshold) return lu_unblocked(A, start, end, trans);
  const i64 half = n / 2;
  const i64 pow = std::min<i64>(16, next_pow2(half));
  const i64 bs = next_multiple_of(half, pow);
  i64 n_trans = 0;
  Mat<T> W = A.sub(0, start, m, n);
  n_trans += lu_rec(W, 0, bs, trans, recursion_threshold);
  {
    Mat<T> A00 = W.sub(0, 0, bs, bs), A01 = W.sub(0, bs, bs, n - bs), A10 = W.sub(bs, 0, m - bs, bs),
           A11 = W.sub(bs, bs, m - bs, n - bs);
    Mat<const T> cA00{A00.p, A00.m, A00.n, A00.rs, A00.cs}, cA10{A10.p, A10.m, A10.n, A10.rs, A10.cs},
        cA01{A01.p, A01.m, A01.n, A01.rs, A01.cs};
    solve_lower<T>(cA00, false, true, A01);
    matmul<T>(A11, true, cA10, false, cA01, false, T(-1));
    n_trans += lu_rec(W.sub(bs, 0, m - bs, n), bs, n, trans + bs, recursion_threshold);
  }
  // apply this level's transpositions to the columns outside [start, end)
  auto swap_cols = [&](Mat<T> M) {
    for (i64 c = 0; c < M.n; ++c)
      for (i64 j = 0; j < n; ++j) std::swap(M(j, c), M(j + trans[j], c));
  };
  swap_cols(A.sub(0, 0, m, start));
  swap_cols(A.sub(0, end, m, ncols - end));
  return n_trans;
}

template <class T>
i64 lu_in_place(Mat<T> A, i64* perm, i64* perm_inv, i64 recursion_threshold) {
  const i64 m = A.m, n = A.n, size = std::min(m, n);
  for (i64 i = 0; i < m; ++i) perm[i] = i;
  std::vector<i64> trans((size_t)size, 0);
  const i64 n_trans = lu_rec(A, 0, size, trans.data(), recursion_threshold);
  for (i64 i = 0; i < size; ++i) std::swap(perm[i], perm[i + trans[i]]);
  if (m < n) {
    Mat<T> left = A.sub(0, 0, m, size), right = A.sub(0, size, m, n - size);
    solve_lower<T>(Mat<const T>{left.p, left.m, left.n, left.rs, left.cs}, false, true, right);
  }
  for (i64 i = 0; i < m; ++i) perm_inv[perm[i]] = i;
  return n_trans;
}

// explicit instantiations
#define ORACLE_INST(T)                                                                                           \
  template void matmul<T>(Mat<T>, bool, Mat<const T>, bool, Mat<const T>, bool, T);                              \
  template void matmul_triangular<T>(Mat<T>, int, bool, Mat<const T>, int, bool, Mat<const T>, int, bool, T);    \
  template void solve_lower<T>(Mat<const T>, bool, bool, Mat<T>);                                                \
  template void solve_upper<T>(Mat<const T>, bool, bool, Mat<T>);                                                \
  template i64 llt_in_place<T>(Mat<T>, real_of<T>::type, real_of<T>::type, i64, i64, i64*);                      \
  template i64 ldlt_in_place<T>(Mat<T>, real_of<T>::type, real_of<T>::type, const signed char*, i64, i64, i64*);  \
  template i64 lu_in_place<T>(Mat<T>, i64*, i64*, i64);
ORACLE_INST(double)
ORACLE_INST(float)
ORACLE_INST(std::complex<double>)
ORACLE_INST(std::complex<float>)

}  // namespace oracle
