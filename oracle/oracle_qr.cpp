// TEST INFRASTRUCTURE — NOT PRODUCT CODE. CPU restatement of faer's Householder QR (no pivoting) and the block
// Householder machinery. See oracle.hpp for the contract.
//
// Reference:
//   reductions::norm_l2                       faer/src/linalg/reductions/norm_l2.rs:6-172
//   householder::make_householder_imp         faer/src/linalg/householder.rs:59-107
//   householder::upgrade_householder_factor   householder.rs:132-272   (T = striu(V^H V), diag = tau)
//   apply_block_householder_*_on_the_left     householder.rs:370-620, 624-719
//   qr::no_pivoting::factor::qr_in_place      faer/src/linalg/qr/no_pivoting/factor.rs:11-86 (unblocked),
//                                             137-256 (blocked), 258-301 (driver), 91-116 (recommended_block_size)
// faer's convention: H = I - v v^H / tau, v_0 = 1, tau = (1 + |v_tail|^2)/2 (= 1/tau_LAPACK);
// block form H_0...H_{b-1} = I - V T^{-1} V^H.
#include <algorithm>
#include <cstring>

#include "oracle.hpp"

namespace oracle {

template <class T> static inline Mat<const T> cst(Mat<T> a) { return Mat<const T>{a.p, a.m, a.n, a.rs, a.cs}; }
template <class T> static inline typename real_of<T>::type abs_(T x) { return std::abs(x); }
template <class T> static inline typename real_of<T>::type abs2_(T x) { return x * x; }
template <class R> static inline R abs2_(std::complex<R> x) { return x.real() * x.real() + x.imag() * x.imag(); }
template <class T> static inline T cj(T x) { return x; }
template <class R> static inline std::complex<R> cj(std::complex<R> x) { return std::conj(x); }

// norm_l2 of a strided vector: three scaled accumulators (x*sqrt(min_pos), x, x*sqrt(max_pos)), then the
// overflow/underflow-safe selection of norm_l2.rs:161-172. (The reference's SIMD lane / pairwise order differs;
// the scalings are exact powers of two, so the result differs only by summation-order rounding.)
template <class T>
typename real_of<T>::type norm_l2(const T* p, i64 n, i64 stride) {
  typedef typename real_of<T>::type R;
  const R min_pos = std::numeric_limits<R>::min();
  const R sml = std::sqrt(min_pos), big = std::sqrt(R(1) / min_pos);
  R acc_sml = 0, acc_med = 0, acc_big = 0;
  for (i64 i = 0; i < n; ++i) {
    const T x = p[i * stride];
    acc_sml += abs2_(T(x * sml));
    acc_med += abs2_(x);
    acc_big += abs2_(T(x * big));
  }
  if (acc_sml >= R(1)) return std::sqrt(acc_sml) * big;
  if (acc_med >= R(1)) return std::sqrt(acc_med);
  return std::sqrt(acc_big) * sml;
}

// householder.rs:59-107. `out`/`in` are column vectors of length len (in == nullptr -> in place).
template <class T>
HouseholderInfo<T> make_householder(T* head, T* out, i64 out_stride, const T* in, i64 in_stride, i64 len) {
  typedef typename real_of<T>::type R;
  const R min_pos = std::numeric_limits<R>::min();
  const R inf = std::numeric_limits<R>::infinity();
  const T* tail = in ? in : out;
  const i64 tstride = in ? in_stride : out_stride;
  const R tail_norm = norm_l2(tail, len, tstride);
  R head_norm = abs_(*head);
  if (head_norm < min_pos) {
    *head = T(0);
    head_norm = 0;
  }
  if (tail_norm < min_pos) return HouseholderInfo<T>{inf, T(inf), head_norm};
  const R norm = std::hypot(head_norm, tail_norm);
  const T sign = head_norm != R(0) ? T(*head * (R(1) / head_norm)) : T(1);
  const T signed_norm = sign * T(norm);
  const T head_with_beta = *head + signed_norm;
  const T inv = T(1) / head_with_beta;
  for (i64 i = 0; i < len; ++i) out[i * out_stride] = tail[i * tstride] * inv;
  *head = -signed_norm;
  const R t = tail_norm * abs_(inv);
  const R tau = R(0.5) * (R(1) + t * t);
  return HouseholderInfo<T>{tau, inv, norm};
}

// qr/no_pivoting/factor.rs:11-86. H: row vector (length H_len, element stride h_stride).
template <class T>
static i64 qr_unblocked(Mat<T> A, T* H, i64 H_len, i64 h_stride, i64 row_start, i64 col_start) {
  typedef typename real_of<T>::type R;
  const i64 m = A.m, n = A.n;
  const R min_pos = std::numeric_limits<R>::min();
  i64 col = col_start, row = row_start;
  while (row < std::min(H_len, m) && col < n) {
    R norm = norm_l2(&A(0, col), row, A.rs);
    T* head = &A(row, col);
    const i64 len = m - row - 1;
    HouseholderInfo<T> info;
    const T* v;
    i64 vstride = A.rs;
    if (row == col) {
      info = make_householder<T>(head, len ? &A(row + 1, col) : head, A.rs, nullptr, 0, len);
      v = len ? &A(row + 1, col) : head;
    } else {
      info = make_householder<T>(head, len ? &A(row + 1, row) : head, A.rs, len ? &A(row + 1, col) : head, A.rs, len);
      const i64 z = std::min(len, col - row);
      for (i64 i = 0; i < z; ++i) A(row + 1 + i, col) = T(0);
      v = len ? &A(row + 1, row) : head;
    }
    norm = std::hypot(info.norm, norm);
    const R eps = std::numeric_limits<R>::epsilon();
    const R leeway = R((double)(m - row) * 16.0);
    const R threshold = eps * leeway * norm;
    const R tau_inv = R(1) / info.tau;
    H[row * h_stride] = T(info.tau);
    if (tau_inv < min_pos) {
      if (info.norm > R(0)) row += 1;
    } else if (info.norm > threshold) {
      for (i64 c = col + 1; c < n; ++c) {
        T dot = A(row, c);
        for (i64 i = 0; i < len; ++i) dot = dot + cj(v[i * vstride]) * A(row + 1 + i, c);
        const T k = -(dot * tau_inv);
        A(row, c) = A(row, c) + k;
        for (i64 i = 0; i < len; ++i) A(row + 1 + i, c) = A(row + 1 + i, c) + k * v[i * vstride];
      }
      row += 1;
    }
    col += 1;
  }
  return row;
}

// householder.rs:132-272
template <class T>
void upgrade_householder_factor(Mat<T> Tf, Mat<const T> V, i64 block_size, i64 prev_block_size) {
  if (block_size == prev_block_size || Tf.m <= prev_block_size) return;
  const i64 n = V.n;
  const i64 block_count = (Tf.m + block_size - 1) / block_size;
  if (block_count > 1) {
    const i64 mid = block_count / 2;
    // NB: the reference splits at `mid` (a BLOCK COUNT used as an index), householder.rs:155-157
    upgrade_householder_factor<T>(Tf.sub(0, 0, mid, mid), V.sub(0, 0, V.m, mid), block_size, prev_block_size);
    upgrade_householder_factor<T>(Tf.sub(mid, mid, Tf.m - mid, Tf.n - mid), V.sub(mid, mid, V.m - mid, V.n - mid),
                                  block_size, prev_block_size);
    return;
  }
  if (prev_block_size < 8) {
    Mat<const T> top = V.sub(0, 0, n, n), bot = V.sub(n, 0, V.m - n, n);
    matmul_triangular<T>(Tf, UNIT_UPPER, false, top.t(), UNIT_UPPER, true, top, UNIT_LOWER, false, T(1));
    matmul_triangular<T>(Tf, UNIT_UPPER, true, bot.t(), RECT, true, bot, RECT, false, T(1));
  } else {
    const i64 prev_block_count = (Tf.m + prev_block_size - 1) / prev_block_size;
    const i64 mid = (prev_block_count / 2) * prev_block_size;
    Mat<T> tl = Tf.sub(0, 0, mid, mid), tr = Tf.sub(0, mid, mid, Tf.n - mid), br = Tf.sub(mid, mid, Tf.m - mid, Tf.n - mid);
    Mat<const T> left = V.sub(0, 0, V.m, mid), right = V.sub(mid, mid, V.m - mid, V.n - mid);
    upgrade_householder_factor<T>(tl, left, block_size, prev_block_size);
    upgrade_householder_factor<T>(br, right, block_size, prev_block_size);
    Mat<const T> left2 = left.sub(mid, 0, left.m - mid, left.n);
    const i64 row_mid = right.n;
    Mat<const T> lt = left2.sub(0, 0, row_mid, left2.n), lb = left2.sub(row_mid, 0, left2.m - row_mid, left2.n);
    Mat<const T> rt = right.sub(0, 0, row_mid, right.n), rb = right.sub(row_mid, 0, right.m - row_mid, right.n);
    matmul_triangular<T>(tr, RECT, false, lt.t(), RECT, true, rt, UNIT_LOWER, false, T(1));
    matmul<T>(tr, true, lb.t(), true, rb, false, T(1));
  }
}

// householder.rs:370-620 (general path; the N == 1 SIMD fast path 394-503 computes the same quantities)
// M <- (I - V T^{-1} V^H) M  (forward = false)   or   M <- (I - V T^{-H} V^H) M  (forward = true), with conj_lhs
// conjugating V and T.
template <class T>
void apply_block_householder_on_the_left(Mat<const T> V, Mat<const T> Tf, bool conj_lhs, Mat<T> M, bool forward) {
  const i64 N = V.n, m = V.m, K = M.n;
  if (N == 0 || K == 0) return;
  std::vector<T> tmpbuf((size_t)(N * K));
  Mat<T> tmp{tmpbuf.data(), N, K, 1, N};
  Mat<const T> Vt = V.sub(0, 0, N, N), Vb = V.sub(N, 0, m - N, N);
  Mat<T> top = M.sub(0, 0, N, K), bot = M.sub(N, 0, m - N, K);
  matmul_triangular<T>(tmp, RECT, false, Vt.t(), UNIT_UPPER, !conj_lhs, cst(top), RECT, false, T(1));
  matmul<T>(tmp, true, Vb.t(), !conj_lhs, cst(bot), false, T(1));
  if (forward)
    solve_lower<T>(Tf.t(), !conj_lhs, false, tmp);
  else
    solve_upper<T>(Tf, conj_lhs, false, tmp);
  matmul_triangular<T>(top, RECT, true, Vt, UNIT_LOWER, conj_lhs, cst(tmp), RECT, false, T(-1));
  matmul<T>(bot, true, Vb, conj_lhs, cst(tmp), false, T(-1));
}

// qr/no_pivoting/factor.rs:137-256
template <class T>
static i64 qr_blocked(Mat<T> A, Mat<T> H, i64 row_start, i64 col_start, i64 blocking_threshold) {
  const i64 m = A.m, n = A.n, size = std::min(m, n);
  const i64 block_size0 = H.m;
  if (block_size0 == 1) return qr_unblocked<T>(A, H.p, H.n, H.cs, row_start, col_start);
  const i64 sub_block_size0 = (m * n < blocking_threshold) ? 1 : block_size0 / 2;
  i64 col = col_start, row = row_start;
  while (row < size && col < n) {
    const i64 block_size = std::min(block_size0, std::min(size - row, n - col));
    const i64 sub_block_size = std::min(block_size, sub_block_size0);
    const i64 start = row;
    i64 offset = 0;
    while (offset < block_size && col < n) {
      const i64 bsz = std::min(n - col, block_size - offset);
      const i64 sbs = std::min(bsz, sub_block_size);
      const i64 new_row = qr_blocked<T>(A.sub(0, 0, m, col + bsz), H.sub(offset, 0, sbs, H.n), row, col, blocking_threshold);
      const i64 local = new_row - row;
      if (local > 0) {
        for (i64 k = 0; k < local;) {
          const i64 s2 = std::min(sbs, local - k);
          if (k > 0) {
            // copy the s2 x s2 upper-triangular T block from rows [offset, offset+s2) to rows [offset+k, ...)
            Mat<T> Hs = H.sub(offset, row + k, H.m - offset, s2);
            for (i64 j = 0; j < s2; ++j)
              for (i64 i = 0; i <= j; ++i) Hs(k + i, j) = Hs(i, j);
          }
          k += s2;
        }
        upgrade_householder_factor<T>(H.sub(offset, row, local, local), cst(A.sub(row, row, m - row, local)), local, sbs);
        if (offset > 0) {
          Mat<T> Hh = H.sub(0, start, offset + local, offset + local);
          Mat<const T> Aa = cst(A.sub(start, start, m - start, offset + local));
          Mat<const T> A0 = Aa.sub(0, 0, offset + local, Aa.n), A1 = Aa.sub(offset + local, 0, Aa.m - (offset + local), Aa.n);
          matmul_triangular<T>(Hh, UNIT_UPPER, false, A0.t(), UNIT_UPPER, true, A0, UNIT_LOWER, false, T(1));
          matmul_triangular<T>(Hh, UNIT_UPPER, true, A1.t(), RECT, true, A1, RECT, false, T(1));
        }
      }
      Mat<T> below = A.sub(row, 0, m - row, n);
      Mat<const T> Q0 = cst(below.sub(0, row, m - row, local));
      Mat<T> A1 = below.sub(0, col + bsz, m - row, n - (col + bsz));
      Mat<const T> Hq = cst(H.sub(offset, row, local, local));
      if (A1.n > 0) apply_block_householder_on_the_left<T>(Q0, Hq, /*conj_lhs = Yes∘Yes =*/false, A1, /*forward=*/true);
      offset += local;
      row += local;
      col += bsz;
    }
  }
  return row;
}

// qr/no_pivoting/factor.rs:258-301. Q_coeff: block_size x min(m, n). returns the rank.
template <class T>
i64 qr_in_place(Mat<T> A, Mat<T> Q_coeff, i64 blocking_threshold) {
  typedef typename real_of<T>::type R;
  const i64 block_size = Q_coeff.m;
  const i64 rank = qr_blocked<T>(A, Q_coeff, 0, 0, blocking_threshold);
  for (i64 j = rank; j < Q_coeff.n; ++j)
    for (i64 i = 0; i < Q_coeff.m; ++i) Q_coeff(i, j) = T(0);
  i64 col = rank / block_size * block_size;
  const i64 n = Q_coeff.n;
  while (col < n) {
    const i64 bs = std::min(block_size, n - col);
    const i64 start = std::max(rank, col);
    // diagonal of Q_coeff[start-col.., start..col+bs) <- +inf
    for (i64 d = 0; start - col + d < Q_coeff.m && start + d < col + bs; ++d)
      Q_coeff(start - col + d, start + d) = T(std::numeric_limits<R>::infinity());
    col += bs;
  }
  return rank;
}

// qr/no_pivoting/factor.rs:91-116
i64 qr_recommended_block_size(i64 nrows, i64 ncols) {
  const i64 prod = nrows * ncols, size = std::min(nrows, ncols);
  i64 bs;
  if (prod > 8192ll * 8192) bs = 256;
  else if (prod > 2048 * 2048) bs = 128;
  else if (prod > 1024 * 1024) bs = 64;
  else if (prod > 512 * 512) bs = 48;
  else if (prod > 128 * 128) bs = 32;
  else if (prod > 32 * 32) bs = 8;
  else if (prod > 16 * 16) bs = 4;
  else bs = 1;
  return std::max<i64>(1, std::min(bs, size));
}

#define ORACLE_QR_INST(T)                                                                           \
  template i64 qr_in_place<T>(Mat<T>, Mat<T>, i64);                                                 \
  template void apply_block_householder_on_the_left<T>(Mat<const T>, Mat<const T>, bool, Mat<T>, bool); \
  template real_of<T>::type norm_l2<T>(const T*, i64, i64);                                         \
  template HouseholderInfo<T> make_householder<T>(T*, T*, i64, const T*, i64, i64);                 \
  template void upgrade_householder_factor<T>(Mat<T>, Mat<const T>, i64, i64);
ORACLE_QR_INST(double)
ORACLE_QR_INST(float)
ORACLE_QR_INST(std::complex<double>)
ORACLE_QR_INST(std::complex<float>)

}  // namespace oracle
