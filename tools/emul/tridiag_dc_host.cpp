// Host run of the divide-and-conquer tridiagonal eigensolver: the SAME scalar routines the GPU kernels call
// (faer-rs_b200/csrc/tridiag_dc_core.cuh: implicit QL leaf, deflation scan, secular root finder) with the kernels' glue
// (merge sort by ranks, Gu-Eisenstat z-hat, vector formation into the merge matrix W, rotations folded into W's rows, output
// ordering by ranks, Q_new = Q_old * W) written as plain loops in the same order of operations. Test infrastructure for
// tests/test_tridiag_dc_cpu.py, not product code.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../faer-rs_b200/csrc/tridiag_dc_core.cuh"

namespace {
const int LEAF = 32;
struct Merge { int lo, mid, hi; };

// balanced tree: Lv levels, 2^Lv leaves of 16..32 rows; block j of level l is [ (j n) >> l, ((j + 1) n) >> l )
int num_levels(int n) {
  int lv = 0;
  while (((long long)n + (1ll << lv) - 1) >> lv > LEAF) ++lv;
  return lv;
}
inline int bnd(long long j, int l, int n) { return (int)((j * n) >> l); }
}  // namespace

extern "C" int tdc_eig(const double* d_in, const double* e_in, int n, double* lam, double* Q /* n x n column-major */) {
  using namespace fb::dc;
  std::vector<double> d(d_in, d_in + n), e(n, 0.0);
  for (int i = 0; i + 1 < n; ++i) e[i] = e_in[i];
  double scale = 0;
  for (int i = 0; i < n; ++i) scale = std::max(scale, std::max(std::fabs(d[i]), std::fabs(e[i])));
  if (!(scale < INFINITY)) return -1;
  if (scale == 0) scale = 1;
  for (int i = 0; i < n; ++i) { d[i] /= scale; e[i] /= scale; }
  const int Lv = num_levels(n);
  std::vector<std::vector<Merge>> levels(Lv);
  std::vector<std::pair<int, int>> leaves;
  for (int l = 0; l < Lv; ++l)
    for (int i = 0; i < (1 << l); ++i) levels[l].push_back({bnd(i, l, n), bnd(2 * i + 1, l + 1, n), bnd(i + 1, l, n)});
  for (int i = 0; i < (1 << Lv); ++i) leaves.push_back({bnd(i, Lv, n), bnd(i + 1, Lv, n)});
  // tear
  std::vector<std::vector<double>> rho(levels.size()), sgn(levels.size());
  for (size_t L = 0; L < levels.size(); ++L)
    for (auto& m : levels[L]) {
      const double beta = e[m.mid - 1];
      rho[L].push_back(std::fabs(beta));
      sgn[L].push_back(beta < 0 ? -1.0 : 1.0);
      d[m.mid - 1] -= std::fabs(beta);
      d[m.mid] -= std::fabs(beta);
    }
  std::vector<double> Qa((size_t)n * n, 0.0), Qb((size_t)n * n, 0.0), W((size_t)n * n), DEL((size_t)n * n);
  // leaves
  for (auto& lf : leaves) {
    const int lo = lf.first, s = lf.second - lf.first;
    double dd[LEAF], ee[LEAF], z[LEAF][LEAF];
    for (int i = 0; i < s; ++i) { dd[i] = d[lo + i]; ee[i] = (i + 1 < s) ? e[lo + i] : 0.0; }
    for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) z[i][j] = i == j;
    if (!ql_implicit(dd, ee, s, [&](int i, double c, double sn) {
          for (int k = 0; k < s; ++k) { const double f = z[k][i + 1]; z[k][i + 1] = sn * z[k][i] + c * f; z[k][i] = c * z[k][i] - sn * f; }
        })) return -2;
    // selection sort ascending with column swaps
    for (int i = 0; i < s; ++i) {
      int kmin = i;
      for (int j = i + 1; j < s; ++j) if (dd[j] < dd[kmin]) kmin = j;
      if (kmin != i) { std::swap(dd[i], dd[kmin]); for (int k = 0; k < s; ++k) std::swap(z[k][i], z[k][kmin]); }
    }
    for (int i = 0; i < s; ++i) d[lo + i] = dd[i];
    for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) Qa[(size_t)(lo + j) * n + lo + i] = z[i][j];
  }
  std::vector<double> dsort(n), zsort(n), dl(n), w(n), ddv(n), rc(n), rs(n), tau(n), zh(n), lamv(n), dnew(n);
  std::vector<int> colsort(n), cnd(n), cdf(n), ra(n), rb(n), org(n), outcol(n);
  double* Qold = Qa.data();
  double* Qnew = Qb.data();
  for (int L = (int)levels.size() - 1; L >= 0; --L) {
    std::fill(Qnew, Qnew + (size_t)n * n, 0.0);
    for (size_t mi = 0; mi < levels[L].size(); ++mi) {
      const Merge m = levels[L][mi];
      const int lo = m.lo, mid = m.mid, hi = m.hi, s = hi - lo, s1 = mid - lo;
      const double rho2 = 2.0 * rho[L][mi];
      // z and ranks
      for (int i = lo; i < hi; ++i) {
        const double zi = (i < mid ? Qold[(size_t)i * n + (mid - 1)] : sgn[L][mi] * Qold[(size_t)i * n + mid]) * M_SQRT1_2;
        int pos;
        if (i < mid) pos = (i - lo) + (int)(std::lower_bound(d.begin() + mid, d.begin() + hi, d[i]) - (d.begin() + mid));
        else pos = (i - mid) + (int)(std::upper_bound(d.begin() + lo, d.begin() + mid, d[i]) - (d.begin() + lo));
        dsort[lo + pos] = d[i]; zsort[lo + pos] = zi; colsort[lo + pos] = i - lo;
      }
      int nrot = 0;
      const int k = deflate_scan(&dsort[lo], &zsort[lo], &colsort[lo], s, rho2, &dl[lo], &w[lo], &cnd[lo], &ddv[lo], &cdf[lo], &ra[lo], &rb[lo],
                                 &rc[lo], &rs[lo], &nrot);
      // secular roots + DELTA
      double zz = 0;
      for (int i = 0; i < k; ++i) zz += w[lo + i] * w[lo + i];
      for (int j = 0; j < k; ++j) {
        int og;
        const double rhoinv = 1.0 / rho2;
        const double t = (k == 1) ? rho2 * zz : secular_root(&dl[lo], k, j, rho2, zz, [&](int o, double tt) { return secular_f(&dl[lo], &w[lo], k, o, tt, rhoinv); }, &og);
        if (k == 1) og = 0;
        lamv[lo + j] = dl[lo + og] + t;
        for (int i = 0; i < k; ++i) DEL[(size_t)j * s + i] = (dl[lo + i] - dl[lo + og]) - t;
      }
      // z-hat
      for (int i = 0; i < k; ++i) {
        double prod = DEL[(size_t)i * s + i];
        for (int j = 0; j < k; ++j) if (j != i) prod *= DEL[(size_t)j * s + i] / (dl[lo + i] - dl[lo + j]);
        zh[lo + i] = std::copysign(std::sqrt(std::fabs(prod)), w[lo + i]);
      }
      // output order (ranks among all s new eigenvalues: roots first on ties)
      const int nd = s - k;
      for (int j = 0; j < k; ++j) {
        int c = j;
        for (int t = 0; t < nd; ++t) c += ddv[lo + t] < lamv[lo + j];
        outcol[lo + j] = c;
      }
      for (int t = 0; t < nd; ++t) {
        int c = 0;
        for (int j = 0; j < k; ++j) c += lamv[lo + j] <= ddv[lo + t];
        for (int u = 0; u < nd; ++u) c += (ddv[lo + u] < ddv[lo + t]) || (ddv[lo + u] == ddv[lo + t] && u < t);
        outcol[lo + k + t] = c;
      }
      // W (s x s, column-major, ld = s)
      std::fill(W.begin(), W.begin() + (size_t)s * s, 0.0);
      for (int j = 0; j < k; ++j) {
        double nrm = 0;
        for (int i = 0; i < k; ++i) { const double v = zh[lo + i] / DEL[(size_t)j * s + i]; nrm += v * v; }
        nrm = std::sqrt(nrm);
        for (int i = 0; i < k; ++i) W[(size_t)outcol[lo + j] * s + cnd[lo + i]] = zh[lo + i] / DEL[(size_t)j * s + i] / nrm;
        dnew[lo + outcol[lo + j]] = lamv[lo + j];
      }
      for (int t = 0; t < nd; ++t) {
        W[(size_t)outcol[lo + k + t] * s + cdf[lo + t]] = 1.0;
        dnew[lo + outcol[lo + k + t]] = ddv[lo + t];
      }
      // rotations folded into the rows of W, last rotation first
      for (int r = nrot - 1; r >= 0; --r) {
        const int a = ra[lo + r], b = rb[lo + r];
        const double c = rc[lo + r], sn = rs[lo + r];
        for (int col = 0; col < s; ++col) {
          const double x = W[(size_t)col * s + a], y = W[(size_t)col * s + b];
          W[(size_t)col * s + a] = c * x - sn * y;
          W[(size_t)col * s + b] = sn * x + c * y;
        }
      }
      // Q_new block = blockdiag(Q1, Q2) * W
      for (int col = 0; col < s; ++col) {
        for (int kk = 0; kk < s; ++kk) {
          const double wv = W[(size_t)col * s + kk];
          if (wv == 0.0) continue;
          const int r0 = kk < s1 ? lo : mid, r1 = kk < s1 ? mid : hi;
          for (int r = r0; r < r1; ++r) Qnew[(size_t)(lo + col) * n + r] += Qold[(size_t)(lo + kk) * n + r] * wv;
        }
      }
      for (int i = 0; i < s; ++i) d[lo + i] = dnew[lo + i];
    }
    std::swap(Qold, Qnew);
  }
  for (int i = 0; i < n; ++i) lam[i] = d[i] * scale;
  std::memcpy(Q, Qold, sizeof(double) * (size_t)n * n);
  return 0;
}
