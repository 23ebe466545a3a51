// Host transcription of ldlf2_kernel (faer-rs_b200/csrc/ldlt_f64.cu) for checking its index arithmetic, ownership rules,
// fma operands and failure path on a machine without a GPU. NOT product code and not a fallback: it is built and run by
// tests/test_ldlf2_emulation_cpu.py only. KEEP IN SYNC with the kernel: same constants, same per-thread state, same
// statements, in the same order. Threads are executed one after the other inside each barrier interval, which is faithful
// because within an interval the kernel only reads the buffers of parity (j & 1) written before the previous barrier and
// writes those of parity ((j + 1) & 1) (or per-thread registers, or global entries it alone owns).
#include <cmath>
#include <cstring>
#include <vector>

typedef long long i64;
namespace {
constexpr int LDL_MAX = 128;
constexpr int LDL_UPD_WARPS = 16;
constexpr int LDL_THREADS = LDL_UPD_WARPS * 32 + LDL_MAX;
constexpr int LDL_CB = LDL_MAX / LDL_UPD_WARPS;

struct Shared {
  double colbuf[2][LDL_MAX];
  double s_inv[2];
  double s_d[2];
  int s_fail[2];
  int s_count;
};
struct Regs {
  double a[4][LDL_CB];
  double dp;
};
}  // namespace

extern "C" int emu_ldlf2(double* A, i64 rs, i64 cs, int n, i64 j0, int regularize, double eps, double delta,
                         const signed char* signs, long long* info) {
  Shared sh;
  std::memset(&sh, 0, sizeof(sh));
  std::vector<Regs> regs(LDL_THREADS);
  if (info[0] >= 0) return 0;
  sh.s_count = 0;

  auto publish_pivot = [&](int jc, double d) {
    if (regularize) {
      const int sign = signs ? (int)signs[j0 + jc] : 0;
      const bool small_or_negative = d <= eps;
      const bool minus_small_or_positive = d >= -eps;
      if (sign == 1 && small_or_negative) {
        d = delta;
        sh.s_count += 1;
      } else if (sign == -1 && minus_small_or_positive) {
        d = -delta;
      } else if (small_or_negative && minus_small_or_positive) {
        d = d < 0.0 ? -delta : delta;
      }
    }
    const int fail = (d == 0.0 || !std::isfinite(d)) ? 1 : 0;
    sh.s_d[jc & 1] = d;
    sh.s_inv[jc & 1] = fail ? 0.0 : 1.0 / d;
    sh.s_fail[jc & 1] = fail;
  };

  // ---- interval 0: loads ----
  for (int tid = 0; tid < LDL_THREADS; ++tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const bool is_upd = warp < LDL_UPD_WARPS;
    const int p = tid - LDL_UPD_WARPS * 32;
    Regs& r = regs[tid];
    r.dp = 0.0;
    if (is_upd) {
      for (int ai = 0; ai < 4; ++ai)
        for (int bi = 0; bi < LDL_CB; ++bi) {
          const int i = lane + 32 * ai, c = warp + LDL_UPD_WARPS * bi;
          r.a[ai][bi] = (i < n && c <= i) ? A[(i64)i * rs + (i64)c * cs] : 0.0;
        }
    } else if (p < n) {
      r.dp = A[(i64)p * rs + (i64)p * cs];
    }
  }
  // ---- interval 1: column 0 published ----
  for (int tid = 0; tid < LDL_THREADS; ++tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const bool is_upd = warp < LDL_UPD_WARPS;
    const int p = tid - LDL_UPD_WARPS * 32;
    Regs& r = regs[tid];
    if (is_upd) {
      if (warp == 0)
        for (int ai = 0; ai < 4; ++ai) sh.colbuf[0][lane + 32 * ai] = r.a[ai][0];
    } else if (p == 0) {
      publish_pivot(0, r.dp);
    }
  }
  // ---- the column loop: one interval per column ----
  for (int j = 0; j < n; ++j) {
    const double* col = sh.colbuf[j & 1];
    const double d = sh.s_d[j & 1];
    if (sh.s_fail[j & 1]) {
      info[0] = j0 + j;                           // tid == 0
      A[(i64)j * rs + (i64)j * cs] = d;
      return 0;
    }
    const double inv = sh.s_inv[j & 1];
    const double nd = -d;
    for (int tid = 0; tid < LDL_THREADS; ++tid) {
      const int lane = tid & 31, warp = tid >> 5;
      const bool is_upd = warp < LDL_UPD_WARPS;
      const int p = tid - LDL_UPD_WARPS * 32;
      Regs& r = regs[tid];
      if (!is_upd) {
        if (p > j && p < n) {
          const double l = col[p] * inv;
          r.dp = std::fma(l * nd, l, r.dp);
          if (p == j + 1) publish_pivot(j + 1, r.dp);
        }
      } else {
        const int jw = j & (LDL_UPD_WARPS - 1);
        if (warp == jw) {
          for (int ai = 0; ai < 4; ++ai) {
            const int i = lane + 32 * ai;
            if (i > j && i < n) A[(i64)i * rs + (i64)j * cs] = col[i] * inv;
            else if (i == j) A[(i64)i * rs + (i64)j * cs] = d;
          }
        }
        if (warp + LDL_UPD_WARPS * (LDL_CB - 1) > j) {
          double li[4];
          for (int ai = 0; ai < 4; ++ai) li[ai] = col[lane + 32 * ai] * inv;
          for (int bi = 0; bi < LDL_CB; ++bi) {
            const int c = warp + LDL_UPD_WARPS * bi;
            if (c > j && c < n) {
              const double lcd = (col[c] * inv) * nd;
              for (int ai = 0; ai < 4; ++ai) {
                const int i = lane + 32 * ai;
                if (32 * ai + 31 >= c) {
                  if (i >= c && i < n) r.a[ai][bi] = std::fma(lcd, li[ai], r.a[ai][bi]);
                }
              }
            }
          }
        }
        if (j + 1 < n && warp == ((j + 1) & (LDL_UPD_WARPS - 1))) {
          const int nbk = (j + 1) / LDL_UPD_WARPS;
          double* nxt = sh.colbuf[(j + 1) & 1];
          for (int ai = 0; ai < 4; ++ai) {
            double v = 0.0;
            for (int bi = 0; bi < LDL_CB; ++bi)
              if (bi == nbk) v = r.a[ai][bi];
            nxt[lane + 32 * ai] = v;
          }
        }
      }
    }
  }
  if (sh.s_count) info[1] += sh.s_count;
  return 0;
}
