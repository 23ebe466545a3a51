// Host build of csrc/ldlt_core.cuh: the flat-map launch sequence of the generic LDLT factorization and the bodies of the solve /
// reconstruct / inverse compositions, run thread by thread on the CPU (tests/test_ldlt_types_emul_cpu.py) over the same index space
// a CUDA launch visits (i rounded up to whole 256-thread blocks), forwards or in reverse thread order.
// TEST INFRASTRUCTURE: nothing in the product links this file.
#include <vector>

#include "../../faer-rs_b200/csrc/ldlt_core.cuh"

using namespace ldl;

namespace {
struct HostRun {
  bool reverse;
  template <class B>
  void operator()(const B& body, i64 nx, i64 ny) {
    if (nx <= 0 || ny <= 0) return;
    const i64 gx = ((nx + 255) / 256) * 256;
    if (!reverse) {
      for (i64 j = 0; j < ny; ++j)
        for (i64 i = 0; i < gx; ++i) body(i, j);
    } else {
      for (i64 j = ny - 1; j >= 0; --j)
        for (i64 i = gx - 1; i >= 0; --i) body(i, j);
    }
  }
};

template <class R, bool CX>
void factor(void* A, i64 rs, i64 cs, i64 n, void* D, const signed char* signs, double delta, double eps, i64* info, int reverse) {
  HostRun run{reverse != 0};
  std::vector<R> w((size_t)(2 * n + 2)), sc(4);
  info[0] = -1; info[1] = 0;
  factor_unblocked<R, CX>(run, (R*)A, rs, cs, n, (R*)D, signs, (R)delta, (R)eps, info, sc.data(), w.data());
}
template <class R, bool CX>
void bodies(int which, void* X, i64 rs, i64 cs, i64 n, i64 k, const void* Dv, i64 dstride, void* out, int reverse) {
  HostRun run{reverse != 0};
  switch (which) {
    case 0: run(RecipDiag<R, CX>{(const R*)Dv, dstride, n, (R*)out}, n, 1); break;
    case 1: run(ScaleRows<R, CX>{(R*)X, rs, cs, n, k, (const R*)Dv}, n, k); break;
    case 2: run(BuildLxD<R, CX>{(const R*)X, rs, cs, (const R*)Dv, dstride, (R*)out, n, n}, n, n); break;
    case 3: run(SetIdentity<R, CX>{(R*)out, n, n}, n, n); break;
    case 4: run(FillUpperAdjoint<R, CX>{(R*)X, n, n, (const R*)Dv}, n, n); break;
  }
}
}  // namespace

extern "C" {
// kind: 0 = f32, 1 = f64, 2 = c32, 3 = c64
void ldlt_emul_factor(int kind, void* A, i64 rs, i64 cs, i64 n, void* D, const signed char* signs, double delta, double eps, i64* info,
                      int reverse) {
  switch (kind) {
    case 0: factor<float, false>(A, rs, cs, n, D, signs, delta, eps, info, reverse); break;
    case 1: factor<double, false>(A, rs, cs, n, D, signs, delta, eps, info, reverse); break;
    case 2: factor<float, true>(A, rs, cs, n, D, signs, delta, eps, info, reverse); break;
    case 3: factor<double, true>(A, rs, cs, n, D, signs, delta, eps, info, reverse); break;
  }
}
void ldlt_emul_body(int kind, int which, void* X, i64 rs, i64 cs, i64 n, i64 k, const void* Dv, i64 dstride, void* out, int reverse) {
  switch (kind) {
    case 0: bodies<float, false>(which, X, rs, cs, n, k, Dv, dstride, out, reverse); break;
    case 1: bodies<double, false>(which, X, rs, cs, n, k, Dv, dstride, out, reverse); break;
    case 2: bodies<float, true>(which, X, rs, cs, n, k, Dv, dstride, out, reverse); break;
    case 3: bodies<double, true>(which, X, rs, cs, n, k, Dv, dstride, out, reverse); break;
  }
}
}
