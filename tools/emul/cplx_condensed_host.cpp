// Host build of csrc/cplx_condensed_core.cuh: the launch sequences of the complex tridiagonalization / bidiagonalization run
// thread by thread on the CPU (tests/test_cplx_condensed_emul_cpu.py). The launcher visits the same (i, j) index space a
// CUDA launch would (i rounded up to whole 256-thread blocks, so the bounds checks of every body are exercised), forwards or —
// `reverse` — backwards: identical results in both orders show that no body depends on another body of the same launch.
// TEST INFRASTRUCTURE: nothing in the product links this file.
#include <cstring>
#include <vector>

#include "../../faer-rs_b200/csrc/cplx_condensed_core.cuh"

using namespace cc;

namespace {
struct HostRun {
  bool reverse;
  long long launches = 0;
  template <class B>
  void operator()(const B& body, i64 nx, i64 ny) {
    ++launches;
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
struct HostWork {
  std::vector<Cx> v, p, w;
  std::vector<double> part, sc;
  Work ws;
  explicit HostWork(i64 len) : v(len + 1), p(len + 1), w(len + 1), part(3 * NP), sc(SC_COUNT) {
    ws.v = v.data(); ws.p = p.data(); ws.w = w.data(); ws.part = part.data(); ws.sc = sc.data();
  }
};
}  // namespace

extern "C" {

// A: n x n interleaved complex, element strides (rs, cs); lower triangle read. Outputs: W (n x n column-major), tau (n - 1),
// d (n), e (n - 1), ph (n complex), tauc (n - 1 complex). Returns the number of launches.
long long cc_emul_tridiag(const double* A, i64 rs, i64 cs, i64 n, double* W_, double* tau, double* d, double* e, double* ph_, double* tauc_,
                          int reverse) {
  Cx* W = (Cx*)W_;
  HostRun run{reverse != 0};
  HostWork hw(n);
  run(BuildHermitian{A, rs, cs, W, n, n}, n, n);
  tridiag_unblocked(run, W, n, n, tau, hw.ws);
  run(TridiagPhases{W, n, n, tau, d, e, (Cx*)ph_, (Cx*)tauc_}, 1, 1);
  return run.launches;
}

// A: interleaved complex view; adjoint = 1 works on A^H. W: m x n column-major with m >= n AFTER the optional adjoint.
long long cc_emul_bidiag(const double* A, i64 rs, i64 cs, i64 m, i64 n, int adjoint, double* W_, double* tl, double* tr, double* d, double* f,
                         double* l_, double* r_, double* tlc_, double* trc_, int reverse) {
  Cx* W = (Cx*)W_;
  HostRun run{reverse != 0};
  HostWork hw(m > n ? m : n);
  run(CopyIn{A, rs, cs, W, m, m, n, adjoint}, m, n);
  bidiag_unblocked(run, W, m, m, n, tl, tr, hw.ws);
  run(BidiagPhases{W, m, n, tl, tr, d, f, (Cx*)l_, (Cx*)r_, (Cx*)tlc_, (Cx*)trc_}, 1, 1);
  return run.launches;
}

// the assembly bodies: out (rows x cols) = [diag(ph) Q, 0; 0, I]; T = transposed n x n corner of W; strided copies out
void cc_emul_scale_rows_embed(const double* Q, i64 ldq, i64 nq, const double* ph, double* out, i64 rows, i64 cols, int reverse) {
  HostRun run{reverse != 0};
  run(ScaleRowsEmbed{Q, ldq, nq, (const Cx*)ph, (Cx*)out, rows, rows, cols}, rows, cols);
}
void cc_emul_transpose_corner(const double* W, i64 ld, double* T, i64 n, int reverse) {
  HostRun run{reverse != 0};
  run(TransposeCorner{(const Cx*)W, ld, (Cx*)T, n, n}, n, n);
}
void cc_emul_copy_out_f64(double* out, i64 rs, i64 cs, const double* src, i64 rows, i64 cols, int reverse) {
  HostRun run{reverse != 0};
  run(CopyOut<double>{out, rs, cs, (const Cx*)src, rows, rows, cols}, rows, cols);
}
void cc_emul_copy_out_f32(float* out, i64 rs, i64 cs, const double* src, i64 rows, i64 cols, int reverse) {
  HostRun run{reverse != 0};
  run(CopyOut<float>{out, rs, cs, (const Cx*)src, rows, rows, cols}, rows, cols);
}
void cc_emul_copy_values_f64(double* S, i64 stride, const double* vals, i64 n) {
  HostRun run{false};
  run(CopyValues<double>{S, stride, vals, n}, n, 1);
}
void cc_emul_widen_c32(const float* A, i64 rs, i64 cs, double* W, i64 m, i64 n) {
  HostRun run{false};
  run(WidenC32{A, rs, cs, (Cx*)W, m, m, n}, m, n);
}


// A: n x n interleaved complex (rs, cs). Outputs: W (n x n column-major: H + reflectors), tau (n - 1), Tf (bs x (n - 1) T blocks).
long long cc_emul_hessenberg(const double* A, i64 rs, i64 cs, i64 n, double* W_, double* tau, double* Tf, i64 bs, int reverse) {
  Cx* W = (Cx*)W_;
  HostRun run{reverse != 0};
  HostWork hw(n);
  run(CopyIn{A, rs, cs, W, n, n, n, 0}, n, n);
  hessenberg_unblocked(run, W, n, n, tau, hw.ws);
  if (n > 1) run(BuildTBlocks{W + 1, n, n - 1, n - 1, tau, (Cx*)Tf, bs, bs}, bs, n - 1);
  return run.launches;
}
}  // extern "C"
