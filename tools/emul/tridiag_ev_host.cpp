// Host build of faer-rs_b200/csrc/tridiag_ev.cuh for tests/test_tridiag_ev_cpu.py (test infrastructure, not product code).
#include "../../faer-rs_b200/csrc/tridiag_ev.cuh"

template <class T>
static void run(const T* d, const T* e, int n, T* out) {
  T glo = d[0], ghi = d[0], emax2 = 0;
  for (int i = 0; i < n; ++i) {
    const T r = (i > 0 ? std::fabs(e[i - 1]) : T(0)) + (i + 1 < n ? std::fabs(e[i]) : T(0));
    glo = std::fmin(glo, d[i] - r);
    ghi = std::fmax(ghi, d[i] + r);
    if (i + 1 < n) emax2 = std::fmax(emax2, e[i] * e[i]);
  }
  const T span = std::fmax(std::fabs(glo), std::fabs(ghi));
  const T pad = T(4) * fb::tev::Lim<T>::eps() * span * T(n) + fb::tev::Lim<T>::safmin();
  glo -= pad;
  ghi += pad;
  for (int k = 0; k < n; ++k) out[k] = fb::tev::st_kth_smallest<T>(d, e, n, k, glo, ghi, emax2);
}
extern "C" void tev_f64(const double* d, const double* e, int n, double* out) { run<double>(d, e, n, out); }
extern "C" void tev_f32(const float* d, const float* e, int n, float* out) { run<float>(d, e, n, out); }
