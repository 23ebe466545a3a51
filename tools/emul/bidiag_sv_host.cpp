// Host build of faer-rs_b200/csrc/bidiag_sv.cuh for tests/test_bidiag_sv_cpu.py: the same functions the GPU kernel calls,
// one "thread" per singular value in a loop. Test infrastructure, not product code.
#include "../../faer-rs_b200/csrc/bidiag_sv.cuh"

template <class T>
static void run(const T* d, const T* e, int n, T* out) {
  T bound = 0, bmax2 = 0;
  for (int i = 0; i < n; ++i) {
    const T a = std::fabs(d[i]), b = i + 1 < n ? std::fabs(e[i]) : T(0), c = i > 0 ? std::fabs(e[i - 1]) : T(0);
    bound = std::fmax(bound, std::fmax(a + b, a + c));  // Gershgorin on T_GK: rows (e_{i-1}, d_i) and (d_i, e_i)
    bmax2 = std::fmax(bmax2, std::fmax(a * a, b * b));
  }
  bound = bound * (T(1) + T(4) * fb::bsv::Lim<T>::eps()) + fb::bsv::Lim<T>::safmin();
  for (int k = 0; k < n; ++k) out[k] = fb::bsv::gk_kth_largest<T>(d, e, n, 1, 1, k, bound, bmax2);
}
extern "C" void bsv_f64(const double* d, const double* e, int n, double* out) { run<double>(d, e, n, out); }
extern "C" void bsv_f32(const float* d, const float* e, int n, float* out) { run<float>(d, e, n, out); }
