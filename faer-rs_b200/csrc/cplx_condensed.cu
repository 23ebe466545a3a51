// `svd` and `self_adjoint_evd` for complex T (c64; c32 is computed in c64 like the f32 entry points are computed in f64).
//
// Reference:
//   svd::svd / svd_imp      faer/src/linalg/svd/mod.rs:530-672, 326-431   (wide inputs through the adjoint, bidiagonalize, phase
//                                                                          normalisation 171-273, real bidiagonal SVD, back-transforms
//                                                                          403-429: left sequence Conj::No, right sequence Conj::Yes on
//                                                                          the transposed rows)
//   evd::self_adjoint_evd   evd/mod.rs:270-418                             (tridiagonalize, real tridiagonal EVD of the phase-normalised
//                                                                          form, back-transform 411-418)
// Arrangement: the reductions to condensed form are the unblocked launch sequences of cplx_condensed_core.cuh (flat maps, one thread
// per row / element; the same source runs thread by thread on the host in tests/test_cplx_condensed_emul_cpu.py). The condensed
// REAL problems go to the solvers the real entry points use (tridiag_dc.cu; the Golub-Kahan / QR-stabilised bidiagonal solver of
// svd_vectors.cu), the back-transforms to the complex block-Householder sequence of cplx.cu with blocks of 32 reflectors (T blocks
// from the BuildTBlocks body). Functional, not tuned: O(n) launches per column and n-thread matrix-vector products — meant for the moderate
// sizes complex users of the C ABI bring, not for the BASELINE sizes of the real path.
// The values-only calls (U / V not wanted) run the same path and skip the back-transforms.
#include <algorithm>

#include "cplx_condensed_core.cuh"
#include "flat_map.cuh"
#include "runtime.cuh"
#include "tensor_ops.cuh"

namespace fb {

namespace {

struct DevWork {
  cc::Work ws;
  void* blocks[5];
  explicit DevWork(i64 len) {
    ws.v = (cc::Cx*)(blocks[0] = ws_alloc((size_t)(len + 1) * sizeof(cc::Cx)));
    ws.p = (cc::Cx*)(blocks[1] = ws_alloc((size_t)(len + 1) * sizeof(cc::Cx)));
    ws.w = (cc::Cx*)(blocks[2] = ws_alloc((size_t)(len + 1) * sizeof(cc::Cx)));
    ws.part = (double*)(blocks[3] = ws_alloc((size_t)3 * cc::NP * sizeof(double)));
    ws.sc = (double*)(blocks[4] = ws_alloc((size_t)cc::SC_COUNT * sizeof(double)));
  }
  void release() {
    for (void* b : blocks) ws_free(b);
  }
};

// block size of the back-transforms and the T blocks of a reflector set (householder.rs:132-272): V = the m x s unit-lower trapezoid at
// `V` (leading dimension ld), tau its s scalars; returns a zero-initialised bs x s factor (pool block) with the blocks filled in
constexpr i64 BT_BS = 32;
inline cc::Cx* build_t_blocks(cudaStream_t st, const cc::Cx* V, i64 ld, i64 m, i64 s, const double* tau, i64 bs) {
  cc::Cx* Tf = (cc::Cx*)ws_alloc((size_t)bs * (size_t)s * sizeof(cc::Cx));
  FB_CUDA_CHECK(cudaMemsetAsync(Tf, 0, (size_t)bs * (size_t)s * sizeof(cc::Cx), st));
  DevRun run{st};
  run(cc::BuildTBlocks{V, ld, m, s, tau, Tf, bs, bs}, bs, s);
  return Tf;
}

// c64 view of the input: the view itself (TO = double) or a compact widened copy (TO = float; *owned receives the block)
inline View<const double> as_c64(cudaStream_t st, View<const double> A, void** owned) {
  (void)st;
  *owned = nullptr;
  return A;
}
inline View<const double> as_c64(cudaStream_t st, View<const float> A, void** owned) {
  const i64 m = A.nrows, n = A.ncols;
  cc::Cx* W = (cc::Cx*)ws_alloc((size_t)std::max<i64>(1, m * n) * sizeof(cc::Cx));
  DevRun run{st};
  run(cc::WidenC32{A.ptr, A.rs, A.cs, W, m, m, n}, m, n);
  *owned = W;
  return View<const double>{(const double*)W, m, n, 1, m};
}

}  // namespace

template <class TO>
bool self_adjoint_evd_cx(cudaStream_t st, View<const TO> A_in, View<TO> U, TO* S, i64 sstride) {
  const i64 n = A_in.nrows;
  FB_ASSERT(A_in.ncols == n, "self_adjoint_evd: square matrix required");
  if (n == 0) return true;
  const bool want_u = U.ptr != nullptr && U.ncols > 0;
  DevRun run{st};
  void* owned = nullptr;
  const View<const double> A = as_c64(st, A_in, &owned);
  cc::Cx* W = (cc::Cx*)ws_alloc((size_t)n * (size_t)n * sizeof(cc::Cx));
  run(cc::BuildHermitian{A.ptr, A.rs, A.cs, W, n, n}, n, n);
  DevWork work(n);
  // tau | d | e | lam (n doubles each), then ph | tauc (n complex each)
  double* reals = (double*)ws_alloc((size_t)(4 * n + 4) * sizeof(double));
  double *tau = reals, *d = reals + n, *e = reals + 2 * n, *lam = reals + 3 * n;
  cc::Cx* cplx = (cc::Cx*)ws_alloc((size_t)(2 * n + 2) * sizeof(cc::Cx));
  cc::Cx *ph = cplx, *tauc = cplx + n;
  FB_CUDA_CHECK(cudaMemsetAsync(reals, 0, (size_t)(4 * n + 4) * sizeof(double), st));
  cc::tridiag_unblocked(run, W, n, n, tau, work.ws);
  run(cc::TridiagPhases{W, n, n, tau, d, e, ph, tauc}, 1, 1);
  bool ok = device_all_finite<double>(st, d, n) && (n < 2 || device_all_finite<double>(st, e, n - 1));
  double* Q = nullptr;
  if (ok) {
    Q = (double*)ws_alloc((size_t)n * (size_t)n * sizeof(double));
    ok = tridiag_dc_f64(st, d, e, n, lam, Q, n);
  }
  if (ok) {
    if (want_u) {
      FB_ASSERT(U.nrows == n && U.ncols == n, "self_adjoint_evd: U must be n x n");
      cc::Cx* Uw = (cc::Cx*)ws_alloc((size_t)n * (size_t)n * sizeof(cc::Cx));
      run(cc::ScaleRowsEmbed{Q, n, n, ph, Uw, n, n, n}, n, n);  // diag(ph) Q
      cc::Cx* Tf = nullptr;
      if (n > 1) {  // rows 1.. <- H_0 H_1 ... H_{n-2} rows 1.. (evd/mod.rs:411-418); reflector k: column k of W below the subdiagonal
        const i64 bs = std::min<i64>(BT_BS, n - 1);
        Tf = build_t_blocks(st, W + 1, n, n - 1, n - 1, tau, bs);
        apply_householder_sequence_left_c64(st, VCD{(const double*)(W + 1), n - 1, n - 1, 1, n}, VCD{(const double*)Tf, bs, n - 1, 1, bs},
                                            false, VD{(double*)(Uw + 1), n - 1, n, 1, n}, false);
      }
      run(cc::CopyOut<TO>{U.ptr, U.rs, U.cs, Uw, n, n, n}, n, n);
      FB_CUDA_CHECK(cudaStreamSynchronize(st));
      if (Tf) ws_free(Tf);
      ws_free(Uw);
    }
    run(cc::CopyValues<TO>{S, sstride, lam, n}, n, 1);
  }
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (Q) ws_free(Q);
  ws_free(cplx);
  ws_free(reals);
  work.release();
  ws_free(W);
  if (owned) ws_free(owned);
  return ok;
}

template <class TO>
bool svd_cx(cudaStream_t st, View<const TO> A_in, View<TO> U, TO* S, i64 sstride, View<TO> V) {
  const bool transpose = A_in.ncols > A_in.nrows;
  const i64 m = transpose ? A_in.ncols : A_in.nrows, n = transpose ? A_in.nrows : A_in.ncols;  // the work matrix M (= A or A^H) is m x n
  // A = U S V^H; for a wide A: A^H = U' S V'^H, so U = V' and V = U' (svd/mod.rs:560-583, 661-669)
  View<TO> Um = transpose ? V : U, Vm = transpose ? U : V;
  const bool want_u = Um.ptr != nullptr && Um.ncols > 0, want_v = Vm.ptr != nullptr && Vm.ncols > 0;
  DevRun run{st};
  if (n == 0) {
    if (want_u) {  // no singular values: the full left factor is the identity
      const i64 ku = Um.ncols;
      cc::Cx* Uw = (cc::Cx*)ws_alloc((size_t)std::max<i64>(1, m * ku) * sizeof(cc::Cx));
      run(cc::ScaleRowsEmbed{nullptr, 1, 0, nullptr, Uw, m, m, ku}, m, ku);
      run(cc::CopyOut<TO>{Um.ptr, Um.rs, Um.cs, Uw, m, m, ku}, m, ku);
      FB_CUDA_CHECK(cudaStreamSynchronize(st));
      ws_free(Uw);
    }
    return true;
  }
  void* owned = nullptr;
  const View<const double> A = as_c64(st, A_in, &owned);
  cc::Cx* W = (cc::Cx*)ws_alloc((size_t)m * (size_t)n * sizeof(cc::Cx));
  run(cc::CopyIn{A.ptr, A.rs, A.cs, W, m, m, n, transpose ? 1 : 0}, m, n);
  DevWork work(m);
  // tl | tr | d | f | s_sorted (n doubles each), then l | r | tlc | trc (n complex each)
  double* reals = (double*)ws_alloc((size_t)(5 * n + 4) * sizeof(double));
  double *tl = reals, *tr = reals + n, *d = reals + 2 * n, *f = reals + 3 * n, *s_sorted = reals + 4 * n;
  cc::Cx* cplx = (cc::Cx*)ws_alloc((size_t)(4 * n + 4) * sizeof(cc::Cx));
  cc::Cx *l = cplx, *r = cplx + n, *tlc = cplx + 2 * n, *trc = cplx + 3 * n;
  FB_CUDA_CHECK(cudaMemsetAsync(reals, 0, (size_t)(5 * n + 4) * sizeof(double), st));
  cc::bidiag_unblocked(run, W, m, m, n, tl, tr, work.ws);
  run(cc::BidiagPhases{W, m, n, tl, tr, d, f, l, r, tlc, trc}, 1, 1);
  bool ok = device_all_finite<double>(st, d, n) && (n < 2 || device_all_finite<double>(st, f, n - 1));
  double *UB = nullptr, *VB = nullptr;
  if (ok) {
    UB = (double*)ws_alloc((size_t)n * (size_t)n * sizeof(double));
    VB = (double*)ws_alloc((size_t)n * (size_t)n * sizeof(double));
    ok = bidiag_svd_vectors_f64(st, d, f, n, s_sorted, UB, VB);  // B_real = UB diag(S) VB^T
  }
  if (ok && want_u) {
    const i64 ku = Um.ncols;  // n (thin) or m (full)
    FB_ASSERT(Um.nrows == m && (ku == n || ku == m), "svd: the left factor must be nrows x {size, nrows}");
    cc::Cx* Uw = (cc::Cx*)ws_alloc((size_t)m * (size_t)ku * sizeof(cc::Cx));
    run(cc::ScaleRowsEmbed{UB, n, n, l, Uw, m, m, ku}, m, ku);  // [diag(l) UB, 0; 0, I]
    // U = H_0 ... H_{n-1} [.]: left reflector k is column k of W below the diagonal (svd/mod.rs:403-412, Conj::No)
    const i64 bs = std::min<i64>(BT_BS, n);
    cc::Cx* Tf = build_t_blocks(st, W, m, m, n, tl, bs);
    apply_householder_sequence_left_c64(st, VCD{(const double*)W, m, n, 1, m}, VCD{(const double*)Tf, bs, n, 1, bs}, false,
                                        VD{(double*)Uw, m, ku, 1, m}, false);
    run(cc::CopyOut<TO>{Um.ptr, Um.rs, Um.cs, Uw, m, m, ku}, m, ku);
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    ws_free(Tf);
    ws_free(Uw);
  }
  if (ok && want_v) {
    FB_ASSERT(Vm.nrows == n && Vm.ncols == n, "svd: the right factor must be ncols x ncols");
    cc::Cx* Vw = (cc::Cx*)ws_alloc((size_t)n * (size_t)n * sizeof(cc::Cx));
    run(cc::ScaleRowsEmbed{VB, n, n, r, Vw, n, n, n}, n, n);  // diag(r) VB
    cc::Cx *T = nullptr, *Tfr = nullptr;
    if (n > 1) {
      // the right reflectors are the rows of W right of the superdiagonal: transposed copy, sequence with Conj::Yes on rows 1..
      // (svd/mod.rs:413-428)
      T = (cc::Cx*)ws_alloc((size_t)n * (size_t)n * sizeof(cc::Cx));
      run(cc::TransposeCorner{W, m, T, n, n}, n, n);
      const i64 bs = std::min<i64>(BT_BS, n - 1);
      Tfr = build_t_blocks(st, T + 1, n, n - 1, n - 1, tr, bs);
      apply_householder_sequence_left_c64(st, VCD{(const double*)(T + 1), n - 1, n - 1, 1, n}, VCD{(const double*)Tfr, bs, n - 1, 1, bs}, true,
                                          VD{(double*)(Vw + 1), n - 1, n, 1, n}, false);
    }
    run(cc::CopyOut<TO>{Vm.ptr, Vm.rs, Vm.cs, Vw, n, n, n}, n, n);
    FB_CUDA_CHECK(cudaStreamSynchronize(st));
    if (Tfr) ws_free(Tfr);
    if (T) ws_free(T);
    ws_free(Vw);
  }
  if (ok) run(cc::CopyValues<TO>{S, sstride, s_sorted, n}, n, 1);
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (VB) ws_free(VB);
  if (UB) ws_free(UB);
  ws_free(cplx);
  ws_free(reals);
  work.release();
  ws_free(W);
  if (owned) ws_free(owned);
  return ok;
}

// svd::bidiag::bidiag_in_place (svd/bidiag.rs:47-256) and evd::tridiag::tridiag_in_place (evd/tridiag.rs:274-529) for complex T as
// extensions (the real dtypes have their HBM-bound kernels in bidiag.cu / tridiag.cu): the unblocked sequences of
// cplx_condensed_core.cuh and the T blocks of the reflectors in the layout the reference returns.
template <class TO>
void bidiag_in_place_cx(cudaStream_t st, View<TO> A, View<TO> Hl, View<TO> Hr) {
  const i64 m = A.nrows, n = A.ncols, bl = Hl.nrows, br = Hr.nrows;
  FB_ASSERT(m >= n, "bidiag_in_place: nrows >= ncols required (the SVD driver works on the adjoint of wide inputs)");
  FB_ASSERT(Hl.ncols == n && Hr.ncols == (n > 0 ? n - 1 : 0) && (n == 0 || bl > 0) && (n <= 1 || br > 0), "bidiag_in_place: Householder factor shapes");
  if (n == 0) return;
  DevRun run{st};
  void* owned = nullptr;
  const View<const double> Ac = as_c64(st, View<const TO>{A.ptr, m, n, A.rs, A.cs}, &owned);
  cc::Cx* W = (cc::Cx*)ws_alloc((size_t)m * (size_t)n * sizeof(cc::Cx));
  run(cc::CopyIn{Ac.ptr, Ac.rs, Ac.cs, W, m, m, n, 0}, m, n);
  DevWork work(m);
  double* taus = (double*)ws_alloc((size_t)(2 * n + 2) * sizeof(double));
  double *tl = taus, *tr = taus + n;
  FB_CUDA_CHECK(cudaMemsetAsync(taus, 0, (size_t)(2 * n + 2) * sizeof(double), st));
  cc::bidiag_unblocked(run, W, m, m, n, tl, tr, work.ws);
  cc::Cx* Tl = build_t_blocks(st, W, m, m, n, tl, bl);
  cc::Cx *T = nullptr, *Tr = nullptr;
  if (n > 1) {  // the right reflectors as columns (svd/bidiag.rs:239-255: the factor of the transposed rows)
    T = (cc::Cx*)ws_alloc((size_t)n * (size_t)n * sizeof(cc::Cx));
    run(cc::TransposeCorner{W, m, T, n, n}, n, n);
    Tr = build_t_blocks(st, T + 1, n, n - 1, n - 1, tr, br);
  }
  run(cc::CopyOut<TO>{A.ptr, A.rs, A.cs, W, m, m, n}, m, n);
  run(cc::CopyOut<TO>{Hl.ptr, Hl.rs, Hl.cs, Tl, bl, bl, n}, bl, n);
  if (n > 1) run(cc::CopyOut<TO>{Hr.ptr, Hr.rs, Hr.cs, Tr, br, br, n - 1}, br, n - 1);
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (Tr) ws_free(Tr);
  if (T) ws_free(T);
  ws_free(Tl);
  ws_free(taus);
  work.release();
  ws_free(W);
  if (owned) ws_free(owned);
}
template <class TO>
void tridiag_in_place_cx(cudaStream_t st, View<TO> A, View<TO> H) {
  const i64 n = A.nrows, b = H.nrows;
  FB_ASSERT(A.ncols == n && H.ncols == (n > 0 ? n - 1 : 0) && (n <= 1 || b > 0), "tridiag_in_place: square A, householder factor b x (n - 1)");
  if (n == 0) return;
  DevRun run{st};
  void* owned = nullptr;
  const View<const double> Ac = as_c64(st, View<const TO>{A.ptr, n, n, A.rs, A.cs}, &owned);
  cc::Cx* W = (cc::Cx*)ws_alloc((size_t)n * (size_t)n * sizeof(cc::Cx));
  run(cc::BuildHermitian{Ac.ptr, Ac.rs, Ac.cs, W, n, n}, n, n);
  DevWork work(n);
  double* tau = (double*)ws_alloc((size_t)(n + 1) * sizeof(double));
  FB_CUDA_CHECK(cudaMemsetAsync(tau, 0, (size_t)(n + 1) * sizeof(double), st));
  cc::tridiag_unblocked(run, W, n, n, tau, work.ws);
  cc::Cx* Tf = n > 1 ? build_t_blocks(st, W + 1, n, n - 1, n - 1, tau, b) : nullptr;
  run(cc::CopyOutLower<TO>{A.ptr, A.rs, A.cs, W, n, n}, n, n);  // only the lower triangle is written (tridiag.rs:274-280)
  if (n > 1) run(cc::CopyOut<TO>{H.ptr, H.rs, H.cs, Tf, b, b, n - 1}, b, n - 1);
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (Tf) ws_free(Tf);
  ws_free(tau);
  work.release();
  ws_free(W);
  if (owned) ws_free(owned);
}
template void bidiag_in_place_cx<double>(cudaStream_t, View<double>, View<double>, View<double>);
template void bidiag_in_place_cx<float>(cudaStream_t, View<float>, View<float>, View<float>);
template void tridiag_in_place_cx<double>(cudaStream_t, View<double>, View<double>);
template void tridiag_in_place_cx<float>(cudaStream_t, View<float>, View<float>);

// evd::hessenberg::hessenberg_in_place (evd/hessenberg.rs:549-567) as an extension (the reference does not export it through its C
// ABI): A <- its upper Hessenberg form H = Q^H A Q in the entries (i, j) with i <= j + 1 and the reflectors of Q = H_0 ... H_{n-2} below
// the subdiagonal; Hf (bs x (n - 1)) <- their T blocks (diag tau, V^H V above it inside each block), the layout the block-Householder
// sequences take. Every dtype runs the c64 launch sequence of cplx_condensed_core.cuh (real input as (x, 0): stays exactly real).
template <class TO, bool CX>
void hessenberg_in_place_t(cudaStream_t st, View<TO> A, View<TO> Hf) {
  const i64 n = A.nrows, bs = Hf.nrows, s = n > 0 ? n - 1 : 0;
  FB_ASSERT(A.ncols == n && Hf.ncols == s && (s == 0 || bs > 0), "hessenberg_in_place: square A, householder factor bs x (n - 1)");
  if (n == 0) return;
  DevRun run{st};
  cc::Cx* W = (cc::Cx*)ws_alloc((size_t)n * (size_t)n * sizeof(cc::Cx));
  if (CX) {
    void* owned = nullptr;
    const View<const double> Ac = as_c64(st, View<const TO>{A.ptr, n, n, A.rs, A.cs}, &owned);
    run(cc::CopyIn{Ac.ptr, Ac.rs, Ac.cs, W, n, n, n, 0}, n, n);
    if (owned) {
      FB_CUDA_CHECK(cudaStreamSynchronize(st));
      ws_free(owned);
    }
  } else {
    run(cc::RealToCx<TO>{A.ptr, A.rs, A.cs, W, n, n, n}, n, n);
  }
  DevWork work(n);
  double* tau = (double*)ws_alloc((size_t)(n + 1) * sizeof(double));
  FB_CUDA_CHECK(cudaMemsetAsync(tau, 0, (size_t)(n + 1) * sizeof(double), st));
  cc::hessenberg_unblocked(run, W, n, n, tau, work.ws);
  cc::Cx* Tf = nullptr;
  if (s > 0) {
    Tf = (cc::Cx*)ws_alloc((size_t)bs * (size_t)s * sizeof(cc::Cx));
    FB_CUDA_CHECK(cudaMemsetAsync(Tf, 0, (size_t)bs * (size_t)s * sizeof(cc::Cx), st));
    run(cc::BuildTBlocks{W + 1, n, n - 1, s, tau, Tf, bs, bs}, bs, s);
  }
  if (CX) {
    run(cc::CopyOut<TO>{A.ptr, A.rs, A.cs, W, n, n, n}, n, n);
    if (s > 0) run(cc::CopyOut<TO>{Hf.ptr, Hf.rs, Hf.cs, Tf, bs, bs, s}, bs, s);
  } else {
    run(cc::CxToReal<TO>{A.ptr, A.rs, A.cs, W, n, n, n}, n, n);
    if (s > 0) run(cc::CxToReal<TO>{Hf.ptr, Hf.rs, Hf.cs, Tf, bs, bs, s}, bs, s);
  }
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (Tf) ws_free(Tf);
  ws_free(tau);
  work.release();
  ws_free(W);
}
template void hessenberg_in_place_t<double, false>(cudaStream_t, View<double>, View<double>);
template void hessenberg_in_place_t<float, false>(cudaStream_t, View<float>, View<float>);
template void hessenberg_in_place_t<double, true>(cudaStream_t, View<double>, View<double>);
template void hessenberg_in_place_t<float, true>(cudaStream_t, View<float>, View<float>);

template bool svd_cx<double>(cudaStream_t, View<const double>, View<double>, double*, i64, View<double>);
template bool svd_cx<float>(cudaStream_t, View<const float>, View<float>, float*, i64, View<float>);
template bool self_adjoint_evd_cx<double>(cudaStream_t, View<const double>, View<double>, double*, i64);
template bool self_adjoint_evd_cx<float>(cudaStream_t, View<const float>, View<float>, float*, i64);

}  // namespace fb
