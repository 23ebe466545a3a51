// c64 (complex<f64>) GEMM / structured GEMM on the f64 DMMA kernel ("4M" formulation).
//
// Reference semantics: faer/src/linalg/matmul/mod.rs:1711-1749 (`matmul_with_conj`: dst = [dst +] alpha *
// conj?(lhs) * conj?(rhs)), triangular.rs:1079-1126 for the structured variant.
//
// A c64 matrix is an interleaved (re, im) array, i.e. two f64 matrices with doubled element strides. G1 takes views
// with arbitrary strides, so the complex product is issued as (up to) eight real DMMA GEMMs on the real / imaginary
// planes — no copies, no de-interleaving pass:
//     P_re = A_re B_re - sa sb A_im B_im          sa = -1 if conj(lhs) else +1, sb likewise
//     P_im = sb A_re B_im + sa A_im B_re
//     dst_re = [dst_re +] ar P_re - ai P_im,   dst_im = [dst_im +] ai P_re + ar P_im        (alpha = ar + i ai)
// Real alpha (the factorizations only use +-1) needs four launches. Unit-triangular operands contribute 1 on the
// real plane and 0 on the imaginary plane (UNIT -> STRICT for the imaginary view).
// Large unstructured products first split both operands into planar (real / imaginary) column-major copies — an O(n^2) pass,
// 0.7 ms at n = 8192 against 120 ms of products — because unit-stride planes are what TMA can read: the four (real alpha)
// products then run on the TMA-fed warp-specialised kernel (gemm_f64_ws.cuh, 36 TFLOP/s) instead of the cp.async kernel on
// stride-2 views (32 TFLOP/s, and twice the L2 -> SM operand traffic); the destination stays interleaved (the kernel's
// epilogue takes any strides). Structured operands and small products keep the no-copy path.
#include "linalg_f64.cuh"
#include "runtime.cuh"

namespace fb {

namespace {
inline int imag_struct(int s) {
  if (s == UNIT_LOWER) return STRICT_LOWER;
  if (s == UNIT_UPPER) return STRICT_UPPER;
  return s;
}
// planar copies of an interleaved complex matrix (any strides, in complex units): re / im column-major with leading dimension ld
__global__ void c64_split_kernel(const double* __restrict__ src, i64 rs, i64 cs, i64 rows, i64 cols, double* __restrict__ re,
                                 double* __restrict__ im, i64 ld) {
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  const i64 i = e % rows, j = e / rows;
  const double2 v = *reinterpret_cast<const double2*>(src + 2 * (i * rs + j * cs));
  re[j * ld + i] = v.x;
  im[j * ld + i] = v.y;
}
}  // namespace

// views are given in COMPLEX element units (ptr to the first complex element, strides in complex elements)
void gemm_c64(cudaStream_t stream, VD dst, int dst_struct, int accum, VCD lhs, int lhs_struct, bool conj_lhs, VCD rhs,
              int rhs_struct, bool conj_rhs, double alpha_re, double alpha_im) {
  auto re = [](auto v) { v.rs *= 2; v.cs *= 2; return v; };
  auto im = [](auto v) { v.ptr += 1; v.rs *= 2; v.cs *= 2; return v; };
  const double sa = conj_lhs ? -1.0 : 1.0, sb = conj_rhs ? -1.0 : 1.0;
  // planar operand copies for large unstructured products (see the header)
  const i64 m = dst.nrows, n = dst.ncols, k = lhs.ncols;
  const bool planar = lhs_struct == RECT && rhs_struct == RECT && m >= 512 && n >= 512 && k >= 256 &&
                      (((uintptr_t)lhs.ptr | (uintptr_t)rhs.ptr) & 15) == 0;
  double* planes = nullptr;
  VCD a_pl[2], b_pl[2];
  if (planar) {
    const i64 lda = (m + 1) & ~(i64)1, ldb = (k + 1) & ~(i64)1;
    planes = (double*)ws_alloc((size_t)(2 * lda * k + 2 * ldb * n) * sizeof(double));
    double* ar = planes;
    double* ai = ar + lda * k;
    double* br = ai + lda * k;
    double* bi = br + ldb * n;
    c64_split_kernel<<<(unsigned)((m * k + 255) / 256), 256, 0, stream>>>(lhs.ptr, lhs.rs, lhs.cs, m, k, ar, ai, lda);
    c64_split_kernel<<<(unsigned)((k * n + 255) / 256), 256, 0, stream>>>(rhs.ptr, rhs.rs, rhs.cs, k, n, br, bi, ldb);
    FB_CUDA_CHECK(cudaGetLastError());
    note_launch();
    note_launch();
    a_pl[0] = VCD{ar, m, k, 1, lda};
    a_pl[1] = VCD{ai, m, k, 1, lda};
    b_pl[0] = VCD{br, k, n, 1, ldb};
    b_pl[1] = VCD{bi, k, n, 1, ldb};
  }
  struct Term { int a_im, b_im; double c_re, c_im; };
  const Term terms[4] = {
      {0, 0, alpha_re, alpha_im},
      {1, 1, -alpha_re * sa * sb, -alpha_im * sa * sb},
      {0, 1, -alpha_im * sb, alpha_re * sb},
      {1, 0, -alpha_im * sa, alpha_re * sa},
  };
  bool first_re = true, first_im = true;
  for (const Term& t : terms) {
    VCD a = planar ? a_pl[t.a_im] : (t.a_im ? im(lhs) : re(lhs));
    VCD b = planar ? b_pl[t.b_im] : (t.b_im ? im(rhs) : re(rhs));
    const int as = t.a_im ? imag_struct(lhs_struct) : lhs_struct;
    const int bs = t.b_im ? imag_struct(rhs_struct) : rhs_struct;
    if (t.c_re != 0.0) {
      gemm_f64(stream, re(dst), dst_struct, first_re ? accum : 1, a, as, b, bs, t.c_re);
      first_re = false;
    }
    if (t.c_im != 0.0) {
      gemm_f64(stream, im(dst), dst_struct, first_im ? accum : 1, a, as, b, bs, t.c_im);
      first_im = false;
    }
  }
  // planes that received no term: Replace still has to define them (alpha == 0 or purely real/imaginary products)
  if (accum == 0) {
    if (first_re) gemm_f64(stream, re(dst), dst_struct, 0, re(lhs), lhs_struct, re(rhs), rhs_struct, 0.0);
    if (first_im) gemm_f64(stream, im(dst), dst_struct, 0, re(lhs), lhs_struct, re(rhs), rhs_struct, 0.0);
  }
  if (planes) {
    // the pool is not stream-ordered: the block may only be handed out again once the products have read it
    FB_CUDA_CHECK(cudaStreamSynchronize(stream));
    ws_free(planes);
  }
}

}  // namespace fb
