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
// A fused complex kernel (one pass over the operands, 4 DMMA per fragment pair) is the next step for this path; the
// strided planes cost 2x the L2->SM operand traffic, which the f64 tensor roofline leaves room for.
#include "linalg_f64.cuh"

namespace fb {

namespace {
inline int imag_struct(int s) {
  if (s == UNIT_LOWER) return STRICT_LOWER;
  if (s == UNIT_UPPER) return STRICT_UPPER;
  return s;
}
}  // namespace

// views are given in COMPLEX element units (ptr to the first complex element, strides in complex elements)
void gemm_c64(cudaStream_t stream, VD dst, int dst_struct, int accum, VCD lhs, int lhs_struct, bool conj_lhs, VCD rhs,
              int rhs_struct, bool conj_rhs, double alpha_re, double alpha_im) {
  auto re = [](auto v) { v.rs *= 2; v.cs *= 2; return v; };
  auto im = [](auto v) { v.ptr += 1; v.rs *= 2; v.cs *= 2; return v; };
  const double sa = conj_lhs ? -1.0 : 1.0, sb = conj_rhs ? -1.0 : 1.0;
  struct Term { int a_im, b_im; double c_re, c_im; };
  const Term terms[4] = {
      {0, 0, alpha_re, alpha_im},
      {1, 1, -alpha_re * sa * sb, -alpha_im * sa * sb},
      {0, 1, -alpha_im * sb, alpha_re * sb},
      {1, 0, -alpha_im * sa, alpha_re * sa},
  };
  bool first_re = true, first_im = true;
  for (const Term& t : terms) {
    VCD a = t.a_im ? im(lhs) : re(lhs);
    VCD b = t.b_im ? im(rhs) : re(rhs);
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
}

}  // namespace fb
