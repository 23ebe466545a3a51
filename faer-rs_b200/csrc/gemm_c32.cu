// c32 (complex<f32>) GEMM / structured GEMM on the f32 kernels ("4M" formulation), the single-precision twin of
// gemm_c64.cu.
//
// Reference semantics: faer/src/linalg/matmul/mod.rs:1711-1749 (`matmul_with_conj`), triangular.rs:1079-1126.
//
// A c32 matrix is an interleaved (re, im) array, i.e. two f32 matrices with doubled element strides; the complex
// product is issued as (up to) eight real f32 GEMMs on the real / imaginary planes:
//     P_re = A_re B_re - sa sb A_im B_im,   P_im = sb A_re B_im + sa A_im B_re      (sa / sb = -1 for a conjugated operand)
//     dst_re = [dst_re +] ar P_re - ai P_im,   dst_im = [dst_im +] ai P_re + ar P_im
// Large unstructured products therefore run on the tcgen05 kernel (its packing pass absorbs the stride-2 planes), the
// rest on the mma.sync 3xTF32 kernel; both give fp32-class accuracy. Unit-triangular operands contribute 1 on the real
// plane and 0 on the imaginary plane (UNIT -> STRICT for the imaginary view).
#include "gemm_f32.cuh"

namespace fb {

namespace {
inline int c32_imag_struct(int s) {
  if (s == UNIT_LOWER) return STRICT_LOWER;
  if (s == UNIT_UPPER) return STRICT_UPPER;
  return s;
}
}  // namespace

// views are given in COMPLEX element units (ptr to the first complex element, strides in complex elements)
void gemm_c32(cudaStream_t stream, VF dst, int dst_struct, int accum, VCF lhs, int lhs_struct, bool conj_lhs, VCF rhs,
              int rhs_struct, bool conj_rhs, float alpha_re, float alpha_im) {
  auto re = [](auto v) { v.rs *= 2; v.cs *= 2; return v; };
  auto im = [](auto v) { v.ptr += 1; v.rs *= 2; v.cs *= 2; return v; };
  const float sa = conj_lhs ? -1.0f : 1.0f, sb = conj_rhs ? -1.0f : 1.0f;
  struct Term { int a_im, b_im; float c_re, c_im; };
  const Term terms[4] = {
      {0, 0, alpha_re, alpha_im},
      {1, 1, -alpha_re * sa * sb, -alpha_im * sa * sb},
      {0, 1, -alpha_im * sb, alpha_re * sb},
      {1, 0, -alpha_im * sa, alpha_re * sa},
  };
  bool first_re = true, first_im = true;
  for (const Term& t : terms) {
    VCF a = t.a_im ? im(lhs) : re(lhs);
    VCF b = t.b_im ? im(rhs) : re(rhs);
    const int as = t.a_im ? c32_imag_struct(lhs_struct) : lhs_struct;
    const int bs = t.b_im ? c32_imag_struct(rhs_struct) : rhs_struct;
    if (t.c_re != 0.0f) {
      gemm_f32(stream, re(dst), dst_struct, first_re ? accum : 1, a, as, b, bs, t.c_re);
      first_re = false;
    }
    if (t.c_im != 0.0f) {
      gemm_f32(stream, im(dst), dst_struct, first_im ? accum : 1, a, as, b, bs, t.c_im);
      first_im = false;
    }
  }
  // planes that received no term: Replace still has to define them
  if (accum == 0) {
    if (first_re) gemm_f32(stream, re(dst), dst_struct, 0, re(lhs), lhs_struct, re(rhs), rhs_struct, 0.0f);
    if (first_im) gemm_f32(stream, im(dst), dst_struct, 0, re(lhs), lhs_struct, re(rhs), rhs_struct, 0.0f);
  }
}

}  // namespace fb
