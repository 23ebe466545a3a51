// f32 GEMM on the 5th-generation tensor cores: error-compensated 3xTF32 with tcgen05.mma (kind::tf32), operands staged
// by TMA into 128B-swizzled shared memory, accumulator in TMEM.
//
// Replaces the mma.sync 3xTF32 kernel (gemm_f32.cu) for large products of `linalg::matmul::matmul` with T = f32
// (reference faer/src/linalg/matmul/mod.rs:1617-1660; the block-Householder applies of the f32 QR are its main user).
//
// Scheme. Every fp32 operand value is split once, a = a_hi + a_lo with a_hi = tf32(a), a_lo = tf32(a - a_hi), by a
// packing pass that also brings BOTH operands into the one layout the MMA kernel consumes: row-major [rows][Kp] with the
// contraction index contiguous ("K-major"), Kp = k rounded up to 32 and zero-padded. The pass reads arbitrary element
// strides (transposes, negative strides), so the MMA kernel has a single code path. Per 32-wide k-block and tile the
// kernel issues, for each of the four 8-deep UMMA steps, a_lo*b_hi, a_hi*b_lo, a_hi*b_hi (small terms first) into the
// same fp32 TMEM accumulator: fp32-class accuracy (the dropped a_lo*b_lo term is ~2^-22 relative).
//
// Kernel anatomy (one 128 x 128 output tile per CTA, 256 threads):
//   warp 0   TMA producer: 4 loads per stage (A_hi, A_lo, B_hi, B_lo: 4 x 16 KB), mbarrier expect_tx
//   warp 1   MMA issuer (one elected lane): waits `full`, 12 tcgen05.mma per stage, tcgen05.commit -> `empty`
//   warp 2   TMEM allocation (128 columns) / deallocation
//   all 8 warps, once their role loops are done: epilogue — tcgen05.ld 32x32b.x32 (warp w: TMEM lane quarter w % 4,
//            column half w / 4), alpha / accumulate with batched loads, coalesced stores with the destination's strides
// 3 stages x 64 KB of shared memory. Out-of-range rows / columns are zero-filled by TMA and masked in the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstdlib>

namespace fb {
namespace tc {

struct Operand {
  const float* p;
  int rows, cols;
  long long rs, cs;  // element strides
};

struct Workspace {
  float* buf = nullptr;
  size_t bytes = 0;
  float last_mma_ms = 0.f;
};

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;         // 16 KB (BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;     // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B, rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) /*LBO (ignored for swizzled K-major)*/ |
         (64ull << 32) /*SBO = 1024 B*/ | (1ull << 46) /*descriptor version (sm_100)*/ | (2ull << 61) /*SWIZZLE_128B*/;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__global__ void __launch_bounds__(256, 1)
gemm_f32_tc_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                   const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, float* __restrict__ C,
                   long long c_rs, long long c_cs, int m, int n, int kblocks_total, int kb_per_split, long long c_split_stride,
                   float alpha, int accum) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grouped rasterisation: the ~148 co-resident CTAs cover 16 m-tiles x ~9 n-tiles, so both operands are shared through
  // L2 (with the plain x-fastest order every wave streamed 64 different A tiles: 15.8 GB of DRAM reads for 1.07 GB of
  // operands at n = 8192, profiles/r01_tc_gemm8192_ncu.txt)
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  constexpr int GROUP_M = 16;
  const int pid = blockIdx.x, width = GROUP_M * tiles_n;
  const int first_m = (pid / width) * GROUP_M;
  const int gsize = min(tiles_m - first_m, GROUP_M);
  const int m0 = (first_m + (pid % width) % gsize) * BM, n0 = ((pid % width) / gsize) * BN;
  // split-K: grid.z slices of the k-blocks, each writing its own partial product
  const int kb0 = blockIdx.z * kb_per_split;
  const int kblocks = min(kb_per_split, kblocks_total - kb0);
  C += (long long)blockIdx.z * c_split_stride;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        tma_load_2d(st + 0 * TILE_BYTES, &mAh, &full[s], (kb0 + kb) * BK, m0);
        tma_load_2d(st + 1 * TILE_BYTES, &mAl, &full[s], (kb0 + kb) * BK, m0);
        tma_load_2d(st + 2 * TILE_BYTES, &mBh, &full[s], (kb0 + kb) * BK, n0);
        tma_load_2d(st + 3 * TILE_BYTES, &mBl, &full[s], (kb0 + kb) * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint64_t a_hi = umma_desc_k_sw128(base + 0 * TILE_BYTES + kk * 32);
          const uint64_t a_lo = umma_desc_k_sw128(base + 1 * TILE_BYTES + kk * 32);
          const uint64_t b_hi = umma_desc_k_sw128(base + 2 * TILE_BYTES + kk * 32);
          const uint64_t b_lo = umma_desc_k_sw128(base + 3 * TILE_BYTES + kk * 32);
          umma_tf32(tmem_base, a_lo, b_hi, IDESC, (kb | kk) != 0 ? 1u : 0u);
          umma_tf32(tmem_base, a_hi, b_lo, IDESC, 1u);
          umma_tf32(tmem_base, a_hi, b_hi, IDESC, 1u);
        }
        umma_commit(&empty[s]);  // the slot is free once these MMAs have read it
      }
      umma_commit(tmem_full);    // accumulator complete
    }
  }
  __syncwarp();
  // ---- epilogue, all 8 warps: warp w reads TMEM lane quarter w % 4 (the only one it may access) and the column half
  // w / 4; the role warps join once their loops are done (short-k products are epilogue-bound) ----
  {
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + 32 * q + lane;
    const int cbeg = (warp >> 2) * (BN / 2);
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + BN / 2; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < m) {
        float* dst = C + (long long)row * c_rs + (long long)(n0 + c0) * c_cs;
        // all 32 old values first (independent loads in flight together), then the stores: a load -> store chain per
        // element made the Add epilogue latency-bound (7 ms for a 65280 x 3840 x 256 update)
        float old[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) old[i] = (accum && n0 + c0 + i < n) ? dst[(long long)i * c_cs] : 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n0 + c0 + i < n) dst[(long long)i * c_cs] = fmaf(alpha, __uint_as_float(v[i]), old[i]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// out[r][kk] = split of src(r, kk) for kk < Kp (zero beyond k); src(r, kk) = p[r * rs + kk * cs]
__global__ void __launch_bounds__(256) pack_split_kernel(const float* __restrict__ src, long long rs, long long cs, int rows,
                                                         int k, int Kp, float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool k_fast = (cs < 0 ? -cs : cs) <= (rs < 0 ? -rs : rs);
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = k_fast ? r0 + i : r0 + tx;
    const int kk = k_fast ? k0 + tx : k0 + i;
    const float v = (r < rows && kk < k) ? src[(long long)r * rs + (long long)kk * cs] : 0.f;
    if (k_fast) t[i][tx] = v;
    else t[tx][i] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, kk = k0 + tx;
    if (r < rows) {
      const float v = t[i][tx];
      uint32_t h, l;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
      const float rem = v - __uint_as_float(h);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rem));
      hi[(size_t)r * Kp + kk] = __uint_as_float(h);
      lo[(size_t)r * Kp + kk] = __uint_as_float(l);
    }
  }
}

// dst = [dst +] alpha * sum_z W_z   (partials are column-major m x n, summed in z order)
__global__ void __launch_bounds__(256) tc_splitk_reduce_kernel(float* __restrict__ C, long long c_rs, long long c_cs, int m, int n,
                                                               int accum, float alpha, const float* __restrict__ W, int splits) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)m * n) return;
  const int row = (int)(e % m), col = (int)(e / m);
  double s = 0.0;
  for (int z = 0; z < splits; ++z) s += (double)W[(long long)z * m * n + e];
  float* cp = C + (long long)row * c_rs + (long long)col * c_cs;
  const float v = alpha * (float)s;
  *cp = accum ? (*cp + v) : v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && f) fn = (EncodeTiledFn)f;
  }
  return fn;
}

inline bool make_map(CUtensorMap* map, float* base, int rows, int Kp) {
  const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)Kp * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  return encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline void release(Workspace* ws) {
  if (ws->buf) cudaFree(ws->buf);
  ws->buf = nullptr;
  ws->bytes = 0;
}

// C(m x n, strides c_rs / c_cs) = [C +] alpha * A(m x k) * B(k x n). Returns false if the problem is not taken
// (driver entry point missing); the caller then uses the mma.sync kernel. `ws` grows on demand (cudaMalloc).
inline bool gemm_f32_tc(cudaStream_t st, float* C, long long c_rs, long long c_cs, int m, int n, int k, int accum, Operand a,
                        Operand b, float alpha, Workspace* ws) {
  if (m <= 0 || n <= 0 || k <= 0 || !encode_fn()) return false;
  const int Kp = (k + BK - 1) / BK * BK;
  const size_t a_elems = (size_t)m * Kp, b_elems = (size_t)n * Kp;
  // split-K when the output has too few 128 x 128 tiles to fill the SMs (tall-skinny V^T M products of the QR)
  const int kblocks = Kp / BK;
  const long long tiles = (long long)((m + BM - 1) / BM) * ((n + BN - 1) / BN);
  int splits = 1;
  if (tiles < 148 && kblocks >= 32) splits = (int)std::min<long long>((296 + tiles - 1) / tiles, kblocks / 8);
  if (splits < 2) splits = 1;
  int kb_per_split = (kblocks + splits - 1) / splits;
  splits = (kblocks + kb_per_split - 1) / kb_per_split;
  const size_t part_elems = splits > 1 ? (size_t)splits * m * n : 0;
  const size_t need = (2 * a_elems + 2 * b_elems + part_elems) * sizeof(float) + 8192;
  if (need > ws->bytes) {
    if (ws->buf) {
      cudaStreamSynchronize(st);
      cudaFree(ws->buf);
    }
    if (cudaMalloc(&ws->buf, need) != cudaSuccess) {
      ws->buf = nullptr;
      ws->bytes = 0;
      return false;
    }
    ws->bytes = need;
  }
  auto up = [](float* p) { return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); };
  float* Ah = up(ws->buf);
  float* Al = up(Ah + a_elems);
  float* Bh = up(Al + a_elems);
  float* Bl = up(Bh + b_elems);
  float* part = up(Bl + b_elems);
  if (reinterpret_cast<char*>(part + part_elems) > reinterpret_cast<char*>(ws->buf) + ws->bytes) return false;
  pack_split_kernel<<<dim3(Kp / 32, (m + 31) / 32), 256, 0, st>>>(a.p, a.rs, a.cs, m, k, Kp, Ah, Al);
  // B packed as its transpose: row j of the packed array is column j of B
  pack_split_kernel<<<dim3(Kp / 32, (n + 31) / 32), 256, 0, st>>>(b.p, b.cs, b.rs, n, k, Kp, Bh, Bl);
  CUtensorMap mAh, mAl, mBh, mBl;
  if (!make_map(&mAh, Ah, m, Kp) || !make_map(&mAl, Al, m, Kp) || !make_map(&mBh, Bh, n, Kp) || !make_map(&mBl, Bl, n, Kp))
    return false;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_f32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return false;
    configured = true;
  }
  const dim3 grid((unsigned)(((m + BM - 1) / BM) * ((n + BN - 1) / BN)), 1u, (unsigned)splits);
  if (splits == 1) {
    gemm_f32_tc_kernel<<<grid, 256, SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, C, c_rs, c_cs, m, n, kblocks, kblocks, 0, alpha, accum);
  } else {
    gemm_f32_tc_kernel<<<grid, 256, SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, part, 1, (long long)m, m, n, kblocks, kb_per_split,
                                                      (long long)m * n, 1.0f, 0);
    const long long total = (long long)m * n;
    tc_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(C, c_rs, c_cs, m, n, accum, alpha, part, splits);
  }
  return cudaGetLastError() == cudaSuccess;
}

}  // namespace tc
}  // namespace fb
