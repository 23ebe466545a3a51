// Tridiagonalization A = Q T Q^H of a self-adjoint matrix (f32 / f64, lower triangle only): ONE persistent cooperative
// kernel + GEMM-built T factors.
//
// Reference: faer/src/linalg/evd/tridiag.rs
//   tridiag_in_place 274-529: per column k — pending rank-2 update (reflector u and vector y of step k-1) applied to
//   column k (307-317), reflector from A[k+1.., k] (325-330), pending update applied to column k+1 (345-353), fused pass
//   over the lower triangle of A22 = A[k+2.., k+2..]: a_ij -= w_j u_i + u_j w_i, y_j = f sum_{i>j} a_ij x_i,
//   z_i += a_ij f x_j (tridiag_fused_op 36-160), y = y + z (377), y += A21/tau (462-466),
//   y_1 = (a11 + <A21, x>)/tau (467-475), b = (y_1 + <x, y>)/(2 tau) (476-485), y -= b (1; x) (486-491);
//   T factors of the reflector blocks via upgrade_householder_factor (506-528).
//
// B200 mapping (same skeleton as bidiag.cu): the stage is HBM-bound with a serial dependency per column, so the column
// loop is one persistent cooperative kernel (1 CTA per SM) with TWO grid barriers per column:
//   vector phase (every CTA, redundantly, identical bits): form column k with the pending update, its norm and the
//     reflector scalars; stage x (scaled tail), u (previous reflector), w (pending y) in shared memory; form column k+1
//     with the pending update (kept in a side buffer) and d = <A21, x>;
//   pass 1 (CTA = range of COLUMNS, square-root-law boundaries so that every CTA owns the same triangle area): rank-2
//     update of the lower triangle (read + write) and the column sums y_j = f sum_{i>j} a_ij x_i — CTA-local;
//   barrier;
//   pass 2 (CTA = strip of ROWS, equal-area boundaries): z_i = f sum_{j<=i} a_ij x_j — CTA-local; the strip owner forms
//     y_i = y_i + z_i + A21_i/tau and its share of <x, y>;
//   barrier; every CTA forms b and goes to the next column.
// Pass 2 re-reads the triangle instead of accumulating z_i across column owners in pass 1 (that needs a cross-CTA
// reduction of n-vectors per column); cost: 3 instead of 2 triangle transfers per column. See DESIGN.md §7.
#include <algorithm>

#include "panel_common.cuh"
#include "runtime.cuh"
#include "tensor_ops.cuh"

namespace fb {

namespace {

constexpr int TD_THREADS = 512;
constexpr int TD_CW = 4;      // columns per pass-1 group
constexpr int TD_RB = 4;      // rows per pass-1 iteration and thread
constexpr int TD_CH = 8192;   // rows resident in shared memory (n <= TD_CH)
constexpr int TD_PC = 64;     // columns finalised per batch
constexpr int TD_NW = TD_THREADS / 32;

template <class T>
struct TdScratch {
  T* vecg;   // [G][3][n] x / u / w of the current column for CTAs whose trailing block does not fit shared memory yet (or null)
  T* ycomb;  // [n]   y before the "- b (1; x)" correction
  T* ycol;   // [n]   column sums of pass 1
  T* cbuf;   // [2][n] column k (+1) with the older pending update applied; [k] holds the diagonal entry
  T* part;   // [2][G][PANEL_NV]
  unsigned long long* bar;
};

template <class T>
__global__ void __launch_bounds__(TD_THREADS, 1) tridiag_kernel(T* A, i64 cs, int n, T* H, i64 hs, TdScratch<T> sc) {
  extern __shared__ unsigned char td_smem_raw[];
  T* xs0 = reinterpret_cast<T*>(td_smem_raw);  // [TD_CH] scaled tail of the current reflector (global row - gb)
  T* us0 = xs0 + TD_CH;                        // [TD_CH] previous reflector
  T* ws0 = us0 + TD_CH;                        // [TD_CH] pending y
  T* part = ws0 + TD_CH;                       // [TD_PC][32]  (pass 2: red[TD_THREADS])
  T* fin = part + TD_PC * 32;                  // [TD_NW][PANEL_NV]
  T* scal = fin + TD_NW * PANEL_NV;            // [16]

  const int tid = threadIdx.x, lane = tid & 31;
  const int G = gridDim.x, bid = blockIdx.x;
  const T min_pos = TLim<T>::min_pos();
  const T sml = t_sqrt(min_pos), big = t_sqrt(T(1) / min_pos);
  unsigned long long nbar = 0;

  // prologue: column 0 (and the diagonal entry) into the side buffer
  for (int g = bid * TD_THREADS + tid; g < n; g += G * TD_THREADS) {
    sc.cbuf[g] = A[g];
    sc.ycomb[g] = T(0);
  }
  ++nbar;
  grid_barrier(sc.bar, nbar * (unsigned long long)G);

  T b_prev = T(0), y1n_prev = T(0);
  for (int k = 0; k < n; ++k) {
    const bool pend = k > 0;
    const T* cur = sc.cbuf + (size_t)(k & 1) * n;
    T* nxt = sc.cbuf + (size_t)((k + 1) & 1) * n;
    const int len = n - k - 1;
    const T y1 = pend ? y1n_prev - b_prev : T(0);
    if (bid == 0 && tid == 0) A[(i64)k * cs + k] = t_ldcg(&cur[k]) - (y1 + y1);  // tridiag.rs:311
    if (len == 0) break;
    const int gb = (k + 1) & ~31;
    // the three vectors cover rows gb .. n-1: in shared memory once that fits (n - gb <= TD_CH), until then in this CTA's
    // private slice of global memory (L2-resident; n > 8192 only). Written and read by this CTA alone.
    const bool vec_in_smem = n - gb <= TD_CH;
    T* x_s = vec_in_smem ? xs0 : sc.vecg + (size_t)bid * 3 * n;
    T* u_s = vec_in_smem ? us0 : x_s + n;
    T* w_s = vec_in_smem ? ws0 : u_s + n;

    // ================= vector phase =================
    // column k with the pending update (tridiag.rs:312-316), u and the corrected pending y, for rows >= k+1
    NormAcc<T> na = {T(0), T(0), T(0)};
    for (int g = gb + tid; g < n; g += TD_THREADS) {
      if (g < k + 1) continue;
      const T u = pend ? t_ldcg(&A[(i64)(k - 1) * cs + g]) : T(0);
      const T yf = pend ? t_ldcg(&sc.ycomb[g]) - b_prev * u : T(0);
      const T c = t_ldcg(&cur[g]) - (y1 * u + yf);
      x_s[g - gb] = c;
      u_s[g - gb] = u;
      w_s[g - gb] = yf;
      if (g > k + 1) na.add(c, sml, big);
    }
    {
      const T vals[3] = {na.s, na.m, na.b};
      block_publish_n<T, 3, TD_NW>(vals, fin, scal + 12);
      __syncthreads();
    }
    const T tail_norm = norm_from_acc(scal[12], scal[13], scal[14], sml, big);
    const HhScalars<T> hh = make_householder_scalars(x_s[k + 1 - gb], tail_norm);
    const T inv = hh.no_tail ? T(1) : hh.inv;
    const T tau_inv = T(1) / hh.tau;  // 0 when tau = +inf
    if (bid == 0 && tid == 0) {
      A[(i64)k * cs + k + 1] = hh.new_head;
      H[(i64)k * hs] = hh.tau;
    }
    const T u1 = u_s[k + 1 - gb], y1p = w_s[k + 1 - gb];
    __syncthreads();  // everybody has read the head before it is overwritten
    // scale the tail; column k+1 with the pending update (tridiag.rs:345-353) and d = <A21, x>
    T dacc = T(0);
    for (int g = gb + tid; g < n; g += TD_THREADS) {
      if (g < k + 1) continue;
      if (g == k + 1) {
        x_s[g - gb] = T(0);
        continue;
      }
      const T x = x_s[g - gb] * inv;
      x_s[g - gb] = x;
      T a21 = t_ldcg(&A[(i64)(k + 1) * cs + g]);
      if (pend) a21 -= u_s[g - gb] * y1p + w_s[g - gb] * u1;
      dacc = fma(a21, x, dacc);
      if (((g >> 7) % G) == bid) {  // one designated CTA per 128-row segment stores the results
        A[(i64)k * cs + g] = x;
        nxt[g] = a21;
      }
    }
    T a11 = t_ldcg(&A[(i64)(k + 1) * cs + k + 1]);
    if (pend) a11 -= (u1 * y1p + y1p * u1);
    if (bid == 0 && tid == 0) nxt[k + 1] = a11;
    {
      const T vals[1] = {dacc};
      block_publish_n<T, 1, TD_NW>(vals, fin, scal + 15);
      __syncthreads();
    }
    const T y1n = (a11 + scal[15]) * tau_inv;  // tridiag.rs:467-475

    const int L = n - k - 2;  // order of A22
    T d2 = T(0);
    if (L > 0) {
      const int Gp = min(G, max(1, L / 8));
      // ================= pass 1: columns [c0, c1) of the lower triangle =================
      if (bid < Gp) {
        const int c0 = k + 2 + (int)((double)L * (1.0 - sqrt(1.0 - (double)bid / (double)Gp)));
        const int c1 = bid + 1 == Gp ? n : k + 2 + (int)((double)L * (1.0 - sqrt(1.0 - (double)(bid + 1) / (double)Gp)));
        const int cper = c1 - c0;
        const int TC = min(TD_NW, pow2_ceil((cper + TD_CW - 1) / TD_CW));
        const int TR = TD_THREADS / TC;
        const int rl = tid & (TR - 1), cg = tid / TR, wr = rl >> 5, nwr = TR >> 5;
        for (int cb = c0; cb < c1; cb += TD_PC) {
          const int ce = min(c1, cb + TD_PC);
          for (int jb = cb + cg * TD_CW; jb < ce; jb += TC * TD_CW) {
            const int ncg = min(TD_CW, ce - jb);
            T uj[TD_CW], wj[TD_CW], acc[TD_CW];
#pragma unroll
            for (int c = 0; c < TD_CW; ++c) {
              const bool on = c < ncg;
              uj[c] = on ? u_s[jb + c - gb] : T(0);
              wj[c] = on ? w_s[jb + c - gb] : T(0);
              acc[c] = T(0);
            }
            T* Ac = A + (i64)jb * cs;
            // TD_RB rows x TD_CW columns per iteration: all loads first (16 independent 8-byte loads in flight per
            // thread), then the update, then the stores
            for (int g0 = (jb & ~31) + rl; g0 < n; g0 += TD_RB * TR) {
              T a[TD_RB][TD_CW];
#pragma unroll
              for (int r = 0; r < TD_RB; ++r) {
                const int g = g0 + r * TR;
#pragma unroll
                for (int c = 0; c < TD_CW; ++c)
                  a[r][c] = (g < n && c < ncg && g >= jb + c) ? t_ldcg(&Ac[(i64)c * cs + g]) : T(0);
              }
#pragma unroll
              for (int r = 0; r < TD_RB; ++r) {
                const int g = g0 + r * TR;
                if (g >= n || g < jb) continue;
                const T xi = x_s[g - gb], ui = u_s[g - gb], wi = w_s[g - gb];
                if (pend) {
#pragma unroll
                  for (int c = 0; c < TD_CW; ++c) {
                    a[r][c] = fma(-wj[c], ui, a[r][c]);
                    a[r][c] = fma(-uj[c], wi, a[r][c]);
                  }
#pragma unroll
                  for (int c = 0; c < TD_CW; ++c)
                    if (c < ncg && g >= jb + c) Ac[(i64)c * cs + g] = a[r][c];
                }
#pragma unroll
                for (int c = 0; c < TD_CW; ++c)
                  if (c < ncg && g > jb + c) acc[c] = fma(a[r][c], xi, acc[c]);
              }
            }
#pragma unroll
            for (int c = 0; c < TD_CW; ++c) {
              const T s = warp_sum(acc[c]);
              if (lane == 0 && c < ncg) part[(jb + c - cb) * 32 + wr] = s;
            }
          }
          __syncthreads();
          if (tid < ce - cb) {
            T dot = T(0);
            for (int w = 0; w < nwr; ++w) dot += part[tid * 32 + w];
            sc.ycol[cb + tid] = tau_inv * dot;
          }
          __syncthreads();
        }
      }
      ++nbar;
      grid_barrier(sc.bar, nbar * (unsigned long long)G);

      // ================= pass 2: rows [r0, r1): z_i = f sum_{j <= i} a_ij x_j ; y_i ; <x, y> =================
      T d2acc = T(0);
      if (bid < Gp) {
        const int r0 = k + 2 + (int)((double)L * sqrt((double)bid / (double)Gp));
        const int r1 = bid + 1 == Gp ? n : k + 2 + (int)((double)L * sqrt((double)(bid + 1) / (double)Gp));
        const int rows = r1 - r0;
        const int RS = min(TD_THREADS, 32 * pow2_ceil((rows + 31) / 32));
        const int P = TD_THREADS / RS;
        const int rl = tid & (RS - 1), ph = tid / RS;
        T* red = part;
        for (int sub = 0; sub < rows; sub += RS) {
          const int g = r0 + sub + rl;
          const bool valid = g < r1;
          T acc = T(0);
          if (valid) {
            const T* Ar = A + g;
            int j = k + 2 + ph;
            for (; j + 15 * P <= g; j += 16 * P) {  // 16 independent loads in flight, fixed accumulation order
              T v[16];
#pragma unroll
              for (int q2 = 0; q2 < 16; ++q2) v[q2] = t_ldcg(&Ar[(i64)(j + q2 * P) * cs]);
#pragma unroll
              for (int q2 = 0; q2 < 16; ++q2) acc = fma(v[q2], x_s[j + q2 * P - gb], acc);
            }
#pragma unroll 4
            for (; j <= g; j += P) acc = fma(t_ldcg(&Ar[(i64)j * cs]), x_s[j - gb], acc);
          }
          if (sub > 0) __syncthreads();
          red[ph * RS + rl] = acc;
          __syncthreads();
          if (tid < RS && valid) {
            T s = T(0);
            for (int p = 0; p < P; ++p) s += red[p * RS + rl];
            T y = t_ldcg(&sc.ycol[g]) + tau_inv * s;      // tridiag.rs:377 (y = y + z)
            y += t_ldcg(&nxt[g]) * tau_inv;               // tridiag.rs:462-466
            sc.ycomb[g] = y;
            d2acc = fma(x_s[g - gb], y, d2acc);
          }
        }
      }
      {
        const T vals[1] = {d2acc};
        block_publish_n<T, 1, TD_NW>(vals, fin, sc.part + ((nbar & 1) * G + bid) * PANEL_NV);
        ++nbar;
        grid_barrier(sc.bar, nbar * (unsigned long long)G);
      }
      if (tid < 32) {
        T s[1];
        reduce_partials<T, 1>(sc.part + (((nbar - 1) & 1) * G) * PANEL_NV, G, s);
        if (lane == 0) scal[11] = s[0];
      }
      __syncthreads();
      d2 = scal[11];
    } else {
      // no trailing block: the only visibility requirement is nxt[k+1] for the next column
      ++nbar;
      grid_barrier(sc.bar, nbar * (unsigned long long)G);
    }
    const T b = (y1n + d2) * T(0.5) * tau_inv;  // tridiag.rs:476-485
    y1n_prev = y1n;
    b_prev = b;
  }
}

template <class T>
__global__ void td_row0_to_diag_kernel(T* H, i64 rs, i64 cs, int ncols, int bs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < ncols && (k % bs) != 0) H[(i64)(k % bs) * rs + (i64)k * cs] = H[(i64)k * cs];
}

}  // namespace

template <class T>
void tridiag_in_place(cudaStream_t st, View<T> A, View<T> H) {
  const i64 n = A.nrows;
  FB_ASSERT(A.ncols == n, "tridiag_in_place: square matrix required");
  FB_ASSERT(H.ncols == (n > 0 ? n - 1 : 0), "tridiag_in_place: householder factor must have n - 1 columns");
  if (n == 0) return;
  FB_ASSERT(n <= 1 || H.nrows > 0, "tridiag_in_place: empty Householder factor");
  FB_ASSERT(A.rs == 1, "tridiag_in_place: column-major (row stride 1) matrix required");
  FB_ASSERT(n < (1ll << 30), "tridiag_in_place: dimension too large");
  int dev = 0, num_sms = 0;
  FB_CUDA_CHECK(cudaGetDevice(&dev));
  FB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const int G = num_sms;
  const size_t vec_elems = n > TD_CH ? (size_t)G * 3 * (size_t)n : 0;
  const size_t elems = (size_t)4 * n + (size_t)2 * G * PANEL_NV + 32 + vec_elems;
  char* buf = (char*)ws_alloc(elems * sizeof(T) + 64);
  TdScratch<T> sc;
  sc.vecg = vec_elems ? (T*)buf + ((size_t)4 * n + (size_t)2 * G * PANEL_NV + 8 + 8) : nullptr;
  sc.ycomb = (T*)buf;
  sc.ycol = sc.ycomb + n;
  sc.cbuf = sc.ycol + n;
  sc.part = sc.cbuf + 2 * n;
  sc.bar = (unsigned long long*)(((uintptr_t)(sc.part + (size_t)2 * G * PANEL_NV + 8) + 15) & ~(uintptr_t)15);
  FB_CUDA_CHECK(cudaMemsetAsync(sc.bar, 0, 8, st));
  const size_t smem = ((size_t)3 * TD_CH + TD_PC * 32 + TD_NW * PANEL_NV + 16) * sizeof(T);
  static bool configured = false;
  if (!configured) {
    FB_CUDA_CHECK(cudaFuncSetAttribute(tridiag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  {
    T* Ap = A.ptr;
    i64 cs = A.cs;
    int ni = (int)n;
    T* hp = H.ptr;
    i64 hs = H.cs;
    void* args[] = {&Ap, &cs, &ni, &hp, &hs, &sc};
    FB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)tridiag_kernel<T>, dim3(G), dim3(TD_THREADS), args, smem, st));
    note_launch();
  }
  // T factors (tridiag.rs:506-528)
  const i64 n1 = n - 1, bs = H.nrows;
  if (n1 > 0 && bs > 1) {
    td_row0_to_diag_kernel<T><<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(H.ptr, H.rs, H.cs, (int)n1, (int)bs);
    note_launch();
    View<T> As = A.sub(1, 0, n1, n1);
    for (i64 j = 0; j < n1; j += bs) {
      const i64 b = std::min(bs, n1 - j);
      householder_build_t<T>(st, cview(As.sub(j, j, n1 - j, b)), H.sub(0, j, b, b));
    }
  }
  FB_CUDA_CHECK(cudaGetLastError());
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  ws_free(buf);
}

template void tridiag_in_place<double>(cudaStream_t, View<double>, View<double>);
template void tridiag_in_place<float>(cudaStream_t, View<float>, View<float>);

}  // namespace fb
