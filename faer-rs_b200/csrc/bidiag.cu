// Bidiagonalization A = U B V^H (f32 / f64, m >= n): ONE persistent cooperative kernel + GEMM-built T factors.
//
// Reference: faer/src/linalg/svd/bidiag.rs
//   bidiag_in_place 47-256: per column k — apply the PENDING rank-2 update (u, y, z, v of step k-1) to column k and
//   row k (86-103), left reflector from column k (104-107), fused pass over A22 that applies the pending update and
//   computes y = u^H A22 (bidiag_fused_op 257-373), y = (y + A12)/tau_l ; A12 -= y (151-154), A12 /= |A12| (155-159),
//   z = A22 A12^H (160-167), right reflector from row k (176-186), b = y_0 + <y_1.., v> (187-193), z fix-up (194-219);
//   T factors of the left / right reflector blocks via upgrade_householder_factor (222-255).
//
// B200 mapping. The stage is HBM-bound (per column: one read+write pass and one read pass over the trailing matrix,
// 3 * 8 B * (n-k)^2) with two global dependencies per column, so the whole column loop is one
// persistent cooperative kernel (1 CTA per SM) with exactly TWO grid barriers per column:
//   pass 1 (CTA = contiguous range of COLUMNS, all rows): the vectors u, u_prev, z live in shared memory; each thread
//     streams a group of 4 columns, 4 rows x 4 columns per iteration: a -= u_prev*y_j + z*v_j ; store ; acc_j += u*a.
//     Column dots are CTA-local, so y_j, the updated row entry A12_j and its contribution to |A12|^2 and <y, A12> need no cross-CTA traffic;
//   barrier; every CTA reduces the 7 published partials in a fixed order and forms the right reflector scalars;
//   pass 2 (CTA = strip of ROWS, all columns): z_i = sum_j A22[i, j] * A12_j is CTA-local per row; the strip owner
//     finishes z_i (fix-up formula), forms the next column  A[i, k+1] - u_i*y_{k+1} - z_i  (the pending update applied
//     to the column the next left reflector is built from) and publishes its partial squared norm;
//   barrier; every CTA forms the left reflector scalars of column k+1.
// All reductions are fixed-order (deterministic for a given SM count). Row loops start at a 32-row-aligned global row
// so that warp accesses stay sector-aligned as k advances. Loads of data produced by other CTAs use ld.global.cg.
#include <algorithm>

#include "panel_common.cuh"
#include "runtime.cuh"
#include "tensor_ops.cuh"

namespace fb {

namespace {

constexpr int BD_THREADS = 512;
constexpr int BD_CW = 4;       // columns per pass-1 group
constexpr int BD_RB = 4;       // rows per pass-1 iteration and thread
constexpr int BD_CH = 8192;    // rows (pass 1) / columns (pass 2) resident in shared memory per chunk
constexpr int BD_PC = 64;      // columns finalised per batch
constexpr int BD_NW = BD_THREADS / 32;
constexpr int BD_NV = 8;       // published values per CTA and barrier

template <class T>
struct BdScratch {
  T* ubuf;  // [m] unscaled tail of the current left reflector (global row index)
  T* zbuf;  // [m]
  T* ybuf;  // [n]
  T* a12;   // [n] updated (unnormalised) row k (global column index)
  T* part;  // [2][G][BD_NV]
  T* head;  // [1] head of the next left reflector
  unsigned long long* bar;
};

static_assert(BD_NV == PANEL_NV, "published-value stride");
template <class T, int NV>
__device__ __forceinline__ void block_publish(const T (&vals)[NV], T* fin, T* dst) {
  block_publish_n<T, NV, BD_NW>(vals, fin, dst);
}

// A: column-major (row stride 1), m >= n. Hl / Hr: pointers to row 0 of H_left / H_right, hls / hrs = their column strides.
template <class T>
__global__ void __launch_bounds__(BD_THREADS, 1) bidiag_kernel(T* A, i64 cs, int m, int n, T* Hl, i64 hls, T* Hr, i64 hrs,
                                                               BdScratch<T> sc) {
  extern __shared__ unsigned char bd_smem_raw[];
  T* u_s = reinterpret_cast<T*>(bd_smem_raw);  // [BD_CH]
  T* up_s = u_s + BD_CH;                       // [BD_CH]
  T* z_s = up_s + BD_CH;                       // [BD_CH]   (pass 2: the row A12)
  T* part = z_s + BD_CH;                       // [BD_PC][32] (pass 2: red[BD_THREADS])
  T* a12t = part + BD_PC * 32;                 // [BD_PC]
  T* fin = a12t + BD_PC;                       // [BD_NW][BD_NV]
  T* scal = fin + BD_NW * BD_NV;               // [16]

  const int tid = threadIdx.x, lane = tid & 31;
  const int G = gridDim.x, bid = blockIdx.x;
  const int size = min(m, n);
  const T min_pos = TLim<T>::min_pos();
  const T sml = t_sqrt(min_pos), big = t_sqrt(T(1) / min_pos);
  unsigned long long nbar = 0;

  // ---- prologue: column 0 is the first reflector's (head, tail); no pending update yet ----
  {
    NormAcc<T> na = {T(0), T(0), T(0)};
    for (int g = bid * BD_THREADS + tid; g < m; g += G * BD_THREADS) {
      const T x = A[g];
      if (g == 0) sc.head[0] = x;
      else {
        sc.ubuf[g] = x;
        na.add(x, sml, big);
      }
      sc.zbuf[g] = T(0);
    }
    for (int j = bid * BD_THREADS + tid; j < n; j += G * BD_THREADS) {
      sc.ybuf[j] = T(0);
      sc.a12[j] = T(0);
    }
    const T vals[3] = {na.s, na.m, na.b};
    block_publish<T, 3>(vals, fin, sc.part + ((nbar & 1) * G + bid) * BD_NV);
    ++nbar;
    grid_barrier(sc.bar, nbar * (unsigned long long)G);
  }

  T vscale_prev = T(0);  // v_ess[j] of row k-1 = a12[j] * vscale_prev
  for (int k = 0; k < size; ++k) {
    // ================= left reflector of column k (householder.rs:59-107) =================
    if (tid < 32) {
      T s[3];
      reduce_partials<T, 3>(sc.part + (((nbar - 1) & 1) * G) * BD_NV, G, s);
      const T tail_norm = norm_from_acc(s[0], s[1], s[2], sml, big);
      const HhScalars<T> h = make_householder_scalars(t_ldcg(sc.head), tail_norm);
      if (lane == 0) {
        scal[0] = h.no_tail ? T(1) : h.inv;
        scal[1] = h.no_tail ? T(0) : T(1) / h.tau;
        scal[2] = h.tau;
        scal[3] = h.new_head;
      }
    }
    __syncthreads();
    const T inv_l = scal[0], tl_inv = scal[1];
    if (bid == 0 && tid == 0) {
      A[(i64)k * cs + k] = scal[3];
      Hl[(i64)k * hls] = scal[2];
    }
    // scaled tail into place (rows spread over the CTAs in 128-row segments)
    for (int seg = bid; seg * 128 < m; seg += G) {
      const int g = seg * 128 + (tid & 127);
      if (tid < 128 && g > k && g < m) A[(i64)k * cs + g] = t_ldcg(&sc.ubuf[g]) * inv_l;
    }
    const int nr = n - k - 1;
    if (nr == 0) break;
    const bool pend = k > 0;

    // ================= pass 1: pending update + y = u^H A22, CTA-local per column =================
    T f_all[3] = {T(0), T(0), T(0)}, f_tail[3] = {T(0), T(0), T(0)}, f_dyv = T(0);
    {
      const int cper = (nr + G - 1) / G;
      const int c0 = k + 1 + bid * cper, c1 = min(n, c0 + cper);
      const int TC = min(BD_NW, pow2_ceil((cper + BD_CW - 1) / BD_CW));
      const int TR = BD_THREADS / TC;
      const int rl = tid & (TR - 1), cg = tid / TR, wr = rl >> 5, nwr = TR >> 5;
      const int gb = k & ~31;
      int staged = -1;
      for (int cb = c0; cb < c1; cb += BD_PC) {
        const int ce = min(c1, cb + BD_PC);
        for (int ch0 = gb, chunk = 0; ch0 < m; ch0 += BD_CH, ++chunk) {
          const int ch1 = min(m, ch0 + BD_CH);
          if (staged != chunk) {
            if (staged >= 0) __syncthreads();
            // predicated + unrolled: the 12 loads of four iterations are in flight together
#pragma unroll 4
            for (int g = ch0 + tid; g < ch1; g += BD_THREADS) {
              const bool on = g >= k;
              const T uv = (on && g > k) ? t_ldcg(&sc.ubuf[g]) * inv_l : T(0);
              const T pv = (on && pend) ? t_ldcg(&A[(i64)(k - 1) * cs + g]) : T(0);
              const T zv = (on && pend) ? t_ldcg(&sc.zbuf[g]) : T(0);
              if (on) {
                u_s[g - ch0] = uv;
                up_s[g - ch0] = pv;
                z_s[g - ch0] = zv;
              }
            }
            __syncthreads();
            staged = chunk;
          }
          for (int jb = cb + cg * BD_CW; jb < ce; jb += TC * BD_CW) {
            const int ncg = min(BD_CW, ce - jb);
            T yv[BD_CW], vv[BD_CW], acc[BD_CW];
#pragma unroll
            for (int c = 0; c < BD_CW; ++c) {
              const bool on = c < ncg && pend;
              yv[c] = on ? t_ldcg(&sc.ybuf[jb + c]) : T(0);
              vv[c] = on ? t_ldcg(&sc.a12[jb + c]) * vscale_prev : T(0);
              acc[c] = T(0);
            }
            if (pend && rl == 0 && chunk == 0) {
#pragma unroll
              for (int c = 0; c < BD_CW; ++c)
                if (c < ncg) A[(i64)(jb + c) * cs + (k - 1)] = vv[c];  // v_ess of row k-1 goes to its final place
            }
            T* Ac = A + (i64)jb * cs;
            // BD_RB rows x BD_CW columns per iteration: all BD_RB * BD_CW loads are issued before the first use, so a
            // thread keeps 16 independent 8-byte loads in flight (the HBM stream needs > 40 KB in flight per SM)
            for (int g0 = ch0 + rl; g0 < ch1; g0 += BD_RB * TR) {
              T a[BD_RB][BD_CW];
#pragma unroll
              for (int r = 0; r < BD_RB; ++r) {
                const int g = g0 + r * TR;
                const bool rowon = g < ch1 && g >= k;
#pragma unroll
                for (int c = 0; c < BD_CW; ++c) a[r][c] = (rowon && c < ncg) ? t_ldcg(&Ac[(i64)c * cs + g]) : T(0);
              }
#pragma unroll
              for (int r = 0; r < BD_RB; ++r) {
                const int g = g0 + r * TR;
                if (g >= ch1 || g < k) continue;
                const T uu = u_s[g - ch0];
                if (pend) {
                  const T pp = up_s[g - ch0], zz = z_s[g - ch0];
#pragma unroll
                  for (int c = 0; c < BD_CW; ++c) {
                    a[r][c] = fma(-pp, yv[c], a[r][c]);
                    a[r][c] = fma(-zz, vv[c], a[r][c]);
                  }
                }
                if (g == k) {
#pragma unroll
                  for (int c = 0; c < BD_CW; ++c)
                    if (c < ncg) a12t[jb + c - cb] = a[r][c];
                } else {
#pragma unroll
                  for (int c = 0; c < BD_CW; ++c) {
                    if (c < ncg) {
                      if (pend) Ac[(i64)c * cs + g] = a[r][c];
                      acc[c] = fma(uu, a[r][c], acc[c]);
                    }
                  }
                }
              }
            }
#pragma unroll
            for (int c = 0; c < BD_CW; ++c) {
              const T s = warp_sum(acc[c]);
              if (lane == 0 && c < ncg) {
                T* p = &part[(jb + c - cb) * 32 + wr];
                *p = chunk == 0 ? s : *p + s;
              }
            }
          }
        }
        __syncthreads();
        // ---- finalise the batch: y_j = (dot_j + A12_j)/tau_l ; A12_j -= y_j   (bidiag.rs:151-154)
        if (tid < ce - cb) {
          const int j = cb + tid;
          T dot = T(0);
          for (int w = 0; w < nwr; ++w) dot += part[tid * 32 + w];
          T a12v = a12t[tid];
          const T y = (dot + a12v) * tl_inv;
          a12v -= y;
          sc.ybuf[j] = y;
          sc.a12[j] = a12v;
          const T xs = a12v * sml, xb = a12v * big;
          f_all[0] = fma(xs, xs, f_all[0]);
          f_all[1] = fma(a12v, a12v, f_all[1]);
          f_all[2] = fma(xb, xb, f_all[2]);
          if (j > k + 1) {
            f_tail[0] = fma(xs, xs, f_tail[0]);
            f_tail[1] = fma(a12v, a12v, f_tail[1]);
            f_tail[2] = fma(xb, xb, f_tail[2]);
            f_dyv = fma(y, a12v, f_dyv);
          }
        }
        __syncthreads();
      }
    }
    {
      const T vals[7] = {f_all[0], f_all[1], f_all[2], f_tail[0], f_tail[1], f_tail[2], f_dyv};
      block_publish<T, 7>(vals, fin, sc.part + ((nbar & 1) * G + bid) * BD_NV);
      ++nbar;
      grid_barrier(sc.bar, nbar * (unsigned long long)G);
    }

    // ================= row norm, right reflector of row k (bidiag.rs:155-193) =================
    if (tid < 32) {
      T s[7];
      reduce_partials<T, 7>(sc.part + (((nbar - 1) & 1) * G) * BD_NV, G, s);
      const T norm = norm_from_acc(s[0], s[1], s[2], sml, big);
      const T tail_raw = norm_from_acc(s[3], s[4], s[5], sml, big);
      // (a subnormal norm would overflow its reciprocal in f32: such a row is numerically zero)
      const T norm_inv = norm >= min_pos ? T(1) / norm : T(1);
      const T y1 = t_ldcg(&sc.ybuf[k + 1]);
      const T head_n = t_ldcg(&sc.a12[k + 1]) * norm_inv;
      // (k + 1 < size here: with m >= n the last column leaves through the nr == 0 exit above)
      T vscale = norm_inv, inv_r = T(0), tr_inv = T(0), m_inf = T(1);
      const HhScalars<T> h = make_householder_scalars(head_n, tail_raw * norm_inv);
      const T beta_r = h.new_head, tau_r = h.tau;
      if (!h.no_tail) {
        vscale = norm_inv * h.inv;
        inv_r = h.inv;
        tr_inv = T(1) / h.tau;
        m_inf = T(0);
      }
      const T b = y1 + s[6] * vscale;
      if (lane == 0) {
        scal[4] = norm_inv;
        scal[5] = beta_r;
        scal[6] = inv_r;
        scal[7] = b;
        scal[8] = tr_inv;
        scal[9] = m_inf;
        scal[10] = vscale;
        scal[11] = y1;
        if (bid == 0) {
          A[(i64)(k + 1) * cs + k] = beta_r * norm;  // the stored head is un-normalised again (bidiag.rs:186)
          Hr[(i64)k * hrs] = tau_r;
        }
      }
    }
    __syncthreads();
    const T norm_inv = scal[4], beta_r = scal[5], inv_r = scal[6], bco = scal[7], tr_inv = scal[8], y1 = scal[11];
    const bool m_inf = scal[9] != T(0);
    vscale_prev = scal[10];

    // ================= pass 2: z = A22 A12^H, CTA-local per row; z fix-up; next column =================
    NormAcc<T> na = {T(0), T(0), T(0)};
    {
      const int gb2 = (k + 1) & ~31;
      const int q = (m - gb2 + 32 * G - 1) / (32 * G);
      const int RS = 32 * pow2_ceil(q);  // rows per strip (host guarantees RS <= BD_THREADS)
      const int P = BD_THREADS / RS;
      const int rl = tid & (RS - 1), ph = tid / RS;
      const int g = gb2 + bid * RS + rl;
      const bool valid = g >= k + 1 && g < m;
      const bool active = gb2 + bid * RS < m;
      T acc = T(0);
      T* red = part;
      for (int cc0 = 0; cc0 < nr; cc0 += BD_CH) {
        const int cn = min(BD_CH, nr - cc0);
        if (cc0 > 0) __syncthreads();
        if (active) {
#pragma unroll 4
          for (int jj = tid; jj < cn; jj += BD_THREADS) z_s[jj] = t_ldcg(&sc.a12[k + 1 + cc0 + jj]);
        }
        __syncthreads();
        if (valid) {
          const T* Ar = A + (i64)(k + 1 + cc0) * cs + g;
          // 16 independent loads in flight per thread (fixed order of accumulation: deterministic)
          int jj = ph;
          for (; jj + 15 * P < cn; jj += 16 * P) {
            T v[16];
#pragma unroll
            for (int q2 = 0; q2 < 16; ++q2) v[q2] = t_ldcg(&Ar[(i64)(jj + q2 * P) * cs]);
#pragma unroll
            for (int q2 = 0; q2 < 16; ++q2) acc = fma(v[q2], z_s[jj + q2 * P], acc);
          }
#pragma unroll 4
          for (; jj < cn; jj += P) acc = fma(t_ldcg(&Ar[(i64)jj * cs]), z_s[jj], acc);
        }
      }
      red[ph * RS + rl] = acc;
      __syncthreads();
      if (tid < RS && valid) {
        T s = T(0);
        for (int p = 0; p < P; ++p) s += red[p * RS + rl];
        const T zn = s * norm_inv;
        const T a = t_ldcg(&A[(i64)(k + 1) * cs + g]);
        const T u = t_ldcg(&sc.ubuf[g]) * inv_l;
        T w;
        if (!m_inf) {  // bidiag.rs:195-206
          w = zn - a * beta_r;
          w = w * inv_r;
          w = w - u * bco;
        } else {       // bidiag.rs:208-218
          w = a - u * bco;
        }
        const T z = w * tr_inv;
        sc.zbuf[g] = z;
        // column k+1 with the pending update of this step applied (bidiag.rs:86-96 of the NEXT iteration; v_0 = 1)
        const T cnx = a - (u * y1 + z);
        if (g == k + 1) sc.head[0] = cnx;
        else {
          sc.ubuf[g] = cnx;
          na.add(cnx, sml, big);
        }
      }
    }
    {
      const T vals[3] = {na.s, na.m, na.b};
      block_publish<T, 3>(vals, fin, sc.part + ((nbar & 1) * G + bid) * BD_NV);
      ++nbar;
      grid_barrier(sc.bar, nbar * (unsigned long long)G);
    }
  }
}

template <class T>
__global__ void row0_to_diag_kernel(T* H, i64 rs, i64 cs, int ncols, int bs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < ncols && (k % bs) != 0) H[(i64)(k % bs) * rs + (i64)k * cs] = H[(i64)k * cs];
}

}  // namespace

template <class T>
void bidiag_in_place(cudaStream_t st, View<T> A, View<T> Hl, View<T> Hr) {
  const i64 m = A.nrows, n = A.ncols, size = std::min(m, n);
  FB_ASSERT(m >= n, "bidiag_in_place: nrows >= ncols required (the SVD driver transposes wide inputs, svd/mod.rs:560-575)");
  FB_ASSERT(Hl.ncols == size && Hr.ncols == (size > 0 ? size - 1 : 0), "bidiag_in_place: H_left / H_right column counts");
  FB_ASSERT(Hl.nrows > 0 && (size <= 1 || Hr.nrows > 0), "bidiag_in_place: empty Householder factor");
  if (size == 0) return;
  FB_ASSERT(A.rs == 1, "bidiag_in_place: column-major (row stride 1) matrix required");
  int dev = 0, num_sms = 0;
  FB_CUDA_CHECK(cudaGetDevice(&dev));
  FB_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const int G = num_sms;
  FB_ASSERT(m <= (i64)BD_THREADS * G - 64 && n < (1ll << 30), "bidiag_in_place: matrix too tall for the row-strip pass");

  const size_t elems = (size_t)2 * m + 2 * n + (size_t)2 * G * BD_NV + 8;
  char* buf = (char*)ws_alloc(elems * sizeof(T) + 96);
  BdScratch<T> sc;
  sc.ubuf = (T*)buf;
  sc.zbuf = sc.ubuf + m;
  sc.ybuf = sc.zbuf + m;
  sc.a12 = sc.ybuf + n;
  sc.part = (T*)(((uintptr_t)(sc.a12 + n) + 15) & ~(uintptr_t)15);  // records are read as 16-byte vectors
  sc.head = sc.part + (size_t)2 * G * BD_NV;
  sc.bar = (unsigned long long*)(((uintptr_t)(sc.head + 8) + 15) & ~(uintptr_t)15);
  FB_CUDA_CHECK(cudaMemsetAsync(sc.bar, 0, 8, st));
  const size_t smem = ((size_t)3 * BD_CH + BD_PC * 32 + BD_PC + BD_NW * BD_NV + 16) * sizeof(T);
  static bool configured = false;
  if (!configured) {
    FB_CUDA_CHECK(cudaFuncSetAttribute(bidiag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  {
    T* Ap = A.ptr;
    i64 cs = A.cs;
    int mi = (int)m, ni = (int)n;
    T* hl = Hl.ptr;
    i64 hls = Hl.cs;
    T* hr = Hr.ptr;
    i64 hrs = Hr.cs;
    void* args[] = {&Ap, &cs, &mi, &ni, &hl, &hls, &hr, &hrs, &sc};
    FB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)bidiag_kernel<T>, dim3(G), dim3(BD_THREADS), args, smem, st));
    note_launch();
  }
  // T factors (bidiag.rs:222-255): the diagonal of every block is the row of taus, the strict upper part is V^H V
  const i64 bl = Hl.nrows, br = Hr.nrows;
  if (bl > 1) {
    row0_to_diag_kernel<T><<<(unsigned)((size + 255) / 256), 256, 0, st>>>(Hl.ptr, Hl.rs, Hl.cs, (int)size, (int)bl);
    note_launch();
    for (i64 j = 0; j < size; j += bl) {
      const i64 b = std::min(bl, size - j);
      householder_build_t<T>(st, cview(A.sub(j, j, m - j, b)), Hl.sub(0, j, b, b));
    }
  }
  if (size > 1 && br > 1) {
    const i64 s1 = size - 1;
    row0_to_diag_kernel<T><<<(unsigned)((s1 + 255) / 256), 256, 0, st>>>(Hr.ptr, Hr.rs, Hr.cs, (int)s1, (int)br);
    note_launch();
    View<T> At = A.sub(0, 1, s1, n - 1).t();  // reflector k is column k of At, starting at row k
    for (i64 j = 0; j < s1; j += br) {
      const i64 b = std::min(br, s1 - j);
      householder_build_t<T>(st, cview(At.sub(j, j, At.nrows - j, b)), Hr.sub(0, j, b, b));
    }
  }
  FB_CUDA_CHECK(cudaGetLastError());
  FB_CUDA_CHECK(cudaStreamSynchronize(st));
  ws_free(buf);
}

template void bidiag_in_place<double>(cudaStream_t, View<double>, View<double>, View<double>);
template void bidiag_in_place<float>(cudaStream_t, View<float>, View<float>, View<float>);

}  // namespace fb
