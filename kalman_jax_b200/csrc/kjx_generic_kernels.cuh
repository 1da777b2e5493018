// kjx_generic_kernels.cuh — filter / smoother for a general block-diagonal prior (Sum of Matern / quasi-periodic
// components, spatio-temporal Matern-5/2): state dimension d <= 64 in shared memory, d <= 256 in per-CTA global scratch.
//
// Parallel in time like the d <= 3 path, at CTA granularity: the series is cut into chunks of consecutive steps, one
// CTA per chunk.  generic_fold_kernel reduces every chunk to a filter scan element (A,b,C,eta,J) — O(d^2) per step
// because A(dt) is block-sparse and the update is rank one —, generic_boundary_kernel turns the elements into the
// filtered state entering every chunk and the information (eta, J) of everything after it (two CTAs: prefix, suffix),
// and then the filter / smoother kernels run all chunks concurrently, each in reference order inside its chunk
// (sde_gp.py:271-306 and :349-379).  The RTS gains only depend on filtered moments, so generic_gain_kernel computes
// all of them in parallel before the backward recursion.  With one chunk (short series, or sites that must be
// initialised from the predictive) it is the plain sequential walk.
//
// How a CTA (8 warps) spends a step:
//   * A(dt) is never formed densely: every row of the block-diagonal transition has at most 6 non-zeros
//     (transition_row, kjx_elementwise_kernels.cuh), so A X A^T costs at most 12 d^2 flops instead of 4 d^3;
//   * the update is rank one (f = 1): K = P H^T / S, P -= K (H P);
//   * the dense d^3 work — the two products of the RTS update, the trailing updates of the blocked Cholesky /
//     triangular solves of the gain (utils.py:25-30), the products of the boundary walk — goes through the FP64
//     tensor pipe: mma.sync.m8n8k4.f64 on 16 x 16 blocks per warp (gemm_cta).  Measured on this B200 the DMMA path
//     sustains 37 TFLOP/s against 34 TFLOP/s for plain DFMA (scripts/micro/dmma_rate.cu) with 1/8 of the issue slots;
//   * matrices are stored row-major with a padded leading dimension ld = 4 (mod 8) so that both the row-wise and the
//     column-wise 8 x 4 fragment loads hit every bank exactly twice (conflict-free), and with zero padding up to the
//     next multiple of 4 so the k loop needs no tail;
//   * element-wise passes use (row = warp, column = lane) loops: no integer division anywhere in a step.
#pragma once
#include <stdint.h>
#include <type_traits>
#include "kjx_elementwise_kernels.cuh"
#include "kjx_small_math.cuh"

namespace kjx {

constexpr int KJX_GEN_DMAX = 64;          // the working set (6 padded d x d tiles) fits one SM's shared memory up to here
constexpr int KJX_GEN_DMAX_GLOBAL = 256;  // above it the same kernels keep their matrices in per-CTA global scratch
constexpr int KJX_GEN_THREADS = 256;      // fold / boundary / filter / smoother: one CTA per SM (512 threads cap registers at 128 and spill)
constexpr int KJX_GEN_GAIN_THREADS = 128; // gain kernel: three small CTAs per SM (its serial diagonal-block phases overlap across CTAs)
constexpr int KJX_GEN_WARPS = KJX_GEN_THREADS / 32;   // most warps any generic kernel runs with

template <typename T> struct GenArgs {
  DiscArgs disc;           // block structure of A(dt)
  SiteModel site;
  int d;
  int64_t N;
  const double* Pinf;      // device [d,d]
  const double* H;         // device [d]      (constant measurement row)  or null
  const double* H_steps;   // device [N,d]    (per-step measurement row)  or null
  const T* dt; const T* y; const uint8_t* mask; const T* site_mean; const T* site_var;
  int init_sites;          // sites come from the predictive (sde_gp.py:287, site_params=None)
  T* filt_mean; T* filt_cov; T* out_site_mean; T* out_site_var; T* nlml;
  const T* filt_mean_in; const T* filt_cov_in;
  T* smooth_mean; T* smooth_cov; T* new_site_mean; T* new_site_var;
  int return_full; int update_sites;
  // chunked scan (one CTA per chunk)
  int64_t chunk_len; int nchunks;
  T* sum;                  // [nchunks][3 d^2 + 2 d]  chunk elements: A | C | J | b | eta   (dense, leading dimension d;
                           //                         equilibrated coordinates, see scale_vectors)
  T* st_m; T* st_P;        // [nchunks][d], [nchunks][d^2]  filtered state entering every chunk
  T* in_eta; T* in_J;      // [nchunks][d], [nchunks][d^2]  information of everything after every chunk (equilibrated)
  T* gsum;                 // [ngroups][3 d^2 + 2 d]   group elements of the two-level boundary walk
  T* chunk_nlml;           // [nchunks]
  // per-call precomputation (generic_rows_kernel): the non-zeros of every row of A(dt[n]) for all steps, so the chain
  // kernels load KJX_ROW_W d scalars per step instead of evaluating exp / sincos in double on their critical path
  const T* rows;           // [N + 1][d][roww]; slot N is A(0) = I (the smoother's transition out of the last step, sde_gp.py:337)
  const int* c0n_g;        // [d]  first column | nnz << 24 of every row (step-invariant)
  int roww;                // values stored per row: 4, or 6 (= KJX_ROW_W) when the prior has a 6 x 6 block
  // batched per-step work taken OUT of the chains (steady-state iterations, where it does not feed the recursion)
  T* pred_mean; T* pred_var;   // [N] filter predictive H m_, H P_ H^T -> log-likelihood terms by generic_loglik_kernel
  T* post_mean; T* post_var;   // [N] smoothed marginal H m, H P H^T -> site update by site_update_kernel
  // RTS recursion in affine form (generic_gain_kernel): P_s[n] = Gt[n]^T P_s[n+1] Gt[n] + Lc[n],  m_s[n] = Gt[n]^T m_s[n+1] + gc[n]
  T* Lc; T* gc;            // [N][d][ld] (padded tile layout, like the gains) and [N][d]
  unsigned char* gscratch; // d > KJX_GEN_DMAX: per-CTA working set in global memory instead of shared memory
  size_t gscratch_stride;
};

// padded tile layout of a d x d matrix inside a kernel's working set
struct GenLayout {
  int d;     // logical size
  int ld;    // leading dimension: >= round_up(d, 4) and = 4 (mod 8)
  int mp;    // rows that fragment loads may touch: round_up(d, 16)
  int msz;   // scalars per tile (mp * ld + a tail for column over-reads of the last row)
  int w;     // values stored per row of A(dt) (GenArgs::roww): 4 or 6
};
__host__ __device__ inline GenLayout gen_layout(int d, int w = KJX_ROW_W) {
  GenLayout L;
  L.d = d; L.w = w;
  const int kp = (d + 3) & ~3;
  L.ld = kp + ((kp & 7) ? 0 : 4);
  L.mp = (d + 15) & ~15;
  L.msz = L.mp * L.ld + 16;
  return L;
}

#define KJX_GEN_ROWS(i, n) for (int i = (int)(threadIdx.x >> 5); i < (n); i += (int)(blockDim.x >> 5))
#define KJX_GEN_COLS(j, n) for (int j = (int)(threadIdx.x & 31); j < (n); j += 32)
#define KJX_GEN_VEC(i, n) for (int i = (int)threadIdx.x; i < (n); i += (int)blockDim.x)

// shared-memory carve-up helper
template <typename T> struct SmemCarver {
  unsigned char* p;
  __device__ explicit SmemCarver(unsigned char* base) : p(base) {}
  __device__ T* take(int n) { T* r = reinterpret_cast<T*>(p); p += ((size_t)n * sizeof(T) + 15) / 16 * 16; return r; }
  __device__ int* take_int(int n) { int* r = reinterpret_cast<int*>(p); p += ((size_t)n * sizeof(int) + 15) / 16 * 16; return r; }
};

// rows of A(dt) for all d rows: vals[d][KJX_ROW_W], c0n[d] = first column | nnz << 24
template <typename T>
__device__ __forceinline__ void build_rows(const DiscArgs& da, int d, int w, T dt, T* vals, int* c0n) {
  KJX_GEN_VEC(r, d) {
    int c0, nnz; T v[KJX_ROW_W] = {};
    transition_row<T>(da, r, dt, c0, nnz, v);
    c0n[r] = c0 | (nnz << 24);
#pragma unroll
    for (int q = 0; q < KJX_ROW_W; ++q) if (q < w) vals[r * w + q] = v[q];
  }
}
// all rows of all steps, once per call
template <typename T>
__global__ void __launch_bounds__(256) generic_rows_kernel(const DiscArgs da, int d, int w, int64_t N, const T* __restrict__ dt,
                                                           T* __restrict__ rows, int* __restrict__ c0n_g) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (N + 1) * d) return;
  const int64_t n = e / d; const int r = (int)(e - n * d);
  int c0, nnz; T v[KJX_ROW_W] = {};
  transition_row<T>(da, r, (n < N) ? dt[n] : T(0), c0, nnz, v);
  if (n == 0) c0n_g[r] = c0 | (nnz << 24);
#pragma unroll
  for (int q = 0; q < KJX_ROW_W; ++q) if (q < w) rows[e * w + q] = v[q];
}
// the precomputed rows of step `idx` -> vals (all threads; caller synchronises)
template <typename T>
__device__ __forceinline__ void fetch_rows(const GenArgs<T>& g, int64_t idx, T* vals) {
  const T* src = g.rows + idx * g.roww * (int64_t)g.d;
  KJX_GEN_VEC(e, g.roww * g.d) vals[e] = src[e];
}
template <typename T>
__device__ __forceinline__ void fetch_c0n(const GenArgs<T>& g, int* c0n) { KJX_GEN_VEC(r, g.d) c0n[r] = g.c0n_g[r]; }

// Element-wise pass over a d x d tile with (row = warp, column = lane): every thread first issues the loads of up to 4
// rows, then stores, so the loads are in flight together (a plain load/store loop serialises on possible aliasing).
template <typename T, class LoadF, class StoreF>
__device__ __forceinline__ void tile_map(int d, LoadF load, StoreF store) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i0 = warp; i0 < d; i0 += U * nw)
    for (int j = lane; j < d; j += 32) {
      T v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * nw; if (i < d) v[u] = load(i, j); }
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * nw; if (i < d) store(i, j, v[u]); }
    }
}
// out = A x
template <typename T>
__device__ __forceinline__ void rows_matvec(const GenLayout& L, const T* vals, const int* c0n, const T* x, T* out) {
  KJX_GEN_VEC(i, L.d) {
    const int c0 = c0n[i] & 0xffffff, nn = c0n[i] >> 24;
    T s = T(0);
    for (int q = 0; q < nn; ++q) s += vals[i * L.w + q] * x[c0 + q];
    out[i] = s;
  }
}
// Y = A (X - S)     (S may be null; ldx / lds: leading dimensions of X and S, Y uses L.ld)
template <typename T, typename TS, int W>
__device__ __forceinline__ void rows_left_w(const GenLayout& L, const T* vals, const int* c0n, const T* X, int ldx,
                                            const TS* S, int lds, T* Y) {
  constexpr int U = 4;
  const int d = L.d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i0 = warp; i0 < d; i0 += U * nw) {
    int c0[U], nn[U]; T v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nw;
      nn[u] = -1; c0[u] = 0;
      if (i < d) {
        const int c = c0n[i]; c0[u] = c & 0xffffff; nn[u] = c >> 24;
#pragma unroll
        for (int q = 0; q < W; ++q) v[u][q] = vals[i * W + q];
      }
    }
    for (int j = lane; j < d; j += 32) {
      T s[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s[u] = T(0);
        if (nn[u] >= 0) {
          const T* x = X + c0[u] * ldx + j;
#pragma unroll
          for (int q = 0; q < W; ++q)
            if (q < nn[u]) s[u] += v[u][q] * (S ? x[q * ldx] - (T)S[(c0[u] + q) * lds + j] : x[q * ldx]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) if (nn[u] >= 0) Y[(i0 + u * nw) * ld + j] = s[u];
    }
  }
}
// (the row width is uniform over the launch: one well-predicted branch; the 4-wide build keeps its registers and loads)
template <typename T, typename TS>
__device__ __forceinline__ void rows_left(const GenLayout& L, const T* vals, const int* c0n, const T* X, int ldx,
                                          const TS* S, int lds, T* Y) {
  if (L.w == 4) rows_left_w<T, TS, 4>(L, vals, c0n, X, ldx, S, lds, Y);
  else rows_left_w<T, TS, KJX_ROW_W>(L, vals, c0n, X, ldx, S, lds, Y);
}
// X <- A X in place: the rows of one diagonal block of A only depend on the same rows of X.  bstart: first rows of the
// blocks, nblk of them.
template <typename T, int W>
__device__ __forceinline__ void rows_left_inplace_w(const GenLayout& L, const T* vals, const int* c0n, const int* bstart, int nblk, T* X) {
  const int d = L.d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int b = warp; b < nblk; b += nw) {
    const int i = bstart[b], nn = c0n[i] >> 24;
    for (int j = lane; j < d; j += 32) {
      T x[W], s[W];
#pragma unroll
      for (int q = 0; q < W; ++q) x[q] = (q < nn) ? X[(i + q) * ld + j] : T(0);
#pragma unroll
      for (int r = 0; r < W; ++r) {
        s[r] = T(0);
        if (r < nn) {
#pragma unroll
          for (int q = 0; q < W; ++q) if (q < nn) s[r] += vals[(i + r) * W + q] * x[q];
        }
      }
#pragma unroll
      for (int r = 0; r < W; ++r) if (r < nn) X[(i + r) * ld + j] = s[r];
    }
  }
}
template <typename T>
__device__ __forceinline__ void rows_left_inplace(const GenLayout& L, const T* vals, const int* c0n, const int* bstart, int nblk, T* X) {
  if (L.w == 4) rows_left_inplace_w<T, 4>(L, vals, c0n, bstart, nblk, X);
  else rows_left_inplace_w<T, KJX_ROW_W>(L, vals, c0n, bstart, nblk, X);
}
// Z = sign * (Y A^T + Pinf) + (Base ? Base : 0)      (ldp: leading dimension of Pinf; Y, Z, Base use L.ld)
template <typename T, typename TP, int W>
__device__ __forceinline__ void rows_right_add_w(const GenLayout& L, const T* vals, const int* c0n, const T* Y, const TP* Pinf, int ldp,
                                                 T* Z, const T* Base, T sign) {
  constexpr int U = 4;
  const int d = L.d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = lane; j < d; j += 32) {
    const int c0 = c0n[j] & 0xffffff, nn = c0n[j] >> 24;
    T v[W];
#pragma unroll
    for (int q = 0; q < W; ++q) v[q] = vals[j * W + q];
    for (int i0 = warp; i0 < d; i0 += U * nw) {
      T s[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * nw;
        s[u] = T(0);
        if (i < d) {
          const T* y = Y + i * ld + c0;
          T t = T(0);
#pragma unroll
          for (int q = 0; q < W; ++q) if (q < nn) t += y[q] * v[q];
          t += (T)Pinf[i * ldp + j];
          s[u] = Base ? Base[i * ld + j] + sign * t : sign * t;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * nw; if (i < d) Z[i * ld + j] = s[u]; }
    }
  }
}
template <typename T, typename TP>
__device__ __forceinline__ void rows_right_add(const GenLayout& L, const T* vals, const int* c0n, const T* Y, const TP* Pinf, int ldp,
                                               T* Z, const T* Base = nullptr, T sign = T(1)) {
  if (L.w == 4) rows_right_add_w<T, TP, 4>(L, vals, c0n, Y, Pinf, ldp, Z, Base, sign);
  else rows_right_add_w<T, TP, KJX_ROW_W>(L, vals, c0n, Y, Pinf, ldp, Z, Base, sign);
}
// dense [d][d] global <-> padded tile
template <typename T, typename S>
__device__ __forceinline__ void load_dense(const GenLayout& L, const S* src, T* dst) {   // src may have been written by this kernel: no __restrict__ / read-only path
  const int d = L.d, ld = L.ld;
  tile_map<T>(d, [&](int i, int j) { return (T)src[(size_t)i * d + j]; }, [&](int i, int j, T v) { dst[i * ld + j] = v; });
}
template <typename T>
__device__ __forceinline__ void store_dense(const GenLayout& L, const T* src, T* __restrict__ dst) {
  const int d = L.d, ld = L.ld;
  tile_map<T>(d, [&](int i, int j) { return src[i * ld + j]; }, [&](int i, int j, T v) { dst[(size_t)i * d + j] = v; });
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum_i a[i] b[i] over a warp (every lane gets the result)
template <typename T> __device__ __forceinline__ T warp_dot(int d, const T* a, const T* b) {
  T s = T(0);
  KJX_GEN_COLS(i, d) s += a[i] * b[i];
  return warp_sum(s);
}
// part[w][j] = sum over the rows i = w, w + nwg, ... of h[i] X[i][j], for the nwg warps starting at warp w0 (other warps
// skip).  Zero entries of h are skipped: H is sparse for Sum priors.
template <typename T>
__device__ __forceinline__ void rowvec_partial(const GenLayout& L, const T* h, const T* X, T* part, int w0, int nwg) {
  const int w = (int)(threadIdx.x >> 5) - w0;
  if (w < 0 || w >= nwg) return;
  KJX_GEN_COLS(j, L.d) {
    T s = T(0);
    for (int i = w; i < L.d; i += nwg) { const T hi = h[i]; if (hi != T(0)) s += hi * X[i * L.ld + j]; }
    part[w * L.d + j] = s;
  }
}
template <typename T> __device__ __forceinline__ T partial_total(int d, const T* part, int j, int nwg) {
  T s = T(0);
  for (int w = 0; w < nwg; ++w) s += part[w * d + j];
  return s;
}

// ------------------------------------------------------------------------------------------------
// CTA-wide product on the FP64 tensor pipe.  C(i,j) = sum_k opA(i,k) opB(k,j) for i < m, j < n, k < kk, handed to
// epi(i, j, value); opA(i,k) = TA ? A[k*ld + i] : A[i*ld + k], opB(k,j) = TB ? B[j*ld + k] : B[k*ld + j].
// Every warp takes 16 x 16 blocks of C (2 x 2 m8n8k4 tiles: 4 fragment loads feed 4 DMMAs per k-step of 4).
// Requirements: operands zero (and finite) in the k range [kk, round_up(kk,4)); rows / columns up to the next multiple
// of 16 readable (their products are discarded).  All threads of the CTA must call; no barrier inside.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <typename T, bool TA, bool TB, class Epi>
__device__ __forceinline__ void gemm_cta(int m, int n, int kk, const T* A, const T* B, int ld, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if constexpr (std::is_same<T, double>::value) {
    const int nbn = (n + 15) >> 4, nblk = ((m + 15) >> 4) * nbn;
    const int r = lane >> 2, q = lane & 3;
    const int kp = (kk + 3) & ~3;
    const int astep = TA ? 4 * ld : 4, a8 = TA ? 8 : 8 * ld;
    const int bstep = TB ? 4 : 4 * ld, b8 = TB ? 8 * ld : 8;
    for (int blk = warp; blk < nblk; blk += nwarp) {
      const int bi = (blk / nbn) << 4, bj = (blk % nbn) << 4;
      const double* ap = TA ? A + q * ld + bi + r : A + (bi + r) * ld + q;
      const double* bp = TB ? B + (bj + r) * ld + q : B + q * ld + bj + r;
      double c00a = 0, c00b = 0, c01a = 0, c01b = 0, c10a = 0, c10b = 0, c11a = 0, c11b = 0;
#pragma unroll 2
      for (int k = 0; k < kp; k += 4) {
        const double a0 = ap[0], a1 = ap[a8], b0 = bp[0], b1 = bp[b8];
        dmma884(c00a, c00b, a0, b0);
        dmma884(c01a, c01b, a0, b1);
        dmma884(c10a, c10b, a1, b0);
        dmma884(c11a, c11b, a1, b1);
        ap += astep; bp += bstep;
      }
      __syncwarp();        // in-place use (C overlapping this block's own operands): all fragment loads done before any store
      // lane holds C[r][2q], C[r][2q+1] of each 8 x 8 tile
      const int i0 = bi + r, i1 = bi + 8 + r, j0 = bj + 2 * q, j1 = bj + 8 + 2 * q;
      if (i0 < m) {
        if (j0 < n) epi(i0, j0, c00a);
        if (j0 + 1 < n) epi(i0, j0 + 1, c00b);
        if (j1 < n) epi(i0, j1, c01a);
        if (j1 + 1 < n) epi(i0, j1 + 1, c01b);
      }
      if (i1 < m) {
        if (j0 < n) epi(i1, j0, c10a);
        if (j0 + 1 < n) epi(i1, j0 + 1, c10b);
        if (j1 < n) epi(i1, j1, c11a);
        if (j1 + 1 < n) epi(i1, j1 + 1, c11b);
      }
    }
  } else {
    for (int i = warp; i < m; i += nwarp)
      for (int j = lane; j < n; j += 32) {
        T s = T(0);
        for (int k = 0; k < kk; ++k) s += (TA ? A[k * ld + i] : A[i * ld + k]) * (TB ? B[j * ld + k] : B[k * ld + j]);
        epi(i, j, s);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// Solve M X = [R | Rb | rv] in place by blocked LU with partial pivoting (M: d x d, overwritten by its factors; R, Rb
// (optional): d x d and rv: d, overwritten by the solution).  Used on I + P J / I + J C, whose eigenvalues are real and >= 1 but which
// are not symmetric (and whose leading minors may vanish), so no Cholesky and no unpivoted elimination.
// Panels of 8 columns: warp 0 factors the panel (pivot search by shuffle, row swap, rank-one update inside the panel),
// then one thread per remaining column applies the swaps and the unit-lower 8 x 8 solve, and the rank-8 trailing
// update of [M | R] goes through gemm_cta — 8x less shared-memory traffic than rank-one elimination, which is what
// bounds that one.  Back substitution is blocked the same way.  piv: d ints, dscr: 8 scalars of scratch.
// M and R must not be the last tiles of the working set (fragment loads of the trailing update may run past the
// tile by up to 16 rows; those products are discarded).
template <typename T>
__device__ __forceinline__ void lu_solve_blocked(const GenLayout& L, T* M, T* R, T* Rb, T* rv, int* piv, T* dscr) {
  const int nrb = Rb ? L.d : 0;                             // optional second block of d right-hand sides
  constexpr int NB = 8;
  const int d = L.d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = 0; k0 < d; k0 += NB) {
    const int kb = (d - k0 < NB) ? d - k0 : NB;
    if (warp == 0 && d - k0 <= 64) {
      // the whole panel in registers: lane l owns rows k0 + l and k0 + 32 + l; row k0 + j is slot 0 of lane j
      const int i0 = k0 + lane, i1 = k0 + 32 + lane;
      const bool ok0 = i0 < d, ok1 = i1 < d;
      T a0[NB], a1[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        a0[c] = (ok0 && c < kb) ? M[i0 * ld + k0 + c] : T(0);
        a1[c] = (ok1 && c < kb) ? M[i1 * ld + k0 + c] : T(0);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < kb) {
          const int k = k0 + j;
          const T v0 = (ok0 && lane >= j) ? fabs(a0[j]) : T(-1), v1 = ok1 ? fabs(a1[j]) : T(-1);
          T best = (v1 > v0) ? v1 : v0; int bi = (v1 > v0) ? i1 : i0;
          if (!(best >= T(0))) { best = T(-1); bi = -1; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const T ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (ob > best || (ob == best && (bi < 0 || oi < bi)))) { best = ob; bi = oi; }
          }
          const int p = (bi < 0) ? k : bi;                   // column of NaNs: keep going
          if (lane == 0) piv[k] = p;
          const int lp = (p - k0) & 31; const bool sp = (p - k0) >= 32;
          T pr[NB], kr[NB];
#pragma unroll
          for (int c = 0; c < NB; ++c) { pr[c] = __shfl_sync(0xffffffffu, sp ? a1[c] : a0[c], lp); kr[c] = __shfl_sync(0xffffffffu, a0[c], j); }
          if (p != k) {
#pragma unroll
            for (int c = 0; c < NB; ++c) {
              if (lane == lp) { if (sp) a1[c] = kr[c]; else a0[c] = kr[c]; }
              if (lane == j) a0[c] = pr[c];
            }
          }
          const T pinv = T(1) / pr[j];
          if (ok0 && lane > j) {
            const T l = a0[j] * pinv; a0[j] = l;
#pragma unroll
            for (int c = 0; c < NB; ++c) if (c > j) a0[c] -= l * pr[c];
          }
          if (ok1) {
            const T l = a1[j] * pinv; a1[j] = l;
#pragma unroll
            for (int c = 0; c < NB; ++c) if (c > j) a1[c] -= l * pr[c];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (ok0 && c < kb) M[i0 * ld + k0 + c] = a0[c];
        if (ok1 && c < kb) M[i1 * ld + k0 + c] = a1[c];
      }
    } else if (warp == 0) {
      for (int j = 0; j < kb; ++j) {
        const int k = k0 + j;
        T best = T(-1); int bi = -1;
        for (int i = k + lane; i < d; i += 32) { const T v = fabs(M[i * ld + k]); if (v > best) { best = v; bi = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const T ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (oi >= 0 && (ob > best || (ob == best && (bi < 0 || oi < bi)))) { best = ob; bi = oi; }
        }
        const int p = (bi < 0) ? k : bi;                     // column of NaNs: keep going
        if (lane == 0) piv[k] = p;
        if (p != k && lane < kb) { const T t = M[k * ld + k0 + lane]; M[k * ld + k0 + lane] = M[p * ld + k0 + lane]; M[p * ld + k0 + lane] = t; }
        __syncwarp();
        const T pinv = T(1) / M[k * ld + k];
        T prow[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) prow[c] = (c > j && c < kb) ? M[k * ld + k0 + c] : T(0);
        for (int i = k + 1 + lane; i < d; i += 32) {
          T* row = M + i * ld + k0;
          T rowv[NB];
#pragma unroll
          for (int c = 0; c < NB; ++c) rowv[c] = (c >= j && c < kb) ? row[c] : T(0);
          const T l = rowv[j] * pinv;
          row[j] = l;
#pragma unroll
          for (int c = 0; c < NB; ++c) if (c > j && c < kb) row[c] = rowv[c] - l * prow[c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // swaps + U12 = L11^-1 [M12 | R1 | rv1] for every column right of the panel (one thread per column)
    const int nright = d - k0 - kb;
    for (int c = threadIdx.x; c < nright + d + nrb + 1; c += blockDim.x) {
      T* base; int st;
      if (c < nright) { base = M + k0 + kb + c; st = ld; } else if (c < nright + d) { base = R + (c - nright); st = ld; }
      else if (c < nright + d + nrb) { base = Rb + (c - nright - d); st = ld; } else { base = rv; st = 1; }
      T x[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        if (j < kb) {
          const int k = k0 + j, p = piv[k];
          T v = base[p * st];
          if (p != k) base[p * st] = base[k * st];          // row k's old value moves down; row k is rewritten below
          const T* lrow = M + k * ld + k0;
#pragma unroll
          for (int q = 0; q < j; ++q) v -= lrow[q] * x[q];
          x[j] = v;
          base[k * st] = v;
        }
      }
    }
    __syncthreads();
    if (nright > 0) {            // kb == NB here
      const T* L21 = M + (k0 + NB) * ld + k0;
      T* M22 = M + (k0 + NB) * ld + k0 + NB; T* R2 = R + (k0 + NB) * ld;
      gemm_cta<T, false, false>(nright, nright, NB, L21, M + k0 * ld + k0 + NB, ld, [&](int i, int j, T v) { M22[i * ld + j] -= v; });
      gemm_cta<T, false, false>(nright, d, NB, L21, R + k0 * ld, ld, [&](int i, int j, T v) { R2[i * ld + j] -= v; });
      if (Rb) { T* Rb2 = Rb + (k0 + NB) * ld; gemm_cta<T, false, false>(nright, d, NB, L21, Rb + k0 * ld, ld, [&](int i, int j, T v) { Rb2[i * ld + j] -= v; }); }
      for (int i = threadIdx.x; i < nright; i += blockDim.x) {
        T v = rv[k0 + NB + i];
#pragma unroll
        for (int q = 0; q < NB; ++q) v -= L21[i * ld + q] * rv[k0 + q];
        rv[k0 + NB + i] = v;
      }
      __syncthreads();
    }
  }
  // U X = (forward-substituted right-hand sides), from the last panel up
  for (int k0 = ((d - 1) / NB) * NB; k0 >= 0; k0 -= NB) {
    const int kb = (d - k0 < NB) ? d - k0 : NB;
    if (threadIdx.x < kb) dscr[threadIdx.x] = T(1) / M[(k0 + threadIdx.x) * ld + k0 + threadIdx.x];
    __syncthreads();
    for (int c = threadIdx.x; c < d + nrb + 1; c += blockDim.x) {
      T* base = (c < d) ? R + c : ((c < d + nrb) ? Rb + (c - d) : rv); const int st = (c < d + nrb) ? ld : 1;
      T x[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) x[j] = (j < kb) ? base[(k0 + j) * st] : T(0);
#pragma unroll
      for (int j = NB - 1; j >= 0; --j) {
        if (j < kb) {
          const T* urow = M + (k0 + j) * ld + k0;
          T v = x[j];
#pragma unroll
          for (int q = NB - 1; q > j; --q) if (q < kb) v -= urow[q] * x[q];
          x[j] = v * dscr[j];
        }
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) if (j < kb) base[(k0 + j) * st] = x[j];
    }
    __syncthreads();
    if (k0 > 0) {
      const T* U01 = M + k0;                                 // rows 0..k0-1, columns k0..k0+kb-1
      gemm_cta<T, false, false>(k0, d, kb, U01, R + k0 * ld, ld, [&](int i, int j, T v) { R[i * ld + j] -= v; });
      if (Rb) gemm_cta<T, false, false>(k0, d, kb, U01, Rb + k0 * ld, ld, [&](int i, int j, T v) { Rb[i * ld + j] -= v; });
      for (int i = threadIdx.x; i < k0; i += blockDim.x) {
        T v = rv[i];
        for (int q = 0; q < kb; ++q) v -= U01[i * ld + q] * rv[k0 + q];
        rv[i] = v;
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Blocked in-place Cholesky (lower) of the d x d tile W (utils.py:29).  Panels of 16 columns: warp 0 factors the 16 x 16
// diagonal block with one row per lane held in registers (right-looking, shuffles) and replaces it by its INVERSE
// (strict upper part zeroed); the panel below (L21 = A21 L11^-T) and the trailing update then go through gemm_cta.
// 3 barriers per panel.  On return the tile holds L below the diagonal blocks and L11^-1 in them, which is what
// chol_solve_blocked wants.  A non-positive pivot gives NaN, like a failed potrf propagated by jaxlib.  scr16: 16 scalars.
template <typename T> __device__ __forceinline__ void chol_blocked(const GenLayout& L, T* W, T* scr16) {
  const int d = L.d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = 0; k0 < d; k0 += 16) {
    const int kb = (d - k0 < 16) ? d - k0 : 16;
    T* D = W + k0 * ld + k0;
    if (warp == 0) {
      // lane i < kb owns row i of the block; lanes >= kb carry an identity row so the arithmetic stays finite
      T a[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) a[c] = (lane < kb && c < kb && c <= lane) ? D[lane * ld + c] : ((c == lane) ? T(1) : T(0));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const T pv = __shfl_sync(0xffffffffu, a[j], j);
        const T r = rsqrt_(pv);
        a[j] = (lane == j) ? r : a[j] * r;                           // column j of L (rows > j); lane j keeps 1 / L[j][j]
#pragma unroll
        for (int c = j + 1; c < 16; ++c) {
          const T lcj = __shfl_sync(0xffffffffu, a[j], c);             // L[c][j]
          if (lane >= c) a[c] -= a[j] * lcj;
        }
      }
      // L goes back to the tile (1 / L[j][j] to scr16), then lane c < 16 computes column c of L^-1 by forward
      // substitution with broadcast reads of L — the factor's registers are dead by then
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < lane && lane < kb) D[lane * ld + c] = a[c];
          if (c == lane) scr16[lane] = a[c];
        }
      }
      __syncwarp();
      T x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        T v = (i == lane) ? T(1) : T(0);
        if (i < kb) {
#pragma unroll
          for (int k = 0; k < i; ++k) v -= D[i * ld + k] * x[k];
        }
        x[i] = (i >= lane) ? v * scr16[i] : T(0);
      }
      __syncwarp();
      // lane c holds column c of the inverse: X[i][c] = x[i]; write the block (zero above the diagonal)
      if (lane < 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (i < kb && lane < kb) D[i * ld + lane] = (i >= lane) ? x[i] : T(0);
      }
    }
    __syncthreads();
    const int rem = d - k0 - kb;
    if (rem > 0) {               // kb == 16 here
      T* P21 = W + (k0 + 16) * ld + k0;
      // L21 = A21 L11^-T, in place (every 16 x 16 block of the product only reads its own rows of A21)
      gemm_cta<T, false, true>(rem, 16, 16, P21, D, ld, [&](int i, int j, T v) { P21[i * ld + j] = v; });
      __syncthreads();
      T* C = W + (k0 + 16) * ld + (k0 + 16);
      gemm_cta<T, false, true>(rem, rem, 16, P21, P21, ld, [&](int i, int j, T v) { C[i * ld + j] -= v; });
      __syncthreads();
    }
  }
}

// X <- solve(L L^T, X) for the ncol right-hand-side columns of X (utils.py:30) with the tile chol_blocked left behind:
// block forward / backward substitution where the diagonal solves are products with the inverted diagonal blocks (in
// place: a 16 x 16 block of the product only reads its own block of X), so everything is gemm_cta.  The two halves are
// separate calls because the gain kernel needs the intermediate Z = L^-1 X (Z^T Z = X^T (L L^T)^-1 X).
template <typename T> __device__ __forceinline__ void chol_forward_blocked(const GenLayout& L, const T* Lw, T* X, int ncol) {
  const int d = L.d, ld = L.ld;
  for (int k0 = 0; k0 < d; k0 += 16) {                        // forward: L z = b
    const int kb = (d - k0 < 16) ? d - k0 : 16;
    T* X1 = X + k0 * ld;
    gemm_cta<T, false, false>(kb, ncol, kb, Lw + k0 * ld + k0, X1, ld, [&](int i, int j, T v) { X1[i * ld + j] = v; });
    __syncthreads();
    const int rem = d - k0 - kb;
    if (rem > 0) {
      T* C = X + (k0 + 16) * ld;
      gemm_cta<T, false, false>(rem, ncol, 16, Lw + (k0 + 16) * ld + k0, X1, ld, [&](int i, int j, T v) { C[i * ld + j] -= v; });
      __syncthreads();
    }
  }
}
template <typename T> __device__ __forceinline__ void chol_backward_blocked(const GenLayout& L, const T* Lw, T* X, int ncol) {
  const int d = L.d, ld = L.ld;
  for (int k0 = ((d - 1) >> 4) << 4; k0 >= 0; k0 -= 16) {     // backward: L^T x = z
    const int kb = (d - k0 < 16) ? d - k0 : 16;
    T* X1 = X + k0 * ld;
    gemm_cta<T, true, false>(kb, ncol, kb, Lw + k0 * ld + k0, X1, ld, [&](int i, int j, T v) { X1[i * ld + j] = v; });
    __syncthreads();
    if (k0 > 0) {
      // rows [0, k0) -= L[k0.., 0..k0)^T X[k0..]   (k range kb; row d of both tiles is zero padding when kb % 4 != 0)
      gemm_cta<T, true, false>(k0, ncol, kb, Lw + k0 * ld, X1, ld, [&](int i, int j, T v) { X[i * ld + j] -= v; });
      __syncthreads();
    }
  }
}
template <typename T> __device__ __forceinline__ void chol_solve_blocked(const GenLayout& L, const T* Lw, T* X, int ncol) {
  chol_forward_blocked<T>(L, Lw, X, ncol);
  chol_backward_blocked<T>(L, Lw, X, ncol);
}

// ------------------------------------------------------------------------------------------------
// Cholesky for the gain kernel (d <= 64): panels of 8 columns, the WHOLE column panel (all rows below the diagonal
// included) in the registers of warp 0 — lane l owns rows k0 + l and k0 + 32 + l, all indices static so nothing spills
// (the 16-column version above keeps a[16] in local memory: its one-warp diagonal phases were 36 % of the gain kernel in
// the ncu source view) —, then a rank-8 update of the trailing block on the FP64 tensor pipe by all warps.  8 panels and
// 16 barriers for d = 59.  On return the strict lower triangle of W holds L and rinv[j] = 1 / L[j][j] (the diagonal of W
// holds L[j][j]); a non-positive pivot gives NaN like a failed potrf propagated by jaxlib (utils.py:29).
template <typename T> __device__ __forceinline__ void chol_panel8(const GenLayout& L, T* W, T* rinv) {
  constexpr int NB = 8;
  const int d = L.d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = 0; k0 < d; k0 += NB) {
    const int kb = (d - k0 < NB) ? d - k0 : NB;
    if (warp == 0) {
      const int i0 = k0 + lane, i1 = k0 + 32 + lane;
      const bool ok0 = i0 < d, ok1 = i1 < d;
      T a0[NB], a1[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        a0[c] = (ok0 && c < kb) ? W[i0 * ld + k0 + c] : ((c == lane) ? T(1) : T(0));      // identity padding keeps the arithmetic finite
        a1[c] = (ok1 && c < kb) ? W[i1 * ld + k0 + c] : T(0);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const T pv = __shfl_sync(0xffffffffu, a0[j], j);
        const T r = rsqrt_(pv);
        a0[j] = (lane == j) ? pv * r : a0[j] * r;            // L[j][j] = sqrt(pivot); rows below: column j of L
        a1[j] *= r;
        if (lane == j && j < kb) rinv[k0 + j] = r;
#pragma unroll
        for (int c = j + 1; c < NB; ++c) {
          const T lcj = __shfl_sync(0xffffffffu, a0[j], c);  // L[k0 + c][k0 + j]
          if (lane >= c) a0[c] -= a0[j] * lcj;
          a1[c] -= a1[j] * lcj;
        }
      }
#pragma unroll
      for (int c = 0; c < NB; ++c) {
        if (ok0 && c < kb && lane >= c) W[i0 * ld + k0 + c] = a0[c];
        if (ok1 && c < kb) W[i1 * ld + k0 + c] = a1[c];
      }
    }
    __syncthreads();
    const int rem = d - k0 - kb;
    if (rem > 0) {               // kb == NB here
      const T* L21 = W + (k0 + NB) * ld + k0;
      T* C = W + (k0 + NB) * ld + (k0 + NB);
      gemm_cta<T, false, true>(rem, rem, NB, L21, L21, ld, [&](int i, int jj, T v) { C[i * ld + jj] -= v; });
      __syncthreads();
    }
  }
}
// X <- L^-1 X (forward) / X <- L^-T X (backward) for the ncol columns of X with the factor chol_panel8 left in W:
// panels of 16 rows; inside a panel one thread per column substitutes with the panel's 16 entries in registers (L is
// read by broadcast), the rest of the rows are updated by a rank-16 product on the FP64 tensor pipe.
template <typename T> __device__ __forceinline__ void tri_forward(const GenLayout& L, const T* W, const T* rinv, T* X, int ncol) {
  const int d = L.d, ld = L.ld;
  for (int k0 = 0; k0 < d; k0 += 16) {
    const int kb = (d - k0 < 16) ? d - k0 : 16;
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
      T x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = (i < kb) ? X[(k0 + i) * ld + c] : T(0);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < kb) {
          const T z = x[i] * rinv[k0 + i];
          x[i] = z;
#pragma unroll
          for (int q = i + 1; q < 16; ++q) if (q < kb) x[q] -= W[(k0 + q) * ld + k0 + i] * z;
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < kb) X[(k0 + i) * ld + c] = x[i];
    }
    __syncthreads();
    const int rem = d - k0 - kb;
    if (rem > 0) {               // kb == 16 here
      T* C = X + (k0 + 16) * ld;
      gemm_cta<T, false, false>(rem, ncol, 16, W + (k0 + 16) * ld + k0, X + k0 * ld, ld, [&](int i, int j, T v) { C[i * ld + j] -= v; });
      __syncthreads();
    }
  }
}
template <typename T> __device__ __forceinline__ void tri_backward(const GenLayout& L, const T* W, const T* rinv, T* X, int ncol) {
  const int d = L.d, ld = L.ld;
  for (int k0 = ((d - 1) >> 4) << 4; k0 >= 0; k0 -= 16) {
    const int kb = (d - k0 < 16) ? d - k0 : 16;
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
      T x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = (i < kb) ? X[(k0 + i) * ld + c] : T(0);
#pragma unroll
      for (int i = 15; i >= 0; --i) {
        if (i < kb) {
          const T z = x[i] * rinv[k0 + i];
          x[i] = z;
#pragma unroll
          for (int q = 0; q < i; ++q) x[q] -= W[(k0 + i) * ld + k0 + q] * z;      // L^T[q][i] = L[i][q]
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < kb) X[(k0 + i) * ld + c] = x[i];
    }
    __syncthreads();
    if (k0 > 0) {
      // rows [0, k0) -= L[k0.., 0..k0)^T X[k0..]; the k range is kb: rows d.. of both tiles are zero padding
      gemm_cta<T, true, false>(k0, ncol, kb, W + k0 * ld, X + k0 * ld, ld, [&](int i, int j, T v) { X[i * ld + j] -= v; });
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// working-set sizes (bytes); the same carve-up lives in global scratch when d > KJX_GEN_DMAX
template <typename T> size_t generic_fold_smem(int d) {
  const GenLayout L = gen_layout(d);
  return (size_t)(6 * L.msz + KJX_ROW_W * d + 9 * d + KJX_GEN_WARPS * d + 16 + 32) * sizeof(T) + (size_t)(d + 4) * sizeof(int) + 256;
}
template <typename T> size_t generic_boundary_smem(int d) {
  const GenLayout L = gen_layout(d);
  return (size_t)(4 * L.msz + 7 * d + 16 + 32) * sizeof(T) + (size_t)(2 * d + 40) * sizeof(int) + 512;
}
template <typename T> size_t generic_filter_smem(int d) {
  const GenLayout L = gen_layout(d);
  return (size_t)(3 * L.msz + KJX_ROW_W * d + 5 * d + KJX_GEN_WARPS * d + 16 + 32) * sizeof(T) + (size_t)(d + 4) * sizeof(int) + 256;
}
template <typename T> size_t generic_gain_smem(int d) {
  const GenLayout L = gen_layout(d);
  return (size_t)(2 * L.msz + (KJX_ROW_W + 3) * d + 16) * sizeof(T) + (size_t)(2 * d + 8) * sizeof(int) + 256;
}
template <typename T> size_t generic_smoother_smem(int d) {
  const GenLayout L = gen_layout(d);
  return (size_t)(6 * L.msz + KJX_ROW_W * d + 9 * d + KJX_GEN_WARPS * d + 16 + 32) * sizeof(T) + (size_t)(3 * d + 44) * sizeof(int) + 512;
}

// Scan quantities are kept in EQUILIBRATED coordinates x~ = D^-1 x with D = sqrt(diag(Pinf)): the components of a Sum
// prior differ by many orders of magnitude in scale (high harmonics of a quasi-periodic component), and the solves of
// the scan would otherwise only be accurate relative to the largest component — the small ones then enter the chunks
// with errors that the smoother amplifies.  In these coordinates Pinf has a unit diagonal.  The recursions inside the
// chunks (reference order) are NOT scaled.
template <typename T> __device__ __forceinline__ void scale_vectors(int d, const double* Pinf, T* Dv, T* Dinv) {
  KJX_GEN_VEC(i, d) {
    const double v = Pinf[(size_t)i * d + i];
    const double s = (v > 0.0) ? sqrt(v) : 1.0;
    Dv[i] = (T)s; Dinv[i] = (T)(1.0 / s);
  }
}

// ------------------------------------------------------------------------------------------------
// fold: one CTA per chunk reduces its steps to the scan element (A, b, C, eta, J) of SURVEY.md 7.3
template <typename T, bool GS>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_fold_kernel(const GenArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // GS: the working set lives in per-CTA global scratch (d > 64); otherwise in shared memory — a compile-time choice so
  // that the shared-memory build addresses its tiles with LDS / STS and 32-bit offsets instead of generic 64-bit pointers
  SmemCarver<T> sc(GS ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const GenLayout L = gen_layout(g.d, g.roww);
  const int d = L.d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* A = sc.take(L.msz); T* A2 = sc.take(L.msz); T* C = sc.take(L.msz); T* Y = sc.take(L.msz); T* J = sc.take(L.msz);
  T* Pinf = sc.take(L.msz);
  T* vals = sc.take(KJX_ROW_W * d); T* b = sc.take(d); T* eta = sc.take(d); T* t1 = sc.take(d); T* Hrow = sc.take(d);
  T* av = sc.take(d); T* cv = sc.take(d); T* part = sc.take(KJX_GEN_WARPS * d); T* sh = sc.take(16);
  T* Dv = sc.take(d); T* Dinv = sc.take(d);
  int* c0n = sc.take_int(d);
  const int64_t n0 = (int64_t)blockIdx.x * g.chunk_len, n1 = (n0 + g.chunk_len < g.N) ? n0 + g.chunk_len : g.N;
  KJX_GEN_VEC(e, 6 * L.msz) A[e] = T(0);                   // the six tiles are contiguous
  scale_vectors<T>(d, g.Pinf, Dv, Dinv);
  fetch_c0n<T>(g, c0n);
  __syncthreads();
  {
    const double* Pg = g.Pinf;
    tile_map<T>(d, [&](int i, int j) { return (T)Pg[(size_t)i * d + j] * Dinv[i] * Dinv[j]; }, [&](int i, int j, T v) { Pinf[i * ld + j] = v; });
  }
  KJX_GEN_VEC(i, d) { A[i * ld + i] = T(1); b[i] = T(0); eta[i] = T(0); if (g.H) Hrow[i] = (T)g.H[i] * Dv[i]; }
  __syncthreads();
  for (int64_t n = n0; n < n1; ++n) {
    {                                                       // rows of A~ = D^-1 A D from the precomputed rows of A(dt[n])
      const int w = L.w;
      const T* src = g.rows + n * w * (int64_t)d;
      KJX_GEN_VEC(e, w * d) {
        const int r = e / w, q = e - r * w, c0 = c0n[r] & 0xffffff, nn = c0n[r] >> 24;
        vals[e] = (q < nn) ? src[e] * Dv[c0 + q] * Dinv[r] : T(0);
      }
    }
    if (g.H_steps) KJX_GEN_VEC(i, d) Hrow[i] = (T)g.H_steps[n * d + i] * Dv[i];
    __syncthreads();
    rows_left<T, T>(L, vals, c0n, A, ld, nullptr, 0, A2);   // A2 <- F A
    rows_left<T, T>(L, vals, c0n, C, ld, Pinf, ld, Y);      // Y  <- F (C - Pinf)
    rows_matvec<T>(L, vals, c0n, b, t1);                    // t1 <- F b
    __syncthreads();
    { T* t = A; A = A2; A2 = t; }
    rows_right_add<T, T>(L, vals, c0n, Y, Pinf, ld, C);     // C <- F (C - Pinf) F^T + Pinf
    KJX_GEN_VEC(i, d) b[i] = t1[i];
    __syncthreads();
    const bool masked = g.mask ? (g.mask[n] != 0) : false;
    if (!masked) {
      constexpr int HW = KJX_GEN_WARPS / 2;                 // half of the warps per product
      rowvec_partial<T>(L, Hrow, A, part, 0, HW);           // a = A^T h
      rowvec_partial<T>(L, Hrow, C, part + HW * d, HW, HW); // c = C' h (C' symmetric)
      __syncthreads();
      KJX_GEN_VEC(j, d) { av[j] = partial_total<T>(d, part, j, HW); cv[j] = partial_total<T>(d, part + HW * d, j, HW); }
      if (warp == 0) {
        T sv = T(0), beta = T(0);
        KJX_GEN_COLS(j, d) { sv += Hrow[j] * partial_total<T>(d, part + HW * d, j, HW); beta += Hrow[j] * b[j]; }
        sv = warp_sum(sv); beta = warp_sum(beta);
        if (lane == 0) {
          T yt, R;
          if (g.site_mean) { yt = g.site_mean[n]; R = g.site_var[n]; } else { yt = g.y[n]; R = T(g.site.lik_hyp); }
          sh[0] = inv1(sv + R); sh[1] = yt - beta;
        }
      }
      __syncthreads();
      const T sinv = sh[0], res = sh[1];
      for (int j = lane; j < d; j += 32) {
        const T aj = av[j], cj = cv[j];
        for (int i0 = warp; i0 < d; i0 += 4 * KJX_GEN_WARPS) {
          T vj[4], va[4], vc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * KJX_GEN_WARPS;
            if (i < d) {
              const T ai = av[i] * sinv, ki = cv[i] * sinv;
              vj[u] = J[i * ld + j] + ai * aj; va[u] = A[i * ld + j] - ki * aj; vc[u] = C[i * ld + j] - ki * cj;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * KJX_GEN_WARPS;
            if (i < d) { J[i * ld + j] = vj[u]; A[i * ld + j] = va[u]; C[i * ld + j] = vc[u]; }
          }
        }
      }
      KJX_GEN_VEC(i, d) { eta[i] += av[i] * (res * sinv); b[i] += cv[i] * sinv * res; }
    }
    __syncthreads();
  }
  T* out = g.sum + (size_t)blockIdx.x * (3 * d * d + 2 * d);
  store_dense<T>(L, A, out); store_dense<T>(L, C, out + d * d); store_dense<T>(L, J, out + 2 * d * d);
  KJX_GEN_VEC(i, d) { out[3 * d * d + i] = b[i]; out[3 * d * d + d + i] = eta[i]; }
}

// ------------------------------------------------------------------------------------------------
// boundary: from the chunk elements to the filtered state entering every chunk (walk forwards) and the information
// (eta, J) of everything after every chunk (walk backwards).  Per element one blocked-LU solve and three d^3 products,
// with the element's matrices staged into the tile that is free at that point.  A walk is a dependent chain, so for
// many chunks it is done in two levels (mode): groups of `gsz` consecutive chunks are first combined into one element
// each (generic_group_kernel, one CTA per group), a TOP walk over the group elements gives the state / information at
// the group boundaries (2 CTAs), and INNER walks fill in the chunks of every group concurrently (2 CTAs per group):
// depth ~ 3 sqrt(nchunks) instead of nchunks.  FLAT is the single-level walk (2 CTAs) for short series.
// Everything is in the equilibrated coordinates of the fold (see scale_vectors); the states entering the chunks are
// stored unscaled (the filter recursion is unscaled), the informations stay scaled (only the smoother's chunk-end
// solve reads them, and it solves in scaled coordinates too).
enum { KJX_WALK_FLAT = 0, KJX_WALK_TOP = 1, KJX_WALK_INNER = 2 };

template <typename T, bool GS>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_boundary_kernel(const GenArgs<T> g, int mode, int gsz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // GS: the working set lives in per-CTA global scratch (d > 64); otherwise in shared memory — a compile-time choice so
  // that the shared-memory build addresses its tiles with LDS / STS and 32-bit offsets instead of generic 64-bit pointers
  SmemCarver<T> sc(GS ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const GenLayout L = gen_layout(g.d, g.roww);
  const int d = L.d, dd = d * d, ld = L.ld;
  T* P = sc.take(L.msz); T* M = sc.take(L.msz); T* R = sc.take(L.msz); T* Tm = sc.take(L.msz);
  T* m = sc.take(d); T* w = sc.take(d); T* rv = sc.take(d); T* Dv = sc.take(d); T* Dinv = sc.take(d); T* dscr = sc.take(16);
  int* piv = sc.take_int(d);
  const size_t ssz = (size_t)3 * dd + 2 * d;
  const bool forward = (blockIdx.x & 1) == 0;
  const int grp = blockIdx.x >> 1;
  // the walk: `count` elements E[0..count), element k sits between output slots out(k) (before it) ... see below
  const T* E; int count, first;
  if (mode == KJX_WALK_TOP) { E = g.gsum; count = (g.nchunks + gsz - 1) / gsz; first = 0; }
  else if (mode == KJX_WALK_INNER) { first = grp * gsz; count = (g.nchunks - first < gsz) ? g.nchunks - first : gsz; E = g.sum + (size_t)first * ssz; }
  else { E = g.sum; count = g.nchunks; first = 0; }
  // chunk whose entering state (forward) / following information (backward) belongs to position k of the walk
  auto slot = [&](int k) -> int {
    if (mode == KJX_WALK_TOP) return forward ? k * gsz : (((k + 1) * gsz < g.nchunks) ? (k + 1) * gsz : g.nchunks) - 1;
    return first + k;
  };
  KJX_GEN_VEC(e, 4 * L.msz) P[e] = T(0);
  KJX_GEN_VEC(i, d) m[i] = T(0);
  scale_vectors<T>(d, g.Pinf, Dv, Dinv);
  __syncthreads();
  if (forward) {
    if (mode == KJX_WALK_INNER) {                          // start from the state the top walk left at the group's first chunk
      const T* sP = g.st_P + (size_t)first * dd; const T* sm = g.st_m + (size_t)first * d;
      tile_map<T>(d, [&](int i, int j) { return sP[(size_t)i * d + j] * Dinv[i] * Dinv[j]; }, [&](int i, int j, T v) { P[i * ld + j] = v; });
      KJX_GEN_VEC(i, d) m[i] = sm[i] * Dinv[i];
    } else {
      const double* Pg = g.Pinf;                                                     // sde_gp.py:265
      tile_map<T>(d, [&](int i, int j) { return (T)Pg[(size_t)i * d + j] * Dinv[i] * Dinv[j]; }, [&](int i, int j, T v) { P[i * ld + j] = v; });
    }
    __syncthreads();
    for (int k = 0; k < count; ++k) {
      if (!(mode == KJX_WALK_INNER && k == 0)) {
        const int c = slot(k);
        T* sP = g.st_P + (size_t)c * dd;
        if (c == 0) {                                        // the prior itself, bit for bit
          const double* Pg = g.Pinf;
          tile_map<T>(d, [&](int i, int j) { return (T)Pg[(size_t)i * d + j]; }, [&](int i, int j, T v) { sP[(size_t)i * d + j] = v; });
        } else {
          tile_map<T>(d, [&](int i, int j) { return P[i * ld + j] * Dv[i] * Dv[j]; }, [&](int i, int j, T v) { sP[(size_t)i * d + j] = v; });
        }
        KJX_GEN_VEC(i, d) g.st_m[(size_t)c * d + i] = m[i] * Dv[i];
      }
      if (k + 1 == count) break;
      const T* A = E + (size_t)k * ssz; const T* C = A + dd; const T* J = A + 2 * dd; const T* b = A + 3 * dd; const T* eta = b + d;
      load_dense<T>(L, J, Tm);
      __syncthreads();
      // (I + P J) X = [P | m + P eta]
      gemm_cta<T, false, false>(d, d, d, P, Tm, ld, [&](int i, int j, T v) { M[i * ld + j] = v + ((i == j) ? T(1) : T(0)); });
      tile_map<T>(d, [&](int i, int j) { return P[i * ld + j]; }, [&](int i, int j, T v) { R[i * ld + j] = v; });
      KJX_GEN_VEC(i, d) { T s = m[i]; for (int q = 0; q < d; ++q) s += P[i * ld + q] * eta[q]; rv[i] = s; }
      __syncthreads();
      lu_solve_blocked<T>(L, M, R, (T*)nullptr, rv, piv, dscr);               // R <- X = (I+PJ)^-1 P,  rv <- (I+PJ)^-1 (m + P eta)
      load_dense<T>(L, A, M);
      __syncthreads();
      gemm_cta<T, false, false>(d, d, d, M, R, ld, [&](int i, int j, T v) { Tm[i * ld + j] = v; });        // Tm = A X
      KJX_GEN_VEC(i, d) { T s = b[i]; for (int q = 0; q < d; ++q) s += M[i * ld + q] * rv[q]; w[i] = s; }
      __syncthreads();
      gemm_cta<T, false, true>(d, d, d, Tm, M, ld, [&](int i, int j, T v) { P[i * ld + j] = v + C[i * d + j]; });   // P' = A X A^T + C
      KJX_GEN_VEC(i, d) m[i] = w[i];
      __syncthreads();
    }
  } else {
    // information (eta, J): J in P, eta in m
    if (mode == KJX_WALK_INNER) {                          // start from the information the top walk left after the group
      const int c = first + count - 1;
      load_dense<T>(L, g.in_J + (size_t)c * dd, P);
      KJX_GEN_VEC(i, d) m[i] = g.in_eta[(size_t)c * d + i];
      __syncthreads();
    }
    for (int k = count - 1; k >= 0; --k) {
      if (!(mode == KJX_WALK_INNER && k == count - 1)) {
        const int c = slot(k);
        store_dense<T>(L, P, g.in_J + (size_t)c * dd);
        KJX_GEN_VEC(i, d) g.in_eta[(size_t)c * d + i] = m[i];
      }
      if (k == 0) break;
      const T* A = E + (size_t)k * ssz; const T* C = A + dd; const T* Ji = A + 2 * dd; const T* b = A + 3 * dd; const T* etai = b + d;
      // (I + J_j C_i) Y = [J_j A_i | eta_j - J_j b_i]
      load_dense<T>(L, C, Tm);
      KJX_GEN_VEC(i, d) { T s = m[i]; for (int q = 0; q < d; ++q) s -= P[i * ld + q] * b[q]; rv[i] = s; }
      __syncthreads();
      gemm_cta<T, false, false>(d, d, d, P, Tm, ld, [&](int i, int j, T v) { M[i * ld + j] = v + ((i == j) ? T(1) : T(0)); });
      __syncthreads();
      load_dense<T>(L, A, Tm);
      __syncthreads();
      gemm_cta<T, false, false>(d, d, d, P, Tm, ld, [&](int i, int j, T v) { R[i * ld + j] = v; });
      __syncthreads();
      lu_solve_blocked<T>(L, M, R, (T*)nullptr, rv, piv, dscr);               // R <- Y, rv <- the vector part
      gemm_cta<T, true, false>(d, d, d, Tm, R, ld, [&](int i, int j, T v) { M[i * ld + j] = v + Ji[i * d + j]; });   // J' = A_i^T Y + J_i
      KJX_GEN_VEC(i, d) { T s = etai[i]; for (int q = 0; q < d; ++q) s += Tm[q * ld + i] * rv[q]; w[i] = s; }
      __syncthreads();
      KJX_GEN_VEC(i, d) m[i] = w[i];
      tile_map<T>(d, [&](int i, int j) { return (i == j) ? M[i * ld + j] : T(0.5) * (M[i * ld + j] + M[j * ld + i]); },   // keep J' exactly symmetric
                  [&](int i, int j, T v) { P[i * ld + j] = v; });
      __syncthreads();
    }
  }
}

// group elements: CTA `grp` combines the elements of its gsz consecutive chunks, earliest first, into gsum[grp]
// (the accumulator lives there).  acc (x) e  (Sarkka & Garcia-Fernandez 2021, Lemma 8) with
// M = I + C_a J_e and ONE solve M [X_A | X_C | x_b] = [A_a | C_a | b_a + C_a eta_e]:
//   A = A_e X_A,  b = A_e x_b + b_e,  C = A_e X_C A_e^T + C_e,
//   J = A_a^T J_e X_A + J_a                      (push-through: M^-T J_e = J_e M^-1),
//   eta = A_a^T (v - J_e X_C v) + eta_a,  v = eta_e - J_e b_a      (M^-T = I - J_e M^-1 C_a).
template <typename T> size_t generic_group_smem(int d) {
  const GenLayout L = gen_layout(d);
  return (size_t)(5 * L.msz + 8 * d + 16 + 32) * sizeof(T) + (size_t)(d + 8) * sizeof(int) + 512;
}
template <typename T, bool GS>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_group_kernel(const GenArgs<T> g, int gsz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // GS: the working set lives in per-CTA global scratch (d > 64); otherwise in shared memory — a compile-time choice so
  // that the shared-memory build addresses its tiles with LDS / STS and 32-bit offsets instead of generic 64-bit pointers
  SmemCarver<T> sc(GS ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const GenLayout L = gen_layout(g.d, g.roww);
  const int d = L.d, dd = d * d, ld = L.ld;
  T* T0 = sc.take(L.msz); T* T1 = sc.take(L.msz); T* T2 = sc.take(L.msz); T* T3 = sc.take(L.msz); T* T4 = sc.take(L.msz);
  T* rv = sc.take(d); T* v = sc.take(d); T* w1 = sc.take(d); T* w2 = sc.take(d); T* bn = sc.take(d); T* en = sc.take(d); T* dscr = sc.take(16);
  int* piv = sc.take_int(d);
  const size_t ssz = (size_t)3 * dd + 2 * d;
  const int first = blockIdx.x * gsz, count = (g.nchunks - first < gsz) ? g.nchunks - first : gsz;
  T* acc = g.gsum + (size_t)blockIdx.x * ssz;
  KJX_GEN_VEC(e, 5 * L.msz) T0[e] = T(0);
  for (size_t e = threadIdx.x; e < ssz; e += blockDim.x) acc[e] = g.sum[(size_t)first * ssz + e];
  __syncthreads();
  T* Aa = acc; T* Ca = acc + dd; T* Ja = acc + 2 * dd; T* ba = acc + 3 * dd; T* ea = ba + d;
  for (int k = 1; k < count; ++k) {
    const T* Ae = g.sum + (size_t)(first + k) * ssz; const T* Ce = Ae + dd; const T* Je = Ae + 2 * dd; const T* be = Ae + 3 * dd; const T* ee = be + d;
    load_dense<T>(L, Aa, T1); load_dense<T>(L, Ca, T2); load_dense<T>(L, Je, T3);
    __syncthreads();
    gemm_cta<T, false, false>(d, d, d, T2, T3, ld, [&](int i, int j, T x) { T0[i * ld + j] = x + ((i == j) ? T(1) : T(0)); });   // M = I + C_a J_e
    KJX_GEN_VEC(i, d) {
      T s = ba[i], t = ee[i];
      for (int q = 0; q < d; ++q) { s += T2[i * ld + q] * ee[q]; t -= T3[i * ld + q] * ba[q]; }
      rv[i] = s; v[i] = t;                                   // b_a + C_a eta_e ;  v = eta_e - J_e b_a
    }
    __syncthreads();
    lu_solve_blocked<T>(L, T0, T1, T2, rv, piv, dscr);       // T1 <- X_A, T2 <- X_C, rv <- x_b
    gemm_cta<T, false, false>(d, d, d, T3, T1, ld, [&](int i, int j, T x) { T4[i * ld + j] = x; });          // T4 = J_e X_A
    KJX_GEN_VEC(i, d) { T s = T(0); for (int q = 0; q < d; ++q) s += T2[i * ld + q] * v[q]; w1[i] = s; }     // X_C v
    load_dense<T>(L, Aa, T0);
    __syncthreads();
    KJX_GEN_VEC(i, d) { T s = v[i]; for (int q = 0; q < d; ++q) s -= T3[i * ld + q] * w1[q]; w2[i] = s; }    // M^-T v
    __syncthreads();
    gemm_cta<T, true, false>(d, d, d, T0, T4, ld, [&](int i, int j, T x) { T3[i * ld + j] = x + Ja[i * d + j]; });   // J = A_a^T J_e X_A + J_a
    KJX_GEN_VEC(i, d) { T s = ea[i]; for (int q = 0; q < d; ++q) s += T0[q * ld + i] * w2[q]; en[i] = s; }
    __syncthreads();
    tile_map<T>(d, [&](int i, int j) { return (i == j) ? T3[i * ld + j] : T(0.5) * (T3[i * ld + j] + T3[j * ld + i]); },
                [&](int i, int j, T x) { Ja[(size_t)i * d + j] = x; });
    load_dense<T>(L, Ae, T0);
    __syncthreads();
    gemm_cta<T, false, false>(d, d, d, T0, T1, ld, [&](int i, int j, T x) { Aa[(size_t)i * d + j] = x; });   // A = A_e X_A
    gemm_cta<T, false, false>(d, d, d, T0, T2, ld, [&](int i, int j, T x) { T4[i * ld + j] = x; });          // T4 = A_e X_C
    KJX_GEN_VEC(i, d) { T s = be[i]; for (int q = 0; q < d; ++q) s += T0[i * ld + q] * rv[q]; bn[i] = s; }
    __syncthreads();
    gemm_cta<T, false, true>(d, d, d, T4, T0, ld, [&](int i, int j, T x) { Ca[(size_t)i * d + j] = x + Ce[i * d + j]; });   // C = A_e X_C A_e^T + C_e
    KJX_GEN_VEC(i, d) { ba[i] = bn[i]; ea[i] = en[i]; }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// log Z of every step from the recorded predictive moments (sde_gp.py:287, 300-301), embarrassingly parallel; one partial
// sum per block (fixed order), reduced by nlml_reduce_kernel
template <typename T>
__global__ void __launch_bounds__(128) generic_loglik_kernel(const SiteModel site, int64_t N, const T* __restrict__ y,
                                                             const uint8_t* __restrict__ mask, const T* __restrict__ pm,
                                                             const T* __restrict__ pv, T* __restrict__ partial) {
  __shared__ T red[4];
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  T lml = T(0);
  if (n < N && !(mask && mask[n] != 0)) lml = filter_loglik<T>(site, y[n], pm[n], pv[n]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lml += __shfl_down_sync(0xffffffffu, lml, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lml;
  __syncthreads();
  if (threadIdx.x == 0) partial[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
}

// ------------------------------------------------------------------------------------------------
// forward filter, sde_gp.py:271-306
template <typename T, bool GS>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_filter_kernel(const GenArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // GS: the working set lives in per-CTA global scratch (d > 64); otherwise in shared memory — a compile-time choice so
  // that the shared-memory build addresses its tiles with LDS / STS and 32-bit offsets instead of generic 64-bit pointers
  SmemCarver<T> sc(GS ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const GenLayout L = gen_layout(g.d, g.roww);
  const int d = L.d, dd = d * d, ld = L.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* P = sc.take(L.msz); T* Y = sc.take(L.msz); T* Pinf = sc.take(L.msz);
  T* vals = sc.take(KJX_ROW_W * d); T* m = sc.take(d); T* mp = sc.take(d); T* Hrow = sc.take(d); T* HP = sc.take(d);
  T* part = sc.take(KJX_GEN_WARPS * d);
  T* sh = sc.take(16);                 // scalars: 0 pm, 2 site_mean, 4 Sinv, 5 masked, 6 nlml
  int* c0n = sc.take_int(d);
  // this CTA's chunk and the filtered state entering it (chunk 0: the prior, sde_gp.py:265)
  const int64_t n0 = (int64_t)blockIdx.x * g.chunk_len, n1 = (n0 + g.chunk_len < g.N) ? n0 + g.chunk_len : g.N;
  KJX_GEN_VEC(e, 3 * L.msz) P[e] = T(0);
  fetch_c0n<T>(g, c0n);
  __syncthreads();
  load_dense<T>(L, g.Pinf, Pinf);
  load_dense<T>(L, g.st_P + (size_t)blockIdx.x * dd, P);
  KJX_GEN_VEC(i, d) { m[i] = g.st_m[(size_t)blockIdx.x * d + i]; if (g.H) Hrow[i] = (T)g.H[i]; }
  if (threadIdx.x == 0) sh[6] = T(0);
  __syncthreads();
  for (int64_t n = n0; n < n1; ++n) {
    fetch_rows<T>(g, n, vals);                                                      // :276
    if (g.H_steps) KJX_GEN_VEC(i, d) Hrow[i] = (T)g.H_steps[n * d + i];             // :282
    __syncthreads();
    rows_matvec<T>(L, vals, c0n, m, mp);                                            // :277
    rows_left<T, T>(L, vals, c0n, P, ld, Pinf, ld, Y);                              // Y <- A (P - Pinf)
    __syncthreads();
    rows_right_add<T, T>(L, vals, c0n, Y, Pinf, ld, P);                             // P <- P_ = . A^T + Pinf   (:278)
    __syncthreads();
    rowvec_partial<T>(L, Hrow, P, part, 0, KJX_GEN_WARPS);                                            // HP = H P_
    __syncthreads();
    KJX_GEN_VEC(j, d) HP[j] = partial_total<T>(d, part, j, KJX_GEN_WARPS);
    if (warp == 0) {                   // warp 0: the step's scalar work; cubature points spread over its lanes
      T pm = T(0), pv = T(0);
      KJX_GEN_COLS(j, d) { pm += Hrow[j] * mp[j]; pv += partial_total<T>(d, part, j, KJX_GEN_WARPS) * Hrow[j]; }   // :283-284
      pm = warp_sum(pm); pv = warp_sum(pv);
      const bool masked = g.mask ? (g.mask[n] != 0) : false;
      T smean, svar, lml = T(0);
      if (g.init_sites) {
        const SiteOut<T> o = site_update<T, WarpRed>(g.site, masked ? pm : g.y[n], pm, pv, false, T(0), T(0));   // :286-287
        smean = o.mean; svar = o.var; lml = o.lml;
      } else {
        if (g.site_mean) { smean = g.site_mean[n]; svar = g.site_var[n]; }           // :290
        else { smean = g.y[n]; svar = T(g.site.lik_hyp); }                           // Gaussian + EP initialisation
        // log Z is always the true likelihood's (:287).  It does not feed the recursion, so steady-state iterations
        // only record the predictive here and generic_loglik_kernel evaluates all steps at once afterwards.
        if (g.pred_mean) { if (lane == 0) { g.pred_mean[n] = pm; g.pred_var[n] = pv; } }
        else if (!masked) lml = filter_loglik<T, WarpRed>(g.site, g.y[n], pm, pv);
      }
      if (masked) lml = T(0);                                                        // :300
      if (lane == 0) {
        sh[0] = pm; sh[2] = smean; sh[4] = inv1(pv + svar); sh[5] = masked ? T(1) : T(0);   // :292-294
        sh[6] += lml;                                                                // :301
        if (g.out_site_mean) { g.out_site_mean[n] = smean; g.out_site_var[n] = svar; }
      }
    }
    __syncthreads();
    const bool masked = sh[5] != T(0);
    const T Sinv = sh[4], innov = sh[2] - sh[0];
    KJX_GEN_VEC(i, d) {
      const T mi = masked ? mp[i] : mp[i] + HP[i] * Sinv * innov;                    // :295 / :298
      m[i] = mi;
      if (g.filt_mean) g.filt_mean[n * d + i] = mi;
    }
    T* fc = g.filt_cov ? g.filt_cov + n * (int64_t)dd : nullptr;
    if (!masked || fc)
      tile_map<T>(d, [&](int i, int j) { return masked ? P[i * ld + j] : P[i * ld + j] - HP[i] * Sinv * HP[j]; },    // :296 / :299
                  [&](int i, int j, T v) { P[i * ld + j] = v; if (fc) fc[i * d + j] = v; });
    __syncthreads();
  }
  if (threadIdx.x == 0) g.chunk_nlml[blockIdx.x] = sh[6];
}

// ------------------------------------------------------------------------------------------------
// The same filter for priors with at most 16 diagonal blocks (every Sum of Matern / quasi-periodic components up to
// d = 64): P lives in REGISTERS, thread (bi, bj) owning the (block bi, block bj) tile of at most 4 x 4 entries.  A is
// block diagonal, so the prediction A (P - Pinf) A^T + Pinf is local to a thread — no pass over shared memory and no
// barrier — and only the row vector H P needs a reduction over block rows.  Three barriers per step instead of six and
// none of the tile traffic that bounds the kernel above; bit-compatible with it up to the order of that one reduction.
template <typename T> size_t generic_filter_blocks_smem(int d) {
  return (size_t)(KJX_ROW_W * d + 4 * d + 16 * d + 16) * sizeof(T) + (size_t)(d + 40) * sizeof(int) + 256;
}
// BS: the largest diagonal block of A (4, or 6 with a Subband Matern-5/2 component): size of the register tiles
template <typename T, int BS>
__global__ void __launch_bounds__(KJX_GEN_THREADS, 1) generic_filter_blocks_kernel(const GenArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(smem_raw);
  const GenLayout L = gen_layout(g.d, g.roww);
  const int d = g.d, dd = d * d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* vals = sc.take(KJX_ROW_W * d); T* m = sc.take(d); T* mp = sc.take(d); T* Hrow = sc.take(d); T* HP = sc.take(d);
  T* part = sc.take(16 * d);
  T* sh = sc.take(16);                 // scalars: 0 pm, 2 site_mean, 4 Sinv, 5 masked, 6 nlml
  int* c0n = sc.take_int(d); int* bstart = sc.take_int(17); int* bsize = sc.take_int(17);
  const int64_t n0 = (int64_t)blockIdx.x * g.chunk_len, n1 = (n0 + g.chunk_len < g.N) ? n0 + g.chunk_len : g.N;
  fetch_c0n<T>(g, c0n);                                      // block structure of A
  __syncthreads();
  if (threadIdx.x == 0) {
    int nb = 0;
    for (int i = 0; i < d; ++i) if ((c0n[i] & 0xffffff) == i) { bstart[nb] = i; bsize[nb] = c0n[i] >> 24; ++nb; }
    bstart[16] = nb; sh[6] = T(0);
  }
  KJX_GEN_VEC(i, d) { m[i] = g.st_m[(size_t)blockIdx.x * d + i]; if (g.H) Hrow[i] = (T)g.H[i]; }
  __syncthreads();
  const int nblk = bstart[16];
  const bool owner = (int)threadIdx.x < nblk * nblk;
  const int bi = owner ? (int)threadIdx.x / nblk : 0, bj = owner ? (int)threadIdx.x % nblk : 0;
  const int r0 = bstart[bi], ni = owner ? bsize[bi] : 0, c0 = bstart[bj], nj = owner ? bsize[bj] : 0;
  constexpr int RW = (BS > 4) ? KJX_ROW_W : 4;     // the host launches BS = 6 exactly when the rows are 6 wide
  T Pb[BS][BS], Qb[BS][BS];
  {
    const T* sP = g.st_P + (size_t)blockIdx.x * dd;
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
      for (int b = 0; b < BS; ++b) {
        const bool in = a < ni && b < nj;
        Pb[a][b] = in ? sP[(size_t)(r0 + a) * d + c0 + b] : T(0);
        Qb[a][b] = in ? (T)g.Pinf[(size_t)(r0 + a) * d + c0 + b] : T(0);
      }
  }
  for (int64_t n = n0; n < n1; ++n) {
    fetch_rows<T>(g, n, vals);                                                      // :276
    if (g.H_steps) KJX_GEN_VEC(i, d) Hrow[i] = (T)g.H_steps[n * d + i];             // :282
    __syncthreads();
    rows_matvec<T>(L, vals, c0n, m, mp);                                            // :277
    if (owner) {
      T Fi[BS][BS], Fj[BS][BS], Y[BS][BS];
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int q = 0; q < BS; ++q) { Fi[a][q] = (a < ni) ? vals[(r0 + a) * RW + q] : T(0); Fj[a][q] = (a < nj) ? vals[(c0 + a) * RW + q] : T(0); }
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) {                                               // Y = A (P - Pinf)
          T s = T(0);
#pragma unroll
          for (int q = 0; q < BS; ++q) if (q < ni) s += Fi[a][q] * (Pb[q][b] - Qb[q][b]);
          Y[a][b] = s;
        }
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) {                                               // P_ = Y A^T + Pinf        (:278)
          T t = T(0);
#pragma unroll
          for (int q = 0; q < BS; ++q) if (q < nj) t += Y[a][q] * Fj[b][q];
          Pb[a][b] = (a < ni && b < nj) ? t + Qb[a][b] : T(0);
        }
#pragma unroll
      for (int b = 0; b < BS; ++b) {                                                 // this block row's share of H P_
        if (b < nj) {
          T s = T(0);
#pragma unroll
          for (int a = 0; a < BS; ++a) if (a < ni) s += Hrow[r0 + a] * Pb[a][b];
          part[bi * d + c0 + b] = s;
        }
      }
    }
    __syncthreads();
    KJX_GEN_VEC(j, d) HP[j] = partial_total<T>(d, part, j, nblk);
    if (warp == 0) {                   // warp 0: the step's scalar work; cubature points spread over its lanes
      T pm = T(0), pv = T(0);
      KJX_GEN_COLS(j, d) { pm += Hrow[j] * mp[j]; pv += partial_total<T>(d, part, j, nblk) * Hrow[j]; }   // :283-284
      pm = warp_sum(pm); pv = warp_sum(pv);
      const bool masked = g.mask ? (g.mask[n] != 0) : false;
      T smean, svar, lml = T(0);
      if (g.init_sites) {
        const SiteOut<T> o = site_update<T, WarpRed>(g.site, masked ? pm : g.y[n], pm, pv, false, T(0), T(0));   // :286-287
        smean = o.mean; svar = o.var; lml = o.lml;
      } else {
        if (g.site_mean) { smean = g.site_mean[n]; svar = g.site_var[n]; }           // :290
        else { smean = g.y[n]; svar = T(g.site.lik_hyp); }                           // Gaussian + EP initialisation
        // log Z is always the true likelihood's (:287).  It does not feed the recursion, so steady-state iterations
        // only record the predictive here and generic_loglik_kernel evaluates all steps at once afterwards.
        if (g.pred_mean) { if (lane == 0) { g.pred_mean[n] = pm; g.pred_var[n] = pv; } }
        else if (!masked) lml = filter_loglik<T, WarpRed>(g.site, g.y[n], pm, pv);
      }
      if (masked) lml = T(0);                                                        // :300
      if (lane == 0) {
        sh[0] = pm; sh[2] = smean; sh[4] = inv1(pv + svar); sh[5] = masked ? T(1) : T(0);   // :292-294
        sh[6] += lml;                                                                // :301
        if (g.out_site_mean) { g.out_site_mean[n] = smean; g.out_site_var[n] = svar; }
      }
    }
    __syncthreads();
    const bool masked = sh[5] != T(0);
    const T Sinv = sh[4], innov = sh[2] - sh[0];
    KJX_GEN_VEC(i, d) {
      const T mi = masked ? mp[i] : mp[i] + HP[i] * Sinv * innov;                    // :295 / :298
      m[i] = mi;
      if (g.filt_mean) g.filt_mean[n * d + i] = mi;
    }
    if (owner) {
      T* fc = g.filt_cov ? g.filt_cov + n * (int64_t)dd : nullptr;
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) {
          if (a < ni && b < nj) {
            if (!masked) Pb[a][b] = Pb[a][b] - HP[r0 + a] * Sinv * HP[c0 + b];       // :296 / :299
            if (fc) fc[(size_t)(r0 + a) * d + c0 + b] = Pb[a][b];
          }
        }
    }
    // (the next step's first barrier separates these reads of HP / sh / mp from their next writes)
  }
  __syncthreads();
  if (threadIdx.x == 0) g.chunk_nlml[blockIdx.x] = sh[6];
}

// ------------------------------------------------------------------------------------------------
// RTS gains, embarrassingly parallel over time: Gt[n] = solve(P_pred[n], A[n+1] Pf[n]) depends only on the FILTERED
// moments (sde_gp.py:351-359), so a grid of CTAs computes them all before the (sequential) backward recursion — and with
// them the constant terms of that recursion written in affine form,
//     P_s[n] = Pf + G (P_s[n+1] - P_pred) G^T = G P_s[n+1] G^T + Lc,   Lc = Pf - G P_pred G^T = Pf - Z^T Z,
//     m_s[n] = mf + G (m_s[n+1] - A mf)       = G m_s[n+1] + gc,       gc = mf - G A mf,
// with G = Gt^T and Z = L^-1 A Pf the intermediate of the Cholesky solve (P_pred = L L^T): the chain kernel is left with
// two products per step.  Gt and Lc go out in the padded tile layout ([d][ld]: 16-byte rows) so that the smoother can
// fetch each with one bulk copy.  Working set: two tiles in shared memory (61 KB at d = 59, three CTAs per SM).
template <typename T> __host__ __device__ inline int gen_vec_stride(int d) { return (d + 3) & ~3; }   // 16-byte multiple for both types
// RW: row width of A(dt) as a compile-time constant (4 or KJX_ROW_W = g.roww): this kernel sits at its register limit,
// and carrying both row-product variants behind a run-time branch cost 16 % (6.8 against 5.8 ms at C3).
template <typename T, bool GS, int RW>
__global__ void __launch_bounds__(KJX_GEN_GAIN_THREADS, 3) generic_gain_kernel(const GenArgs<T> g, T* __restrict__ Gt_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // GS: the working set lives in per-CTA global scratch (d > 64); otherwise in shared memory — a compile-time choice so
  // that the shared-memory build addresses its tiles with LDS / STS and 32-bit offsets instead of generic 64-bit pointers
  SmemCarver<T> sc(GS ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const GenLayout L = gen_layout(g.d, g.roww);
  const int d = L.d, dd = d * d, ld = L.ld, tsz = d * ld, gs = gen_vec_stride<T>(d), lane = threadIdx.x & 31;
  T* W = sc.take(L.msz); T* Gt = sc.take(L.msz);
  T* vals = sc.take(KJX_ROW_W * d); T* mpv = sc.take(d); T* rinv = sc.take(d); T* scr16 = sc.take(16);
  int* c0n = sc.take_int(d); int* bstart = sc.take_int(d + 1);
  KJX_GEN_VEC(e, 2 * L.msz) W[e] = T(0);
  fetch_c0n<T>(g, c0n);                                      // block structure of A: first rows of the diagonal blocks
  __syncthreads();
  if (threadIdx.x == 0) { int nb = 0; for (int i = 0; i < d; ++i) if ((c0n[i] & 0xffffff) == i) bstart[nb++] = i; bstart[d] = nb; }
  __syncthreads();
  const int nblk = bstart[d];
  const double* Pinf = g.Pinf;
  for (int64_t n = blockIdx.x; n < g.N; n += gridDim.x) {
    fetch_rows<T>(g, n + 1, vals);                                                  // A(dt[n+1]), A = I after the last step: :337, :351
    const T* Pf = g.filt_cov_in + n * (int64_t)dd;
    const T* mf = g.filt_mean_in + n * (int64_t)d;
    tile_map<T>(d, [&](int i, int j) { return Pf[i * d + j] - (T)Pinf[i * d + j]; }, [&](int i, int j, T v) { W[i * ld + j] = v; });
    __syncthreads();
    rows_left_w<T, T, RW>(L, vals, c0n, W, ld, nullptr, 0, Gt);                     // Gt <- A (Pf - Pinf)
    rows_matvec<T>(L, vals, c0n, mf, mpv);                                          // A mf                      (:352)
    __syncthreads();
    rows_right_add_w<T, double, RW>(L, vals, c0n, Gt, Pinf, d, W, nullptr, T(1));   // W <- P_pred = . A^T + Pinf   (:354)
    __syncthreads();
    load_dense<T>(L, Pf, Gt);
    __syncthreads();
    rows_left_inplace_w<T, RW>(L, vals, c0n, bstart, nblk, Gt);                     // Gt <- A Pf                  (:353)
    __syncthreads();
    // utils.py:29; Gt <- Z = L^-1 A Pf.  Shared-memory build (d <= 64): the register-panel factorisation (64 rows)
    if constexpr (!GS) { chol_panel8<T>(L, W, rinv); tri_forward<T>(L, W, rinv, Gt, d); }
    else { chol_blocked<T>(L, W, scr16); chol_forward_blocked<T>(L, W, Gt, d); }
    if (g.Lc) {
      T* Lo = g.Lc + n * (int64_t)tsz;
      gemm_cta<T, true, false>(d, d, d, Gt, Gt, ld, [&](int i, int j, T v) { Lo[i * ld + j] = Pf[i * d + j] - v; });   // Lc = Pf - Z^T Z
      KJX_GEN_ROWS(i, d) for (int j = d + lane; j < ld; j += 32) Lo[i * ld + j] = T(0);   // padding columns: the smoother's k loops read them
      __syncthreads();
    }
    if constexpr (!GS) tri_backward<T>(L, W, rinv, Gt, d); else chol_backward_blocked<T>(L, W, Gt, d);   // Gt <- L^-T Z   (:359, utils.py:30)
    T* Go = Gt_out + n * (int64_t)tsz;
    KJX_GEN_ROWS(i, d) KJX_GEN_COLS(j, ld) Go[i * ld + j] = Gt[i * ld + j];             // whole padded rows (padding stays zero)
    if (g.gc) KJX_GEN_VEC(i, d) { T v = mf[i]; for (int k = 0; k < d; ++k) v -= Gt[k * ld + i] * mpv[k]; g.gc[n * (int64_t)gs + i] = v; }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// bulk copies global -> shared through the TMA unit with mbarrier completion (cp.async.bulk; SASS UBLKCP)
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "KJX_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra KJX_MBAR_DONE;\n"
      "bra KJX_MBAR_WAIT;\n"
      "KJX_MBAR_DONE:\n"
      "}\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                  "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// backward smoother, sde_gp.py:337-379.  Last step of the chunk: filtered x information of everything after the chunk
// (two-filter identity; the chunk that ends the series: smoothed = filtered, sde_gp.py:339).  Every other step: the
// affine recursion with the terms generic_gain_kernel prepared — two d^3 products on the FP64 tensor pipe per step, the
// next step's (Gt, Lc, gc) arriving by bulk copy (TMA, mbarrier) while the current one is being multiplied.  The marginal
// H m, H P H^T of every step goes to post_mean / post_var; the site update is a batched kernel afterwards.
template <typename T, bool GS>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_smoother_kernel(const GenArgs<T> g, const T* __restrict__ Gt_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2];
  // GS: the working set lives in per-CTA global scratch (d > 64); otherwise in shared memory — a compile-time choice so
  // that the shared-memory build addresses its tiles with LDS / STS and 32-bit offsets instead of generic 64-bit pointers
  SmemCarver<T> sc(GS ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const GenLayout L = gen_layout(g.d, g.roww);
  const int d = L.d, dd = d * d, ld = L.ld, tsz = d * ld, gs = gen_vec_stride<T>(d), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* Ps = sc.take(L.msz); T* T1 = sc.take(L.msz);
  T* GtS[2]; T* LcS[2]; T* gvS[2];
  GtS[0] = sc.take(L.msz); GtS[1] = sc.take(L.msz); LcS[0] = sc.take(L.msz); LcS[1] = sc.take(L.msz);
  gvS[0] = sc.take(gs); gvS[1] = sc.take(gs);
  T* ms = sc.take(d); T* mnew = sc.take(d); T* mf = sc.take(d); T* Hrow = sc.take(d);
  T* rv = sc.take(d); T* part = sc.take(KJX_GEN_WARPS * d); T* Dv = sc.take(d); T* Dinv = sc.take(d); T* dscr = sc.take(16);
  int* piv = sc.take_int(d);
  constexpr bool tma = !GS;                                // tiles in shared memory: bulk copies; in global scratch: plain copies
  const int64_t n0 = (int64_t)blockIdx.x * g.chunk_len, n1 = (n0 + g.chunk_len < g.N) ? n0 + g.chunk_len : g.N;
  KJX_GEN_VEC(e, 6 * L.msz) Ps[e] = T(0);
  scale_vectors<T>(d, g.Pinf, Dv, Dinv);
  if (threadIdx.x == 0 && tma) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  KJX_GEN_VEC(i, d) if (g.H) Hrow[i] = (T)g.H[i];
  __syncthreads();
  // marginal of the function value of step n from (ms, Ps): H m, H P H^T                 (:368-369, :373)
  auto project = [&](int64_t n) {
    rowvec_partial<T>(L, Hrow, Ps, part, 0, KJX_GEN_WARPS);
    __syncthreads();
    if (warp == 0) {
      T pm = T(0), pv = T(0);
      KJX_GEN_COLS(j, d) { pm += Hrow[j] * ms[j]; pv += partial_total<T>(d, part, j, KJX_GEN_WARPS) * Hrow[j]; }
      pm = warp_sum(pm); pv = warp_sum(pv);
      if (lane == 0) { g.post_mean[n] = pm; g.post_var[n] = pv; }
    }
  };
  {
    const int64_t n = n1 - 1;
    T* Mm = GtS[0]; T* Rr = GtS[1]; T* Pf = LcS[0]; T* Wj = LcS[1];           // scratch roles of this step (the LU's
    // operands must not be the last tiles: its fragment loads may read up to 16 rows past a tile)
    if (g.H_steps) KJX_GEN_VEC(i, d) Hrow[i] = (T)g.H_steps[n * d + i];
    const T* Pg = g.filt_cov_in + n * (int64_t)dd;
    if (n1 == g.N) {
      load_dense<T>(L, Pg, Ps);
      KJX_GEN_VEC(i, d) ms[i] = g.filt_mean_in[n * d + i];
      __syncthreads();
    } else {
      // (I + P~ J~) X~ = [P~ | m~ + P~ eta~] in the equilibrated coordinates of the scan (see scale_vectors)
      tile_map<T>(d, [&](int i, int j) { return Pg[(size_t)i * d + j] * Dinv[i] * Dinv[j]; }, [&](int i, int j, T v) { Pf[i * ld + j] = v; });
      load_dense<T>(L, g.in_J + (size_t)blockIdx.x * dd, Wj);
      KJX_GEN_VEC(i, d) mf[i] = g.filt_mean_in[n * d + i] * Dinv[i];
      __syncthreads();
      const T* eta = g.in_eta + (size_t)blockIdx.x * d;
      gemm_cta<T, false, false>(d, d, d, Pf, Wj, ld, [&](int i, int j, T v) { Mm[i * ld + j] = v + ((i == j) ? T(1) : T(0)); });
      tile_map<T>(d, [&](int i, int j) { return Pf[i * ld + j]; }, [&](int i, int j, T v) { Rr[i * ld + j] = v; });
      KJX_GEN_VEC(i, d) { T s = mf[i]; for (int k = 0; k < d; ++k) s += Pf[i * ld + k] * eta[k]; rv[i] = s; }
      __syncthreads();
      lu_solve_blocked<T>(L, Mm, Rr, (T*)nullptr, rv, piv, dscr);         // Rr <- (I + P J)^-1 P,  rv <- (I + P J)^-1 (m + P eta)
      tile_map<T>(d, [&](int i, int j) { return ((i <= j) ? Rr[i * ld + j] : Rr[j * ld + i]) * Dv[i] * Dv[j]; },   // symmetric by construction; keep the upper part
                  [&](int i, int j, T v) { Ps[i * ld + j] = v; });
      KJX_GEN_VEC(i, d) ms[i] = rv[i] * Dv[i];
      __syncthreads();
    }
    if (g.smooth_mean && g.return_full) {
      store_dense<T>(L, Ps, g.smooth_cov + n * (int64_t)dd);
      KJX_GEN_VEC(i, d) g.smooth_mean[n * d + i] = ms[i];
    }
    project(n);
    __syncthreads();
    // the scratch tiles go back to zero padding (rows d.. and columns d.. must be finite zeros for the fragment loads)
    KJX_GEN_VEC(e, 4 * L.msz) GtS[0][e] = T(0);
    __syncthreads();
  }
  const unsigned tbytes = (unsigned)(tsz * sizeof(T)), gbytes = (unsigned)(gs * sizeof(T));
  auto fetch = [&](int64_t n, int st) {                    // (Gt, Lc, gc) of step n -> stage st
    if (tma) {
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the stage's last generic-proxy accesses are ordered before the copies
        mbar_expect_tx(&bars[st], 2 * tbytes + gbytes);
        bulk_load(GtS[st], Gt_in + n * (int64_t)tsz, tbytes, &bars[st]);
        bulk_load(LcS[st], g.Lc + n * (int64_t)tsz, tbytes, &bars[st]);
        bulk_load(gvS[st], g.gc + n * (int64_t)gs, gbytes, &bars[st]);
      }
    } else {
      const T* G0 = Gt_in + n * (int64_t)tsz; const T* L0 = g.Lc + n * (int64_t)tsz;
      KJX_GEN_VEC(e, tsz) { GtS[st][e] = G0[e]; LcS[st][e] = L0[e]; }
      KJX_GEN_VEC(i, d) gvS[st][i] = g.gc[n * (int64_t)gs + i];
    }
  };
  unsigned phase[2] = {0u, 0u};
  if (n1 - 2 >= n0) fetch(n1 - 2, 0);
  for (int64_t n = n1 - 2; n >= n0; --n) {
    const int st = (int)((n1 - 2 - n) & 1);
    if (n - 1 >= n0 && tma) fetch(n - 1, st ^ 1);          // stage st^1 was last read two barriers ago
    if (tma) { mbar_wait(&bars[st], phase[st]); phase[st] ^= 1u; } else __syncthreads();
    const T* Gt = GtS[st]; const T* Lc = LcS[st]; const T* gv = gvS[st];
    if (g.H_steps) KJX_GEN_VEC(i, d) Hrow[i] = (T)g.H_steps[n * d + i];
    gemm_cta<T, false, false>(d, d, d, Ps, Gt, ld, [&](int i, int j, T v) { T1[i * ld + j] = v; });             // T1 = P_s[n+1] Gt
    KJX_GEN_VEC(i, d) { T v = gv[i]; for (int k = 0; k < d; ++k) v += Gt[k * ld + i] * ms[k]; mnew[i] = v; }   // :360
    __syncthreads();
    T* sc_out = (g.smooth_mean && g.return_full) ? g.smooth_cov + n * (int64_t)dd : nullptr;
    gemm_cta<T, true, false>(d, d, d, Gt, T1, ld, [&](int i, int j, T v) {                                     // :361
      const T r = Lc[i * ld + j] + v;
      Ps[i * ld + j] = r;
      if (sc_out) sc_out[i * d + j] = r;
    });
    KJX_GEN_VEC(i, d) { ms[i] = mnew[i]; if (sc_out) g.smooth_mean[n * d + i] = mnew[i]; }
    __syncthreads();
    if (!tma && n - 1 >= n0) fetch(n - 1, st ^ 1);
    project(n);
  }
}

// (log_lik, site) of every step from the smoothed marginals (sde_gp.py:373-379), embarrassingly parallel.  Without stored
// sites (Gaussian likelihood + EP before the first update) the previous site is (y, sigma^2), utils.py:241-244.
template <typename T>
__global__ void __launch_bounds__(128) generic_site_update_kernel(const SiteModel sm, int64_t N, const T* __restrict__ y,
                                                                  const T* __restrict__ pm, const T* __restrict__ pv,
                                                                  const T* sprev_mean, const T* sprev_var, T* site_mean, T* site_var) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const T pmean = sprev_mean ? sprev_mean[k] : y[k], pvar = sprev_var ? sprev_var[k] : T(sm.lik_hyp);
  const SiteOut<T> o = site_update<T>(sm, y[k], pm[k], pv[k], true, pmean, pvar);
  site_mean[k] = o.mean; site_var[k] = o.var;
}

}  // namespace kjx
