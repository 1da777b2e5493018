// kjx_generic_kernels.cuh — filter / smoother for a general block-diagonal prior of state dimension d <= 64
// (Sum of Matern / quasi-periodic components, spatio-temporal Matern-5/2 with up to 21 inducing points).
//
// One CTA owns one time series and walks it in reference order (sde_gp.py:271-306 and :349-379); the d x d
// matrices live in shared memory (6 d^2 scalars in the smoother: 167 KB at d = 59) and the threads of the CTA
// share the O(d^2) / O(d^3) work of each step:
//   * A(dt) is never formed densely: every row of the block-diagonal transition has at most 4 non-zeros
//     (transition_row, kjx_elementwise_kernels.cuh), so A X A^T costs 8 d^2 flops instead of 4 d^3;
//   * the update is rank one (f = 1): K = P H^T / S, P -= K (H P);
//   * the RTS gain is the reference's Cholesky solve (utils.py:25-30): right-looking factorisation in place,
//     then forward/backward substitution with one thread per right-hand side.
// This is the first correct CUDA path for d > 3 (the chunked parallel-in-time scan of the small-d path and a
// multi-CTA tiled path for d > 64 are the next step).
#pragma once
#include <stdint.h>
#include "kjx_elementwise_kernels.cuh"
#include "kjx_small_math.cuh"

namespace kjx {

constexpr int KJX_GEN_DMAX = 64;
constexpr int KJX_GEN_THREADS = 256;

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
};

// rows of A(dt) for all d rows: vals[d][4], c0n[d] = first column | nnz << 24
template <typename T>
__device__ __forceinline__ void build_rows(const DiscArgs& da, int d, T dt, T* vals, int* c0n) {
  for (int r = threadIdx.x; r < d; r += blockDim.x) {
    int c0, nnz; T v[4] = {T(0), T(0), T(0), T(0)};
    transition_row<T>(da, r, dt, c0, nnz, v);
    c0n[r] = c0 | (nnz << 24);
#pragma unroll
    for (int q = 0; q < 4; ++q) vals[r * 4 + q] = v[q];
  }
}
// out = A x
template <typename T>
__device__ __forceinline__ void rows_matvec(int d, const T* vals, const int* c0n, const T* x, T* out) {
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const int c0 = c0n[i] & 0xffffff, nn = c0n[i] >> 24;
    T s = T(0);
    for (int q = 0; q < nn; ++q) s += vals[i * 4 + q] * x[c0 + q];
    out[i] = s;
  }
}
// Y = A X   (X, Y: d x d row-major)
template <typename T>
__device__ __forceinline__ void rows_left(int d, const T* vals, const int* c0n, const T* X, T* Y) {
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    const int c0 = c0n[i] & 0xffffff, nn = c0n[i] >> 24;
    T s = T(0);
    for (int q = 0; q < nn; ++q) s += vals[i * 4 + q] * X[(c0 + q) * d + j];
    Y[e] = s;
  }
}
// Z = Y A^T + Pinf
template <typename T>
__device__ __forceinline__ void rows_right_add(int d, const T* vals, const int* c0n, const T* Y, const T* Pinf, T* Z) {
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    const int c0 = c0n[j] & 0xffffff, nn = c0n[j] >> 24;
    T s = T(0);
    for (int q = 0; q < nn; ++q) s += Y[i * d + c0 + q] * vals[j * 4 + q];
    Z[e] = s + Pinf[e];
  }
}

// shared-memory carve-up helper
template <typename T> struct SmemCarver {
  unsigned char* p;
  __device__ explicit SmemCarver(unsigned char* base) : p(base) {}
  __device__ T* take(int n) { T* r = reinterpret_cast<T*>(p); p += ((size_t)n * sizeof(T) + 15) / 16 * 16; return r; }
  __device__ int* take_int(int n) { int* r = reinterpret_cast<int*>(p); p += ((size_t)n * sizeof(int) + 15) / 16 * 16; return r; }
};

template <typename T> size_t generic_filter_smem(int d) {
  return (size_t)(3 * d * d + 4 * d + 5 * d + 16) * sizeof(T) + (size_t)d * sizeof(int) + 16 * 12;
}
template <typename T> size_t generic_smoother_smem(int d) {
  return (size_t)(6 * d * d + 4 * d + 6 * d + 16) * sizeof(T) + (size_t)d * sizeof(int) + 16 * 14;
}

// ------------------------------------------------------------------------------------------------
// forward filter, sde_gp.py:271-306
template <typename T>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_filter_kernel(const GenArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(smem_raw);
  const int d = g.d, dd = d * d;
  T* P = sc.take(dd); T* X = sc.take(dd); T* Pinf = sc.take(dd);
  T* vals = sc.take(4 * d); T* m = sc.take(d); T* mp = sc.take(d); T* Hrow = sc.take(d); T* HP = sc.take(d); T* K = sc.take(d);
  T* sh = sc.take(16);                 // scalars: 0 pm, 1 pv, 2 site_mean, 3 site_var, 4 Sinv, 5 masked, 6 nlml
  int* c0n = sc.take_int(d);
  for (int e = threadIdx.x; e < dd; e += blockDim.x) { Pinf[e] = (T)g.Pinf[e]; P[e] = (T)g.Pinf[e]; }   // sde_gp.py:265
  for (int i = threadIdx.x; i < d; i += blockDim.x) { m[i] = T(0); if (g.H) Hrow[i] = (T)g.H[i]; }
  if (threadIdx.x == 0) sh[6] = T(0);
  __syncthreads();
  for (int64_t n = 0; n < g.N; ++n) {
    build_rows<T>(g.disc, d, g.dt[n], vals, c0n);                                   // :276
    if (g.H_steps) for (int i = threadIdx.x; i < d; i += blockDim.x) Hrow[i] = (T)g.H_steps[n * d + i];   // :282
    for (int e = threadIdx.x; e < dd; e += blockDim.x) X[e] = P[e] - Pinf[e];
    __syncthreads();
    rows_matvec<T>(d, vals, c0n, m, mp);                                            // :277
    rows_left<T>(d, vals, c0n, X, P);                                               // P <- A (P - Pinf)
    __syncthreads();
    rows_right_add<T>(d, vals, c0n, P, Pinf, X);                                    // X <- P_ = . A^T + Pinf   (:278)
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += blockDim.x) {                             // HP = H P_
      T s = T(0);
      for (int i = 0; i < d; ++i) s += Hrow[i] * X[i * d + j];
      HP[j] = s;
    }
    __syncthreads();
    if (threadIdx.x < 32) {            // warp 0: the step's scalar work; cubature points spread over its lanes
      T pm = T(0), pv = T(0);
      for (int i = 0; i < d; ++i) { pm += Hrow[i] * mp[i]; pv += HP[i] * Hrow[i]; }   // :283-284
      const bool masked = g.mask ? (g.mask[n] != 0) : false;
      T smean, svar, lml = T(0);
      if (g.init_sites) {
        const SiteOut<T> o = site_update<T, WarpRed>(g.site, masked ? pm : g.y[n], pm, pv, false, T(0), T(0));   // :286-287
        smean = o.mean; svar = o.var; lml = o.lml;
      } else {
        if (g.site_mean) { smean = g.site_mean[n]; svar = g.site_var[n]; }           // :290
        else { smean = g.y[n]; svar = T(g.site.lik_hyp); }                           // Gaussian + EP initialisation
        if (!masked) lml = filter_loglik<T, WarpRed>(g.site, g.y[n], pm, pv);        // log Z is always the true likelihood's
      }
      if (masked) lml = T(0);                                                        // :300
      if (threadIdx.x == 0) {
        sh[0] = pm; sh[2] = smean; sh[3] = svar; sh[4] = inv1(pv + svar); sh[5] = masked ? T(1) : T(0);   // :292-294
        sh[6] += lml;                                                                // :301
        if (g.out_site_mean) { g.out_site_mean[n] = smean; g.out_site_var[n] = svar; }
      }
    }
    __syncthreads();
    const bool masked = sh[5] != T(0);
    const T Sinv = sh[4], innov = sh[2] - sh[0];
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      K[i] = HP[i] * Sinv;
      m[i] = masked ? mp[i] : mp[i] + K[i] * innov;                                  // :295 / :298
      if (g.filt_mean) g.filt_mean[n * d + i] = m[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {
      const int i = e / d, j = e % d;
      const T v = masked ? X[e] : X[e] - K[i] * HP[j];                               // :296 / :299
      P[e] = v;
      if (g.filt_cov) g.filt_cov[n * (int64_t)dd + e] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && g.nlml) g.nlml[0] = -sh[6];
}

// ------------------------------------------------------------------------------------------------
// in-place Cholesky (lower, left-looking) of the d x d matrix A in shared memory; rinv[j] = 1 / L[j][j].
// Column j: every row i >= j subtracts its dot product with row j of the already finished columns (one thread per
// row), then the column is scaled: 2 barriers per column and no d^2 trailing update.
template <typename T> __device__ void chol_inplace(int d, T* A, T* rinv) {
  for (int j = 0; j < d; ++j) {
    for (int i = j + threadIdx.x; i < d; i += blockDim.x) {
      T v = A[i * d + j];
      for (int k = 0; k < j; ++k) v -= A[i * d + k] * A[j * d + k];
      A[i * d + j] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const T piv = A[j * d + j];
      const T r = rsqrt_(piv);           // NaN for a negative pivot, like a failed potrf propagated by jaxlib
      rinv[j] = r;
      A[j * d + j] = piv * r;
    }
    __syncthreads();
    const T r = rinv[j];
    for (int i = j + 1 + threadIdx.x; i < d; i += blockDim.x) A[i * d + j] *= r;
    // (row j itself is only read by later columns after the next barrier)
  }
  __syncthreads();
}

// X <- solve(L L^T, X) for the d right-hand-side columns of X (utils.py:30).  The substitutions run over rows; within
// a row every column is independent, so (column, k-slice) pairs are spread over the CTA: `TPC` threads share one
// column's dot product and combine by shuffle.
template <typename T, int TPC> __device__ void chol_substitute(int d, const T* L, const T* rinv, T* X) {
  const int lane_in = threadIdx.x % TPC;
  const int col0 = threadIdx.x / TPC, ncol = blockDim.x / TPC;
  for (int i = 0; i < d; ++i) {                               // forward: L z = b
    for (int c = col0; c < d + (ncol - d % ncol) % ncol; c += ncol) {
      T v = T(0);
      if (c < d) for (int k = lane_in; k < i; k += TPC) v -= L[i * d + k] * X[k * d + c];
#pragma unroll
      for (int o = TPC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (c < d && lane_in == 0) X[i * d + c] = (X[i * d + c] + v) * rinv[i];
    }
    __syncthreads();
  }
  for (int i = d - 1; i >= 0; --i) {                          // backward: L^T x = z
    for (int c = col0; c < d + (ncol - d % ncol) % ncol; c += ncol) {
      T v = T(0);
      if (c < d) for (int k = i + 1 + lane_in; k < d; k += TPC) v -= L[k * d + i] * X[k * d + c];
#pragma unroll
      for (int o = TPC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (c < d && lane_in == 0) X[i * d + c] = (X[i * d + c] + v) * rinv[i];
    }
    __syncthreads();
  }
}

// RTS gains, embarrassingly parallel over time: Gt[n] = solve(P_pred[n], A[n+1] Pf[n]) depends only on the FILTERED
// moments (sde_gp.py:351-359), so a grid of CTAs computes them all before the (sequential) backward recursion.
template <typename T>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_gain_kernel(const GenArgs<T> g, T* __restrict__ Gt_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(smem_raw);
  const int d = g.d, dd = d * d;
  T* Pf = sc.take(dd); T* Pp = sc.take(dd); T* Gt = sc.take(dd); T* W = sc.take(dd); T* Pinf = sc.take(dd);
  T* vals = sc.take(4 * d); T* rinv = sc.take(d);
  int* c0n = sc.take_int(d);
  for (int e = threadIdx.x; e < dd; e += blockDim.x) Pinf[e] = (T)g.Pinf[e];
  __syncthreads();
  for (int64_t n = blockIdx.x; n < g.N; n += gridDim.x) {
    build_rows<T>(g.disc, d, (n + 1 < g.N) ? g.dt[n + 1] : T(0), vals, c0n);        // :337, :351
    for (int e = threadIdx.x; e < dd; e += blockDim.x) { const T v = g.filt_cov_in[n * (int64_t)dd + e]; Pf[e] = v; W[e] = v - Pinf[e]; }
    __syncthreads();
    rows_left<T>(d, vals, c0n, Pf, Gt);                                             // Gt <- A Pf        (:353)
    rows_left<T>(d, vals, c0n, W, Pp);                                              // Pp <- A (Pf - Pinf)
    __syncthreads();
    rows_right_add<T>(d, vals, c0n, Pp, Pinf, W);                                   // W  <- P_pred       (:354)
    __syncthreads();
    chol_inplace<T>(d, W, rinv);                                                    // utils.py:29
    chol_substitute<T, 4>(d, W, rinv, Gt);                                          // :359, utils.py:30
    for (int e = threadIdx.x; e < dd; e += blockDim.x) Gt_out[n * (int64_t)dd + e] = Gt[e];
    __syncthreads();
  }
}

// backward smoother, sde_gp.py:337-379
template <typename T>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_smoother_kernel(const GenArgs<T> g, const T* __restrict__ Gt_in) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(smem_raw);
  const int d = g.d, dd = d * d;
  T* Pf = sc.take(dd); T* Pp = sc.take(dd); T* Gt = sc.take(dd); T* Ps = sc.take(dd); T* W = sc.take(dd); T* Pinf = sc.take(dd);
  T* vals = sc.take(4 * d); T* mf = sc.take(d); T* mp = sc.take(d); T* ms = sc.take(d); T* dm = sc.take(d); T* Hrow = sc.take(d);
  T* rinv = sc.take(d); T* sh = sc.take(16);
  int* c0n = sc.take_int(d);
  for (int e = threadIdx.x; e < dd; e += blockDim.x) { Pinf[e] = (T)g.Pinf[e]; Ps[e] = g.filt_cov_in[(g.N - 1) * (int64_t)dd + e]; }   // :339
  for (int i = threadIdx.x; i < d; i += blockDim.x) { ms[i] = g.filt_mean_in[(g.N - 1) * d + i]; if (g.H) Hrow[i] = (T)g.H[i]; }
  __syncthreads();
  for (int64_t n = g.N - 1; n >= 0; --n) {
    build_rows<T>(g.disc, d, (n + 1 < g.N) ? g.dt[n + 1] : T(0), vals, c0n);        // :337, :351
    if (g.H_steps) for (int i = threadIdx.x; i < d; i += blockDim.x) Hrow[i] = (T)g.H_steps[n * d + i];
    for (int e = threadIdx.x; e < dd; e += blockDim.x) { const T v = g.filt_cov_in[n * (int64_t)dd + e]; Pf[e] = v; W[e] = v - Pinf[e]; }
    for (int i = threadIdx.x; i < d; i += blockDim.x) mf[i] = g.filt_mean_in[n * d + i];
    __syncthreads();
    rows_matvec<T>(d, vals, c0n, mf, mp);                                           // :352
    rows_left<T>(d, vals, c0n, W, Pp);                                              // Pp <- A (Pf - Pinf)
    __syncthreads();
    rows_right_add<T>(d, vals, c0n, Pp, Pinf, W);                                   // W  <- P_pred       (:354)
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) W[e] = Ps[e] - W[e];         // W <- Ps - P_pred
    for (int i = threadIdx.x; i < d; i += blockDim.x) dm[i] = ms[i] - mp[i];
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) Gt[e] = Gt_in[n * (int64_t)dd + e];   // gain from generic_gain_kernel
    __syncthreads();
    // m = mf + G (m - m_pred);  P = Pf + G (P - P_pred) G^T  with G = Gt^T              (:360-361)
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      T v = mf[i];
      for (int k = 0; k < d; ++k) v += Gt[k * d + i] * dm[k];
      ms[i] = v;
    }
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {          // Pp (free now) <- G (Ps - P_pred)
      const int i = e / d, j = e % d;
      T v = T(0);
      for (int k = 0; k < d; ++k) v += Gt[k * d + i] * W[k * d + j];
      Pp[e] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {
      const int i = e / d, j = e % d;
      T v = Pf[e];
      for (int k = 0; k < d; ++k) v += Pp[i * d + k] * Gt[k * d + j];
      Ps[e] = v;
      if (g.smooth_mean && g.return_full) g.smooth_cov[n * (int64_t)dd + e] = v;
    }
    if (g.smooth_mean && g.return_full) for (int i = threadIdx.x; i < d; i += blockDim.x) g.smooth_mean[n * d + i] = ms[i];
    __syncthreads();
    // marginal of the function value: H m, H P H^T                                      (:368-369, :373)
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      T s = T(0);
      for (int i = 0; i < d; ++i) s += Hrow[i] * Ps[i * d + j];
      dm[j] = s;
    }
    __syncthreads();
    if (threadIdx.x < 32) {            // warp 0, cubature points spread over its lanes
      T pm = T(0), pv = T(0);
      for (int i = 0; i < d; ++i) { pm += Hrow[i] * ms[i]; pv += dm[i] * Hrow[i]; }
      if (threadIdx.x == 0 && g.smooth_mean && !g.return_full) { g.smooth_mean[n] = pm; g.smooth_cov[n] = pv; }
      if (g.update_sites) {
        T pmean, pvar;
        if (g.site_mean) { pmean = g.site_mean[n]; pvar = g.site_var[n]; } else { pmean = g.y[n]; pvar = T(g.site.lik_hyp); }
        __syncwarp();                    // every lane has read the previous site before lane 0 may overwrite it in place
        const SiteOut<T> o = site_update<T, WarpRed>(g.site, g.y[n], pm, pv, true, pmean, pvar);   // :375
        if (threadIdx.x == 0) { g.new_site_mean[n] = o.mean; g.new_site_var[n] = o.var; }
      }
    }
    __syncthreads();
  }
}

}  // namespace kjx
