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
// Parallel in time like the d <= 3 path, at CTA granularity: the series is cut into chunks of consecutive steps, one
// CTA per chunk.  generic_fold_kernel reduces every chunk to a filter scan element (A,b,C,eta,J) — O(d^2) per step
// because A(dt) is block-sparse and the update is rank one —, generic_boundary_kernel turns the elements into the
// filtered state entering every chunk and the information (eta, J) of everything after it (two CTAs: prefix, suffix),
// and then the filter / smoother kernels below run all chunks concurrently, each in reference order inside its
// chunk.  With one chunk (short series, or sites that must be initialised from the predictive) it is the plain
// sequential walk.  A multi-CTA tiled path for d > 64 is the next step, see DESIGN.md.
#pragma once
#include <stdint.h>
#include "kjx_elementwise_kernels.cuh"
#include "kjx_small_math.cuh"

namespace kjx {

constexpr int KJX_GEN_DMAX = 64;          // d x d working set (6 d^2 scalars) fits one SM's shared memory up to here
constexpr int KJX_GEN_DMAX_GLOBAL = 256;  // above it the same kernels keep their matrices in per-CTA global scratch
                                          // (L2-resident; correct but untuned: the tiled multi-CTA path is future work)
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
  // chunked scan (one CTA per chunk)
  int64_t chunk_len; int nchunks;
  T* sum;                  // [nchunks][3 d^2 + 2 d]  chunk elements: A | C | J | b | eta
  T* st_m; T* st_P;        // [nchunks][d], [nchunks][d^2]  filtered state entering every chunk
  T* in_eta; T* in_J;      // [nchunks][d], [nchunks][d^2]  information of everything after every chunk
  T* chunk_nlml;           // [nchunks]
  unsigned char* gscratch; // d > KJX_GEN_DMAX: per-CTA working set in global memory instead of shared memory
  size_t gscratch_stride;
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
  return (size_t)(6 * d * d + 4 * d + 7 * d + 16) * sizeof(T) + (size_t)(d + 4) * sizeof(int) + 16 * 16;
}

// ------------------------------------------------------------------------------------------------
// dense d x d products shared by the scan kernels (operands in shared or global memory)
template <typename T> __device__ __forceinline__ void mm_nn(int d, const T* A, const T* B, T* C) {        // C = A B
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    T s = T(0);
    for (int k = 0; k < d; ++k) s += A[i * d + k] * B[k * d + j];
    C[e] = s;
  }
}
template <typename T> __device__ __forceinline__ void mm_nt_add(int d, const T* A, const T* B, const T* S, T* C) {   // C = A B^T + S
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    T s = S[e];
    for (int k = 0; k < d; ++k) s += A[i * d + k] * B[j * d + k];
    C[e] = s;
  }
}
template <typename T> __device__ __forceinline__ void mm_tn_add(int d, const T* A, const T* B, const T* S, T* C) {   // C = A^T B + S
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    T s = S[e];
    for (int k = 0; k < d; ++k) s += A[k * d + i] * B[k * d + j];
    C[e] = s;
  }
}

// Solve M X = R in place by Gaussian elimination with partial pivoting (M: d x d, destroyed; R: d x nr, leading
// dimension ldr, overwritten by X).  Used on I + P J / I + J C, whose eigenvalues are real and >= 1 but which are not
// symmetric, so no Cholesky.  `piv` is one int of shared memory.
template <typename T> __device__ void lu_solve(int d, T* M, T* R, int nr, int ldr, int* piv) {
  for (int k = 0; k < d; ++k) {
    if (threadIdx.x < 32) {
      T best = T(-1); int bi = k;
      for (int i = k + (int)threadIdx.x; i < d; i += 32) { const T v = fabs(M[i * d + k]); if (v > best) { best = v; bi = i; } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (threadIdx.x == 0) *piv = bi;
    }
    __syncthreads();
    const int p = *piv;
    if (p != k) {
      for (int j = threadIdx.x; j < d + nr; j += blockDim.x) {
        T* a = (j < d) ? &M[k * d + j] : &R[k * ldr + (j - d)];
        T* b = (j < d) ? &M[p * d + j] : &R[p * ldr + (j - d)];
        const T t = *a; *a = *b; *b = t;
      }
    }
    __syncthreads();
    const T pinv = T(1) / M[k * d + k];
    for (int i = k + 1 + threadIdx.x; i < d; i += blockDim.x) M[i * d + k] *= pinv;
    __syncthreads();
    const int rem = d - k - 1, wid = rem + nr;
    for (int e = threadIdx.x; e < rem * wid; e += blockDim.x) {
      const int i = k + 1 + e / wid, jj = e % wid;
      const T l = M[i * d + k];
      if (jj < rem) M[i * d + k + 1 + jj] -= l * M[k * d + k + 1 + jj];
      else R[i * ldr + (jj - rem)] -= l * R[k * ldr + (jj - rem)];
    }
    __syncthreads();
  }
  for (int i = d - 1; i >= 0; --i) {
    const T dinv = T(1) / M[i * d + i];
    for (int c = threadIdx.x; c < nr; c += blockDim.x) {
      T v = R[i * ldr + c];
      for (int k = i + 1; k < d; ++k) v -= M[i * d + k] * R[k * ldr + c];
      R[i * ldr + c] = v * dinv;
    }
    __syncthreads();
  }
}

// (m, P) <- smoothed / propagated through the information (eta, J):  X = (I + P J)^-1 [P | m + P eta]
// on return R (d x (d+1), ld d+1) holds [(I+PJ)^-1 P | (I+PJ)^-1 (m + P eta)].   M, R: scratch (d^2, d^2 + d)
template <typename T>
__device__ void info_solve(int d, const T* P, const T* m, const T* J, const T* eta, T* M, T* R, int* piv) {
  const int ld = d + 1;
  for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
    const int i = e / d, j = e % d;
    T s = (i == j) ? T(1) : T(0);
    for (int k = 0; k < d; ++k) s += P[i * d + k] * J[k * d + j];
    M[e] = s;
    R[i * ld + j] = P[e];
  }
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    T s = m[i];
    for (int k = 0; k < d; ++k) s += P[i * d + k] * eta[k];
    R[i * ld + d] = s;
  }
  __syncthreads();
  lu_solve<T>(d, M, R, d + 1, ld, piv);
}

template <typename T> size_t generic_fold_smem(int d) {
  return (size_t)(5 * d * d + 4 * d + 8 * d + 16) * sizeof(T) + (size_t)d * sizeof(int) + 16 * 16;
}
template <typename T> size_t generic_boundary_smem(int d) {
  return (size_t)(4 * d * d + 4 * d + 16) * sizeof(T) + 64 + 16 * 8;
}

// fold: one CTA per chunk reduces its steps to the scan element (A, b, C, eta, J) of SURVEY.md 7.3
template <typename T>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_fold_kernel(const GenArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(g.gscratch ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const int d = g.d, dd = d * d;
  T* A = sc.take(dd); T* C = sc.take(dd); T* J = sc.take(dd); T* X = sc.take(dd); T* Pinf = sc.take(dd);
  T* vals = sc.take(4 * d); T* b = sc.take(d); T* eta = sc.take(d); T* t1 = sc.take(d); T* Hrow = sc.take(d);
  T* av = sc.take(d); T* cv = sc.take(d); T* Kv = sc.take(d); T* sh = sc.take(16);
  int* c0n = sc.take_int(d);
  const int64_t n0 = (int64_t)blockIdx.x * g.chunk_len, n1 = (n0 + g.chunk_len < g.N) ? n0 + g.chunk_len : g.N;
  for (int e = threadIdx.x; e < dd; e += blockDim.x) { Pinf[e] = (T)g.Pinf[e]; A[e] = (e / d == e % d) ? T(1) : T(0); C[e] = T(0); J[e] = T(0); }
  for (int i = threadIdx.x; i < d; i += blockDim.x) { b[i] = T(0); eta[i] = T(0); if (g.H) Hrow[i] = (T)g.H[i]; }
  __syncthreads();
  for (int64_t n = n0; n < n1; ++n) {
    build_rows<T>(g.disc, d, g.dt[n], vals, c0n);
    if (g.H_steps) for (int i = threadIdx.x; i < d; i += blockDim.x) Hrow[i] = (T)g.H_steps[n * d + i];
    __syncthreads();
    rows_left<T>(d, vals, c0n, A, X);                       // X <- F A
    rows_matvec<T>(d, vals, c0n, b, t1);                    // t1 <- F b
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) { A[e] = X[e]; X[e] = C[e] - Pinf[e]; }
    for (int i = threadIdx.x; i < d; i += blockDim.x) b[i] = t1[i];
    __syncthreads();
    rows_left<T>(d, vals, c0n, X, C);                       // C <- F (C - Pinf)
    __syncthreads();
    rows_right_add<T>(d, vals, c0n, C, Pinf, X);            // X <- F (C - Pinf) F^T + Pinf
    __syncthreads();
    const bool masked = g.mask ? (g.mask[n] != 0) : false;
    if (!masked) {
      for (int j = threadIdx.x; j < d; j += blockDim.x) {   // a = A^T h, c = C' h
        T sa = T(0), scv = T(0);
        for (int i = 0; i < d; ++i) { sa += Hrow[i] * A[i * d + j]; scv += X[j * d + i] * Hrow[i]; }
        av[j] = sa; cv[j] = scv;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        T beta = T(0), sv = T(0);
        for (int i = 0; i < d; ++i) { beta += Hrow[i] * b[i]; sv += Hrow[i] * cv[i]; }
        T yt, R;
        if (g.site_mean) { yt = g.site_mean[n]; R = g.site_var[n]; } else { yt = g.y[n]; R = T(g.site.lik_hyp); }
        sh[0] = inv1(sv + R); sh[1] = yt - beta;
      }
      __syncthreads();
      const T sinv = sh[0], res = sh[1];
      for (int i = threadIdx.x; i < d; i += blockDim.x) { Kv[i] = cv[i] * sinv; eta[i] += av[i] * (res * sinv); b[i] += cv[i] * sinv * res; }
      __syncthreads();
      for (int e = threadIdx.x; e < dd; e += blockDim.x) {
        const int i = e / d, j = e % d;
        J[e] += av[i] * av[j] * sinv;
        A[e] -= Kv[i] * av[j];
        C[e] = X[e] - Kv[i] * cv[j];
      }
    } else {
      for (int e = threadIdx.x; e < dd; e += blockDim.x) C[e] = X[e];
    }
    __syncthreads();
  }
  T* out = g.sum + (size_t)blockIdx.x * (3 * dd + 2 * d);
  for (int e = threadIdx.x; e < dd; e += blockDim.x) { out[e] = A[e]; out[dd + e] = C[e]; out[2 * dd + e] = J[e]; }
  for (int i = threadIdx.x; i < d; i += blockDim.x) { out[3 * dd + i] = b[i]; out[3 * dd + d + i] = eta[i]; }
}

// boundary: block 0 walks the chunk elements forwards (filtered state entering every chunk), block 1 backwards
// (information of everything after every chunk).  nchunks is a few hundred at most.
template <typename T>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_boundary_kernel(const GenArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(g.gscratch ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const int d = g.d, dd = d * d, ld = d + 1;
  T* P = sc.take(dd); T* M = sc.take(dd); T* R = sc.take(dd + d); T* Tm = sc.take(dd);
  T* m = sc.take(d); T* u = sc.take(d); T* w = sc.take(d);
  int* piv = sc.take_int(4);
  const size_t ssz = (size_t)3 * dd + 2 * d;
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < dd; e += blockDim.x) P[e] = (T)g.Pinf[e];          // sde_gp.py:265
    for (int i = threadIdx.x; i < d; i += blockDim.x) m[i] = T(0);
    __syncthreads();
    for (int c = 0; c < g.nchunks; ++c) {
      for (int e = threadIdx.x; e < dd; e += blockDim.x) g.st_P[(size_t)c * dd + e] = P[e];
      for (int i = threadIdx.x; i < d; i += blockDim.x) g.st_m[(size_t)c * d + i] = m[i];
      if (c + 1 == g.nchunks) break;
      const T* A = g.sum + c * ssz; const T* C = A + dd; const T* J = A + 2 * dd; const T* b = A + 3 * dd; const T* eta = b + d;
      info_solve<T>(d, P, m, J, eta, M, R, piv);            // R = [(I+PJ)^-1 P | (I+PJ)^-1 (m + P eta)]
      for (int e = threadIdx.x; e < dd; e += blockDim.x) M[e] = R[(e / d) * ld + e % d];
      for (int i = threadIdx.x; i < d; i += blockDim.x) u[i] = R[i * ld + d];
      __syncthreads();
      mm_nn<T>(d, A, M, Tm);                                // Tm = A X
      for (int i = threadIdx.x; i < d; i += blockDim.x) { T s = b[i]; for (int k = 0; k < d; ++k) s += A[i * d + k] * u[k]; m[i] = s; }
      __syncthreads();
      mm_nt_add<T>(d, Tm, A, C, P);                         // P' = A X A^T + C
      __syncthreads();
    }
  } else {
    // information (eta, J): J in P, eta in m
    for (int e = threadIdx.x; e < dd; e += blockDim.x) P[e] = T(0);
    for (int i = threadIdx.x; i < d; i += blockDim.x) m[i] = T(0);
    __syncthreads();
    for (int c = g.nchunks - 1; c >= 0; --c) {
      for (int e = threadIdx.x; e < dd; e += blockDim.x) g.in_J[(size_t)c * dd + e] = P[e];
      for (int i = threadIdx.x; i < d; i += blockDim.x) g.in_eta[(size_t)c * d + i] = m[i];
      if (c == 0) break;
      const T* A = g.sum + c * ssz; const T* C = A + dd; const T* Ji = A + 2 * dd; const T* b = A + 3 * dd; const T* etai = b + d;
      // (I + J_j C_i)^T = I + C_i J_j.  Solve (I + J_j C_i) Y = [J_j A_i | eta_j - J_j b_i]
      for (int e = threadIdx.x; e < dd; e += blockDim.x) {
        const int i = e / d, j = e % d;
        T s = (i == j) ? T(1) : T(0), r = T(0);
        for (int k = 0; k < d; ++k) { s += P[i * d + k] * C[k * d + j]; r += P[i * d + k] * A[k * d + j]; }
        M[e] = s; R[i * ld + j] = r;
      }
      for (int i = threadIdx.x; i < d; i += blockDim.x) {
        T s = m[i];
        for (int k = 0; k < d; ++k) s -= P[i * d + k] * b[k];
        R[i * ld + d] = s;
      }
      __syncthreads();
      lu_solve<T>(d, M, R, d + 1, ld, piv);
      for (int e = threadIdx.x; e < dd; e += blockDim.x) M[e] = R[(e / d) * ld + e % d];
      for (int i = threadIdx.x; i < d; i += blockDim.x) u[i] = R[i * ld + d];
      __syncthreads();
      mm_tn_add<T>(d, A, M, Ji, P);                         // J' = A_i^T Y + J_i
      for (int i = threadIdx.x; i < d; i += blockDim.x) { T s = etai[i]; for (int k = 0; k < d; ++k) s += A[k * d + i] * u[k]; w[i] = s; }
      __syncthreads();
      for (int i = threadIdx.x; i < d; i += blockDim.x) m[i] = w[i];
      // keep J' exactly symmetric
      for (int e = threadIdx.x; e < dd; e += blockDim.x) { const int i = e / d, j = e % d; if (i < j) { const T v = T(0.5) * (P[e] + P[j * d + i]); Tm[e] = v; Tm[j * d + i] = v; } else if (i == j) Tm[e] = P[e]; }
      __syncthreads();
      for (int e = threadIdx.x; e < dd; e += blockDim.x) P[e] = Tm[e];
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward filter, sde_gp.py:271-306
template <typename T>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_filter_kernel(const GenArgs<T> g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(g.gscratch ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const int d = g.d, dd = d * d;
  T* P = sc.take(dd); T* X = sc.take(dd); T* Pinf = sc.take(dd);
  T* vals = sc.take(4 * d); T* m = sc.take(d); T* mp = sc.take(d); T* Hrow = sc.take(d); T* HP = sc.take(d); T* K = sc.take(d);
  T* sh = sc.take(16);                 // scalars: 0 pm, 1 pv, 2 site_mean, 3 site_var, 4 Sinv, 5 masked, 6 nlml
  int* c0n = sc.take_int(d);
  // this CTA's chunk and the filtered state entering it (chunk 0: the prior, sde_gp.py:265)
  const int64_t n0 = (int64_t)blockIdx.x * g.chunk_len, n1 = (n0 + g.chunk_len < g.N) ? n0 + g.chunk_len : g.N;
  for (int e = threadIdx.x; e < dd; e += blockDim.x) { Pinf[e] = (T)g.Pinf[e]; P[e] = g.st_P[(size_t)blockIdx.x * dd + e]; }
  for (int i = threadIdx.x; i < d; i += blockDim.x) { m[i] = g.st_m[(size_t)blockIdx.x * d + i]; if (g.H) Hrow[i] = (T)g.H[i]; }
  if (threadIdx.x == 0) sh[6] = T(0);
  __syncthreads();
  for (int64_t n = n0; n < n1; ++n) {
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
  if (threadIdx.x == 0) g.chunk_nlml[blockIdx.x] = sh[6];
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
template <typename T> size_t generic_gain_smem(int d) {
  return (size_t)(2 * d * d + 4 * d + d + 16) * sizeof(T) + (size_t)d * sizeof(int) + 16 * 8;
}

// Working set: two d x d tiles in shared memory (56 KB at d = 59, so four CTAs share an SM — the factorisation and
// substitutions are barrier-latency-bound); Pf and Pinf are read straight from global / L1.
template <typename T>
__global__ void __launch_bounds__(KJX_GEN_THREADS) generic_gain_kernel(const GenArgs<T> g, T* __restrict__ Gt_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemCarver<T> sc(g.gscratch ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const int d = g.d, dd = d * d;
  T* W = sc.take(dd); T* Gt = sc.take(dd);
  T* vals = sc.take(4 * d); T* rinv = sc.take(d);
  int* c0n = sc.take_int(d);
  for (int64_t n = blockIdx.x; n < g.N; n += gridDim.x) {
    build_rows<T>(g.disc, d, (n + 1 < g.N) ? g.dt[n + 1] : T(0), vals, c0n);        // :337, :351
    const T* Pf = g.filt_cov_in + n * (int64_t)dd;
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {                            // Gt <- A (Pf - Pinf)
      const int i = e / d, j = e % d;
      const int c0 = c0n[i] & 0xffffff, nn = c0n[i] >> 24;
      T s = T(0);
      for (int q = 0; q < nn; ++q) s += vals[i * 4 + q] * (Pf[(c0 + q) * d + j] - (T)g.Pinf[(c0 + q) * d + j]);
      Gt[e] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {                            // W <- P_pred = . A^T + Pinf   (:354)
      const int i = e / d, j = e % d;
      const int c0 = c0n[j] & 0xffffff, nn = c0n[j] >> 24;
      T s = T(0);
      for (int q = 0; q < nn; ++q) s += Gt[i * d + c0 + q] * vals[j * 4 + q];
      W[e] = s + (T)g.Pinf[e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {                            // Gt <- A Pf                  (:353)
      const int i = e / d, j = e % d;
      const int c0 = c0n[i] & 0xffffff, nn = c0n[i] >> 24;
      T s = T(0);
      for (int q = 0; q < nn; ++q) s += vals[i * 4 + q] * Pf[(c0 + q) * d + j];
      Gt[e] = s;
    }
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
  SmemCarver<T> sc(g.gscratch ? g.gscratch + (size_t)blockIdx.x * g.gscratch_stride : smem_raw);
  const int d = g.d, dd = d * d;
  T* Pf = sc.take(dd); T* Pp = sc.take(dd + d); T* Gt = sc.take(dd); T* Ps = sc.take(dd); T* W = sc.take(dd); T* Pinf = sc.take(dd);
  T* vals = sc.take(4 * d); T* mf = sc.take(d); T* mp = sc.take(d); T* ms = sc.take(d); T* dm = sc.take(d); T* Hrow = sc.take(d);
  T* rinv = sc.take(d); T* sh = sc.take(16);
  int* c0n = sc.take_int(d);
  int* piv = sc.take_int(4);
  const int64_t n0 = (int64_t)blockIdx.x * g.chunk_len, n1 = (n0 + g.chunk_len < g.N) ? n0 + g.chunk_len : g.N;
  for (int e = threadIdx.x; e < dd; e += blockDim.x) Pinf[e] = (T)g.Pinf[e];
  for (int i = threadIdx.x; i < d; i += blockDim.x) if (g.H) Hrow[i] = (T)g.H[i];
  __syncthreads();
  for (int64_t n = n1 - 1; n >= n0; --n) {
    if (n == n1 - 1) {
      // Last step of the chunk: smoothed = filtered x information of everything after the chunk (two-filter identity;
      // for the chunk that ends the series the information is zero and this is sde_gp.py:339).  Pp/W are scratch here.
      if (g.H_steps) for (int i = threadIdx.x; i < d; i += blockDim.x) Hrow[i] = (T)g.H_steps[n * d + i];
      for (int e = threadIdx.x; e < dd; e += blockDim.x) Pf[e] = g.filt_cov_in[n * (int64_t)dd + e];
      for (int i = threadIdx.x; i < d; i += blockDim.x) mf[i] = g.filt_mean_in[n * d + i];
      __syncthreads();
      info_solve<T>(d, Pf, mf, g.in_J + (size_t)blockIdx.x * dd, g.in_eta + (size_t)blockIdx.x * d, W, Pp, piv);
      for (int e = threadIdx.x; e < dd; e += blockDim.x) {
        const int i = e / d, j = e % d;
        Ps[e] = (i <= j) ? Pp[i * (d + 1) + j] : Pp[j * (d + 1) + i];          // symmetric by construction; take the upper part
        if (g.smooth_mean && g.return_full) g.smooth_cov[n * (int64_t)dd + e] = Ps[e];
      }
      for (int i = threadIdx.x; i < d; i += blockDim.x) {
        ms[i] = Pp[i * (d + 1) + d];
        if (g.smooth_mean && g.return_full) g.smooth_mean[n * d + i] = ms[i];
      }
      __syncthreads();
    } else {
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
    }  // recursion step
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
