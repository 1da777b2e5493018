// kjx_sampling_kernels.cuh — draws from the state-space prior (sde_gp.py:386-417), d <= 64.
//
// The reference walks every sample through the series one step at a time and factors Q(dt) + 1e-6 I again for every
// sample and step.  Here the two kinds of work are separated:
//   sample_noise_kernel   one CTA per time step (embarrassingly parallel over N): Q = Pinf - A Pinf A^T from the sparse
//                         rows of A, Cholesky of Q + 1e-6 I in shared memory (:408-409), and the step's process noise
//                         W[k][:, s] = chol(.) eps for ALL samples s.  Slot 0 holds chol(Pinf) eps, the initial state (:404).
//   sample_path_kernel    one warp per sample: m <- A m + W[k] (4 non-zeros per row of A), f[k] = H m (:411-414).
// Random numbers: Philox-4x32-10 keyed by the caller's seed with counter (slot, sample, row pair) — counter based, so
// a draw does not depend on grid shape or launch order.  The reference's PRNGKey(i * k + k) seeding cannot be (and need
// not be) reproduced: parity for sampling is in distribution.
#pragma once
#include <stdint.h>
#include "kjx_elementwise_kernels.cuh"

namespace kjx {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// two independent standard normals for the counter (a, b, c)
__device__ __forceinline__ void normal_pair(uint64_t seed, uint32_t a, uint32_t b, uint32_t c, double& z0, double& z1) {
  uint32_t x[4];
  philox4x32_10(a, b, c, 0x6b6a78u, (uint32_t)seed, (uint32_t)(seed >> 32), x);
  // 53-bit uniforms: u1 in (0, 1], u2 in [0, 1)
  const double u1 = ((double)(((uint64_t)(x[0] >> 5) << 26) | (uint64_t)(x[1] >> 6)) + 1.0) * (1.0 / 9007199254740992.0);
  const double u2 = (double)(((uint64_t)(x[2] >> 5) << 26) | (uint64_t)(x[3] >> 6)) * (1.0 / 9007199254740992.0);
  const double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z0 = r * cs; z1 = r * sn;
}

constexpr int KJX_SAMPLE_THREADS = 128;
constexpr int KJX_SAMPLE_MAXS = 64;       // samples per launch (the step's normals live in shared memory)

inline size_t sample_noise_smem(int d, int S) {
  return (size_t)(2 * d * d + KJX_ROW_W * d + d + (size_t)d * S) * sizeof(double) + (size_t)d * sizeof(int) + 64;
}

// W: [N + 1][d][S].  slot 0: chol(Pinf) eps; slot k + 1: chol(Pinf - A_k Pinf A_k^T + 1e-6 I) eps.
__global__ void __launch_bounds__(KJX_SAMPLE_THREADS) sample_noise_kernel(const DiscArgs da, int64_t N, const double* __restrict__ dt,
                                                                          const double* __restrict__ Pinf, int S, int s0, uint64_t seed,
                                                                          double* __restrict__ W) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = da.d;
  double* Q = reinterpret_cast<double*>(smem_raw);
  double* X = Q + d * d;
  double* vals = X + d * d;
  double* rinv = vals + KJX_ROW_W * d;
  double* eps = rinv + d;
  int* c0n = reinterpret_cast<int*>(eps + (size_t)d * S);
  for (int64_t slot = blockIdx.x; slot <= N; slot += gridDim.x) {
    if (slot == 0) {
      for (int e = threadIdx.x; e < d * d; e += blockDim.x) Q[e] = Pinf[e];
    } else {
      const double h = dt[slot - 1];
      for (int r = threadIdx.x; r < d; r += blockDim.x) {
        int c0, nnz; double v[KJX_ROW_W] = {};
        transition_row<double>(da, r, h, c0, nnz, v);
        c0n[r] = c0 | (nnz << 24);
#pragma unroll
        for (int q = 0; q < KJX_ROW_W; ++q) vals[r * KJX_ROW_W + q] = v[q];
      }
      __syncthreads();
      for (int e = threadIdx.x; e < d * d; e += blockDim.x) {          // X = A Pinf
        const int i = e / d, j = e - i * d;
        const int c0 = c0n[i] & 0xffffff, nn = c0n[i] >> 24;
        double s = 0.0;
        for (int q = 0; q < nn; ++q) s += vals[i * KJX_ROW_W + q] * Pinf[(c0 + q) * d + j];
        X[e] = s;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < d * d; e += blockDim.x) {          // Q = Pinf - X A^T + 1e-6 I   (:407-408)
        const int i = e / d, j = e - i * d;
        const int c0 = c0n[j] & 0xffffff, nn = c0n[j] >> 24;
        double s = 0.0;
        for (int q = 0; q < nn; ++q) s += X[i * d + c0 + q] * vals[j * KJX_ROW_W + q];
        Q[e] = Pinf[e] - s + ((i == j) ? 1e-6 : 0.0);
      }
    }
    // the step's standard normals: eps[j][s], one Philox call per pair of rows
    for (int e = threadIdx.x; e < ((d + 1) / 2) * S; e += blockDim.x) {
      const int jp = e / S, s = e - jp * S;
      double z0, z1;
      normal_pair(seed, (uint32_t)slot, (uint32_t)(slot >> 32) ^ ((uint32_t)(s0 + s) << 8), (uint32_t)jp, z0, z1);
      eps[(2 * jp) * S + s] = z0;
      if (2 * jp + 1 < d) eps[(2 * jp + 1) * S + s] = z1;
    }
    __syncthreads();
    // in-place lower Cholesky, left-looking, one thread per row (np.linalg.cholesky: NaN on a non-positive pivot)
    for (int j = 0; j < d; ++j) {
      for (int i = j + threadIdx.x; i < d; i += blockDim.x) {
        double v = Q[i * d + j];
        for (int k = 0; k < j; ++k) v -= Q[i * d + k] * Q[j * d + k];
        Q[i * d + j] = v;
      }
      __syncthreads();
      if (threadIdx.x == 0) { const double pv = Q[j * d + j]; const double r = rsqrt(pv); rinv[j] = r; Q[j * d + j] = pv * r; }
      __syncthreads();
      const double r = rinv[j];
      for (int i = j + 1 + threadIdx.x; i < d; i += blockDim.x) Q[i * d + j] *= r;
      __syncthreads();               // column j is final before any row of the next column reads it (d > 32: several warps)
    }
    __syncthreads();
    double* Wk = W + (size_t)slot * d * S;
    for (int e = threadIdx.x; e < d * S; e += blockDim.x) {            // W = L eps
      const int i = e / S, s = e - i * S;
      double v = 0.0;
      for (int j = 0; j <= i; ++j) v += Q[i * d + j] * eps[j * S + s];
      Wk[e] = v;
    }
    __syncthreads();
  }
}

// one warp per sample; f_sample: [N][S_total] (the reference's [N, func_dim = 1, num_samps]), this launch fills columns
// s0 .. s0 + S
__global__ void __launch_bounds__(KJX_SAMPLE_THREADS) sample_path_kernel(const DiscArgs da, int64_t N, const double* __restrict__ dt,
                                                                         const double* __restrict__ H, const double* __restrict__ W,
                                                                         int S, int s0, int S_total, double* __restrict__ f_sample) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = da.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int s = blockIdx.x * nw + warp;
  if (s >= S) return;
  double* m = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * d;     // two buffers per warp
  double* mn = m + d;
  for (int i = lane; i < d; i += 32) m[i] = W[(size_t)i * S + s];              // slot 0: chol(Pinf) eps
  __syncwarp();
  for (int64_t k = 0; k < N; ++k) {
    const double h = dt[k];
    const double* Wk = W + (size_t)(k + 1) * d * S;
    double acc = 0.0;
    for (int i = lane; i < d; i += 32) {
      int c0, nnz; double v[KJX_ROW_W] = {};
      transition_row<double>(da, i, h, c0, nnz, v);
      double x = Wk[(size_t)i * S + s];
      for (int q = 0; q < nnz; ++q) x += v[q] * m[c0 + q];
      mn[i] = x;
      acc += H[i] * x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) f_sample[(size_t)k * S_total + s0 + s] = acc;
    __syncwarp();
    double* t = m; m = mn; mn = t;
  }
}

}  // namespace kjx
