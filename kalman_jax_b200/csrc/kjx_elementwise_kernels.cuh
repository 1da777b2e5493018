// kjx_elementwise_kernels.cuh — kernels that are embarrassingly parallel over time:
//   site_update_kernel  per-step ApproxInf.update (approximate_inference.py:36-323); also the vmapped
//                       moment match of the NLPD (sde_gp.py:139-142)
//   discretise_kernel   A(dt) = expm(F dt) by closed form per diagonal block (priors.py:163-175,
//                       238-255, 1044-1084, 1163-1181, 1398-1414) and Q = Pinf - A Pinf A^T (sde_gp.py:409)
//   nlml_reduce_kernel  fixed-order reduction of per-block log-likelihood sums
#pragma once
#include <stdint.h>
#include "kjx_sites.cuh"

namespace kjx {

template <typename T>
__global__ void __launch_bounds__(128) site_update_kernel(const SiteModel sm, int64_t N, const T* __restrict__ y,
                                                          const T* __restrict__ pm, const T* __restrict__ pv,
                                                          const T* __restrict__ sprev_mean, const T* __restrict__ sprev_var,
                                                          T* __restrict__ log_lik, T* __restrict__ site_mean,
                                                          T* __restrict__ site_var) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const bool has_prev = sprev_mean != nullptr;
  SiteOut<T> o;
  if (site_mean == nullptr && !has_prev) {
    // log Z only: EP/SLEP/VI all use the power-1 moment match, EEP its linearised energy
    o.lml = filter_loglik<T>(sm, y[k], pm[k], pv[k]);
  } else {
    o = site_update<T>(sm, y[k], pm[k], pv[k], has_prev, has_prev ? sprev_mean[k] : T(0), has_prev ? sprev_var[k] : T(0));
  }
  if (log_lik) log_lik[k] = o.lml;
  if (site_mean) { site_mean[k] = o.mean; site_var[k] = o.var; }
}

struct DiscArgs {
  int d;
  int n_blocks;
  kjx_block_t blocks[KJX_MAX_BLOCKS];
};

// rows (= columns) of one diagonal block of each kind
__host__ __device__ __forceinline__ int block_size(int kind) {
  return kind == KJX_BLOCK_MATERN12 ? 1 : (kind == KJX_BLOCK_MATERN32 || kind == KJX_BLOCK_ROTATION) ? 2 : kind == KJX_BLOCK_MATERN52 ? 3
         : (kind == KJX_BLOCK_QP_MATERN32 || kind == KJX_BLOCK_MATERN72) ? 4 : kind == KJX_BLOCK_SUBBAND_MATERN52 ? 6 : -1;
}
// the row format of A(dt): every row stores KJX_ROW_W values (its non-zeros first), the widest block being 6 x 6
constexpr int KJX_ROW_W = 6;
// Row r of A(dt): at most KJX_ROW_W non-zeros starting at column c0.  Expressions follow the reference's order.
template <typename T>
__device__ __forceinline__ void transition_row(const DiscArgs& da, int r, T dt, int& c0, int& nnz, T (&v)[KJX_ROW_W]) {
  int base = 0;
  for (int b = 0; b < da.n_blocks; ++b) {
    const kjx_block_t& blk = da.blocks[b];
    const int bs = block_size(blk.kind);
    const int span = bs * blk.count;
    if (r < base + span) {
      const int rr = (r - base) % bs;
      c0 = r - rr; nnz = bs;
      const T lam = (T)blk.lam;
      if (blk.kind == KJX_BLOCK_MATERN32) {                 // priors.py:172-174
        const T e = exp(-dt * lam);
        if (rr == 0) { v[0] = e * (dt * lam + T(1)); v[1] = e * (dt * T(1)); }
        else { v[0] = e * (dt * (-(lam * lam))); v[1] = e * (dt * (-lam) + T(1)); }
      } else if (blk.kind == KJX_BLOCK_MATERN52) {          // priors.py:246-255
        const T dtlam = dt * lam, e = exp(-dtlam), lam2 = lam * lam;
        if (rr == 0) { v[0] = e * (dt * (lam * (T(0.5) * dtlam + T(1))) + T(1)); v[1] = e * (dt * (dtlam + T(1))); v[2] = e * (dt * (T(0.5) * dt)); }
        else if (rr == 1) { v[0] = e * (dt * (T(-0.5) * dtlam * lam2)); v[1] = e * (dt * (lam * (T(1) - dtlam)) + T(1)); v[2] = e * (dt * (T(1) - T(0.5) * dtlam)); }
        else { v[0] = e * (dt * (lam2 * lam * (T(0.5) * dtlam - T(1)))); v[1] = e * (dt * (lam2 * (dtlam - T(3)))); v[2] = e * (dt * (lam * (T(0.5) * dtlam - T(2))) + T(1)); }
      } else if (blk.kind == KJX_BLOCK_MATERN12) {          // priors.py:94-104
        v[0] = exp(-dt * lam);
      } else if (blk.kind == KJX_BLOCK_ROTATION) {          // priors.py:397-408, :791-820, :910-939; utils.py:253-265
        T sn, cs;
        sincos((T)blk.omega * dt, &sn, &cs);          // one argument reduction for both
        const T e = (blk.lam != 0.0) ? exp(-dt * lam) : T(1);
        if (rr == 0) { v[0] = e * cs; v[1] = e * (-sn); } else { v[0] = e * sn; v[1] = e * cs; }
      } else if (blk.kind == KJX_BLOCK_MATERN72) {          // priors.py:326-350
        const T lam2 = lam * lam, lam3 = lam2 * lam, dtlam = dt * lam, dtlam2 = dtlam * dtlam, e = exp(-dtlam);
        if (rr == 0) {
          v[0] = e * (dt * (lam * (T(1) + T(0.5) * dtlam + dtlam2 / T(6))) + T(1)); v[1] = e * (dt * (T(1) + dtlam + T(0.5) * dtlam2));
          v[2] = e * (dt * (T(0.5) * dt * (T(1) + dtlam))); v[3] = e * (dt * (dt * dt / T(6)));
        } else if (rr == 1) {
          v[0] = e * (dt * (-dtlam2 * lam2 / T(6))); v[1] = e * (dt * (lam * (T(1) + T(0.5) * dtlam - T(0.5) * dtlam2)) + T(1));
          v[2] = e * (dt * (T(1) + dtlam - T(0.5) * dtlam2)); v[3] = e * (dt * (dt * (T(0.5) - dtlam / T(6))));
        } else if (rr == 2) {
          v[0] = e * (dt * (lam3 * dtlam * (dtlam / T(6) - T(0.5)))); v[1] = e * (dt * (dtlam * lam2 * (T(0.5) * dtlam - T(2))));
          v[2] = e * (dt * (lam * (T(1) - T(2.5) * dtlam + T(0.5) * dtlam2)) + T(1)); v[3] = e * (dt * (T(1) - dtlam + dtlam2 / T(6)));
        } else {
          v[0] = e * (dt * (lam2 * lam2 * (dtlam - T(1) - dtlam2 / T(6)))); v[1] = e * (dt * (lam3 * (T(3.5) * dtlam - T(4) - T(0.5) * dtlam2)));
          v[2] = e * (dt * (lam2 * (T(4) * dtlam - T(6) - T(0.5) * dtlam2))); v[3] = e * (dt * (lam * (T(1.5) * dtlam - T(3) - dtlam2 / T(6))) + T(1));
        }
      } else if (blk.kind == KJX_BLOCK_SUBBAND_MATERN52) {  // priors.py:704-721: (Matern-5/2 polynomial block) (x) R(dt omega)
        const T e = exp(-dt * lam), dtlam = dt * lam, lam2 = lam * lam, lam3 = lam2 * lam;
        T sn, cs;
        sincos((T)blk.omega * dt, &sn, &cs);
        const T r0 = (rr & 1) ? sn : cs, r1 = (rr & 1) ? cs : -sn;   // row (rr&1) of R
        const int br = rr >> 1;
        T f0, f1, f2;
        if (br == 0) { f0 = T(0.5) * (T(2) + dtlam * (T(2) + dtlam)); f1 = dt * (T(1) + dtlam); f2 = T(0.5) * dt * dt; }
        else if (br == 1) { f0 = T(-0.5) * dt * dt * lam3; f1 = T(1) + dtlam * (T(1) - dtlam); f2 = T(-0.5) * dt * (T(-2) + dtlam); }
        else { f0 = T(0.5) * dt * lam3 * (T(-2) + dtlam); f1 = dt * lam2 * (T(-3) + dtlam); f2 = T(0.5) * (T(2) + dtlam * (T(-4) + dtlam)); }
        v[0] = e * (f0 * r0); v[1] = e * (f0 * r1); v[2] = e * (f1 * r0); v[3] = e * (f1 * r1); v[4] = e * (f2 * r0); v[5] = e * (f2 * r1);
      } else {                                              // priors.py:1066-1084, utils.py:253-265
        const T e = exp(-dt * lam);
        T sn, cs;
        sincos((T)blk.omega * dt, &sn, &cs);          // one argument reduction for both
        // R = [[cs, -sn],[sn, cs]];  block = e * [[(1+dt lam) R, dt R], [-dt lam^2 R, (1-dt lam) R]]
        const T r0 = (rr & 1) ? sn : cs, r1 = (rr & 1) ? cs : -sn;   // row (rr&1) of R
        const T f0 = (rr < 2) ? (T(1) + dt * lam) : (-dt * lam * lam);
        const T f1 = (rr < 2) ? dt : (T(1) - dt * lam);
        v[0] = e * (f0 * r0); v[1] = e * (f0 * r1); v[2] = e * (f1 * r0); v[3] = e * (f1 * r1);
      }
      return;
    }
    base += span;
  }
  c0 = 0; nnz = 0;
}

// one warp per time step (grid-stride); dynamic smem: warps * d * (KJX_ROW_W T + 1 int)
template <typename T>
__global__ void __launch_bounds__(128) discretise_kernel(const DiscArgs da, int64_t N, const T* __restrict__ dt,
                                                         T* __restrict__ A, T* __restrict__ Q,
                                                         const double* __restrict__ Pinf) {
  extern __shared__ unsigned char smem_raw[];
  const int d = da.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  T* rows = reinterpret_cast<T*>(smem_raw) + (size_t)warp * d * KJX_ROW_W;
  int* col0 = reinterpret_cast<int*>(reinterpret_cast<T*>(smem_raw) + (size_t)nw * d * KJX_ROW_W) + (size_t)warp * d;
  for (int64_t n = (int64_t)blockIdx.x * nw + warp; n < N; n += (int64_t)gridDim.x * nw) {
    const T h = dt[n];
    for (int r = lane; r < d; r += 32) {
      int c0, nnz; T v[KJX_ROW_W] = {};
      transition_row<T>(da, r, h, c0, nnz, v);
      col0[r] = c0 | (nnz << 24);
#pragma unroll
      for (int q = 0; q < KJX_ROW_W; ++q) rows[r * KJX_ROW_W + q] = v[q];
    }
    __syncwarp();
    T* An = A + n * (int64_t)d * d;
    T* Qn = Q ? Q + n * (int64_t)d * d : nullptr;
    for (int e = lane; e < d * d; e += 32) {
      const int i = e / d, j = e % d;
      const int ci = col0[i] & 0xffffff, ni = col0[i] >> 24;
      An[e] = (j >= ci && j < ci + ni) ? rows[i * KJX_ROW_W + (j - ci)] : T(0);
      if (Qn) {
        const int cj = col0[j] & 0xffffff, nj = col0[j] >> 24;
        T s = T(0);   // (A Pinf A^T)[i][j] over the two rows' non-zeros
        for (int p = 0; p < ni; ++p) {
          T t = T(0);
          for (int q = 0; q < nj; ++q) t += (T)Pinf[(ci + p) * d + (cj + q)] * rows[j * KJX_ROW_W + q];
          s += rows[i * KJX_ROW_W + p] * t;
        }
        Qn[e] = (T)Pinf[e] - s;
      }
    }
    __syncwarp();
  }
}

template <typename T>
__global__ void nlml_reduce_kernel(const T* __restrict__ block_nlml, int64_t nblocks, T* __restrict__ nlml) {
  __shared__ T red[32];
  T s = T(0);
  for (int64_t b = threadIdx.x; b < nblocks; b += blockDim.x) s += block_nlml[b];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    T t = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    nlml[0] = -t;
  }
}

}  // namespace kjx
