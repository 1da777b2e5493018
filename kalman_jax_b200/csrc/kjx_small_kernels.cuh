// kjx_small_kernels.cuh — sm_100a kernels for small state dimension (Matern-3/2: d=2, Matern-5/2: d=3).
//
// Parallel-in-time form of SDEGP.kalman_filter (sde_gp.py:230-309) and
// SDEGP.rauch_tung_striebel_smoother (sde_gp.py:311-384):
//
//   time axis -> thread chunks of L consecutive steps -> blocks of KJX_BLOCK chunks -> shard (one GPU)
//
//   (the filter scan itself lives in kjx_fused_kernels.cuh: fold, segment scans, filter — also used by the
//    stand-alone kjx_filter)
//   smoother_reduce (S1) stand-alone version of the smoother fold (reads the stored filter output)
//   smoother_blockscan (S2) one CTA: reverse scan of block totals; deterministic nlml reduction
//   smoother_apply  (S3) every thread rebuilds the smoothed state entering its chunk from the right and
//                        runs the reference RTS recursion backwards + the per-step site update
//   filter_sequential / smoother_sequential : one dependent chain in reference order (first iteration of
//                        non-Gaussian models, where the site depends on the predicted state)
#pragma once
#include <stdint.h>
#include "kjx_small_math.cuh"

namespace kjx {

constexpr int KJX_BLOCK = 128;          // threads (= thread chunks) per block in F1/F3/S1/S3
constexpr int KJX_SCAN_BLOCK = 256;     // threads of the single-CTA block-total scans

template <typename T, int D> struct SmallPrior {
  T lam;
  T Pinf[D][D];
};

// ------------------------------------------------------------------------------------------------
// traits so one block-scan routine serves both summaries
template <typename S> struct SumOps;
template <typename T, int D> struct SumOps<FSum<T, D>> {
  using Scalar = T;
  static constexpr int NS = FSum<T, D>::NS;
  static KJX_HD void identity(FSum<T, D>& s) { fsum_identity<T, D>(s); }
  static KJX_HD void pack(const FSum<T, D>& s, T* v) { fsum_pack<T, D>(s, v); }
  static KJX_HD void unpack(const T* v, FSum<T, D>& s) { fsum_unpack<T, D>(v, s); }
  static KJX_HD void combine(const FSum<T, D>& a, const FSum<T, D>& b, FSum<T, D>& o) { fsum_combine<T, D>(a, b, o); }
};
template <typename T, int D> struct SumOps<SSum<T, D>> {
  using Scalar = T;
  static constexpr int NS = SSum<T, D>::NS;
  static KJX_HD void identity(SSum<T, D>& s) { ssum_identity<T, D>(s); }
  static KJX_HD void pack(const SSum<T, D>& s, T* v) { ssum_pack<T, D>(s, v); }
  static KJX_HD void unpack(const T* v, SSum<T, D>& s) { ssum_unpack<T, D>(v, s); }
  static KJX_HD void combine(const SSum<T, D>& a, const SSum<T, D>& b, SSum<T, D>& o) { ssum_combine<T, D>(a, b, o); }
};

// summaries live in the workspace as struct-of-arrays: scalar k of item i at base[k * stride + i]
template <typename S> __device__ __forceinline__ void sum_load(const typename SumOps<S>::Scalar* base, size_t stride, size_t i, S& s) {
  typename SumOps<S>::Scalar v[SumOps<S>::NS];
#pragma unroll
  for (int k = 0; k < SumOps<S>::NS; ++k) v[k] = base[k * stride + i];
  SumOps<S>::unpack(v, s);
}
template <typename S> __device__ __forceinline__ void sum_store(typename SumOps<S>::Scalar* base, size_t stride, size_t i, const S& s) {
  typename SumOps<S>::Scalar v[SumOps<S>::NS];
  SumOps<S>::pack(s, v);
#pragma unroll
  for (int k = 0; k < SumOps<S>::NS; ++k) base[k * stride + i] = v[k];
}

template <typename S> __device__ __forceinline__ S sum_shfl(const S& s, int src_lane) {
  using T = typename SumOps<S>::Scalar;
  T v[SumOps<S>::NS];
  SumOps<S>::pack(s, v);
#pragma unroll
  for (int k = 0; k < SumOps<S>::NS; ++k) v[k] = __shfl_sync(0xffffffffu, v[k], src_lane);
  S r; SumOps<S>::unpack(v, r);
  return r;
}

// inclusive scan across the 32 lanes of a warp; REVERSE = suffix scan (element order is always
// earlier ⊗ later)
template <typename S, bool REVERSE> __device__ __forceinline__ S warp_scan_inclusive(const S& mine, int lane) {
  S inc = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int src = REVERSE ? lane + off : lane - off;
    S other = sum_shfl<S>(inc, src & 31);
    const bool ok = REVERSE ? (src < 32) : (src >= 0);
    if (ok) {
      S r;
      if (REVERSE) SumOps<S>::combine(inc, other, r); else SumOps<S>::combine(other, inc, r);
      inc = r;
    }
  }
  return inc;
}

// Block-wide exclusive scan of one summary per thread.  smem: NS * (nwarps + 1) scalars.
// On return `excl` is the combination of all items before (REVERSE: after) this thread's and
// `total` (valid in every thread) the combination of all items of the block.
template <typename S, bool REVERSE, int NTHREADS>
__device__ void block_scan_exclusive(const S& mine, S& excl, S& total, typename SumOps<S>::Scalar* smem) {
  using T = typename SumOps<S>::Scalar;
  constexpr int NS = SumOps<S>::NS;
  constexpr int NW = NTHREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  S inc = warp_scan_inclusive<S, REVERSE>(mine, lane);
  // warp aggregate -> smem
  if (lane == (REVERSE ? 0 : 31)) {
    T v[NS]; SumOps<S>::pack(inc, v);
#pragma unroll
    for (int k = 0; k < NS; ++k) smem[k * (NW + 1) + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
    S a;
    if (lane < NW) {
      T v[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) v[k] = smem[k * (NW + 1) + lane];
      SumOps<S>::unpack(v, a);
    } else {
      SumOps<S>::identity(a);
    }
    S ainc = warp_scan_inclusive<S, REVERSE>(a, lane);
    // exclusive = inclusive of the neighbour
    S aex = sum_shfl<S>(ainc, (REVERSE ? lane + 1 : lane - 1) & 31);
    if (REVERSE ? (lane == 31) : (lane == 0)) SumOps<S>::identity(aex);
    // lanes >= NW hold identity items, so in the REVERSE case lane NW-1's neighbour is identity already
    S tot = sum_shfl<S>(ainc, REVERSE ? 0 : 31);
    __syncwarp();
    if (lane < NW) {
      T v[NS]; SumOps<S>::pack(aex, v);
#pragma unroll
      for (int k = 0; k < NS; ++k) smem[k * (NW + 1) + lane] = v[k];
    }
    if (lane == 0) {
      T v[NS]; SumOps<S>::pack(tot, v);
#pragma unroll
      for (int k = 0; k < NS; ++k) smem[k * (NW + 1) + NW] = v[k];
    }
  }
  __syncthreads();
  S wex;
  {
    T v[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) v[k] = smem[k * (NW + 1) + warp];
    SumOps<S>::unpack(v, wex);
#pragma unroll
    for (int k = 0; k < NS; ++k) v[k] = smem[k * (NW + 1) + NW];
    SumOps<S>::unpack(v, total);
  }
  S lex = sum_shfl<S>(inc, (REVERSE ? lane + 1 : lane - 1) & 31);
  if (REVERSE ? (lane == 31) : (lane == 0)) SumOps<S>::identity(lex);
  if (REVERSE) SumOps<S>::combine(lex, wex, excl); else SumOps<S>::combine(wex, lex, excl);
  __syncthreads();   // smem may be reused by the caller
}

// ------------------------------------------------------------------------------------------------
// kernel arguments
template <typename T, int D> struct SmallArgs {
  SmallPrior<T, D> prior;
  SiteModel site;
  int64_t N;               // steps in this shard
  int L;                   // steps per thread chunk
  int64_t nchunks;         // ceil(N / L)
  int64_t nblocks;         // ceil(nchunks / KJX_BLOCK)
  size_t cstride;          // stride of the per-chunk summary arrays
  size_t bstride;          // stride of the per-block summary arrays
  // inputs
  const T* dt; const T* y; const uint8_t* mask; const T* site_mean; const T* site_var;
  T dt_next;               // dt of the first step of the next shard
  int is_last_shard;       // shard holds step N-1 of the whole series
  int gaussian_init;       // sites are (y, lik_hyp): Gaussian likelihood without stored sites
  // filter outputs
  T* filt_mean; T* filt_cov; T* out_site_mean; T* out_site_var;
  // smoother inputs/outputs
  const T* filt_mean_in; const T* filt_cov_in;
  T* smooth_mean; T* smooth_cov; T* new_site_mean; T* new_site_var;
  int return_full; int update_sites;
  // workspace
  T* f_thread_prefix; T* f_block_total; T* f_block_prefix; T* f_shard_total;
  T* s_thread_suffix; T* s_block_total; T* s_block_suffix; T* s_shard_total;
  T* block_nlml; T* nlml_out;
  const T* f_state_in;     // [D + D*D] filtered state entering the shard
  const T* s_state_in;     // [D + D*D] smoothed state entering the shard from the right
};

template <typename T, int D> __device__ __forceinline__ void load_pinf(const SmallPrior<T, D>& p, T (&Pinf)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) Pinf[i][j] = p.Pinf[i][j];
}
template <typename T, int D> __device__ __forceinline__ void load_state(const T* p, FiltState<T, D>& x) {
#pragma unroll
  for (int i = 0; i < D; ++i) x.m[i] = p[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) x.P[i][j] = p[D + i * D + j];
}
template <typename T, int D> __device__ __forceinline__ void load_state_at(const T* pm, const T* pP, int64_t k, FiltState<T, D>& x) {
#pragma unroll
  for (int i = 0; i < D; ++i) x.m[i] = pm[k * D + i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) x.P[i][j] = pP[k * D * D + i * D + j];
}
template <typename T, int D> __device__ __forceinline__ void store_state(T* pm, T* pP, int64_t k, const FiltState<T, D>& x) {
#pragma unroll
  for (int i = 0; i < D; ++i) pm[k * D + i] = x.m[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) pP[k * D * D + i * D + j] = x.P[i][j];
}

// pseudo-observation of step k as the filter uses it (sde_gp.py:287-290)
template <typename T, int D>
__device__ __forceinline__ void step_site(const SmallArgs<T, D>& a, int64_t k, T& ytilde, T& R) {
  if (a.gaussian_init) { ytilde = a.y[k]; R = T(a.site.lik_hyp); }   // utils.py:241-244
  else { ytilde = a.site_mean[k]; R = a.site_var[k]; }
}

// ------------------------------------------------------------------------------------------------
// F2 / S2: single CTA scan over the per-block totals.
template <typename S, bool REVERSE>
__device__ void blockscan_body(const typename SumOps<S>::Scalar* totals, typename SumOps<S>::Scalar* prefix,
                               typename SumOps<S>::Scalar* shard_total, int64_t nblocks, size_t bstride,
                               typename SumOps<S>::Scalar* smem) {
  const int64_t per = (nblocks + KJX_SCAN_BLOCK - 1) / KJX_SCAN_BLOCK;
  const int64_t b0 = (int64_t)threadIdx.x * per;
  const int64_t b1 = (b0 + per < nblocks) ? b0 + per : nblocks;
  S acc; SumOps<S>::identity(acc);
  if (!REVERSE) {
    for (int64_t b = b0; b < b1; ++b) { S it; sum_load<S>(totals, bstride, b, it); S r; SumOps<S>::combine(acc, it, r); acc = r; }
  } else {
    for (int64_t b = b1 - 1; b >= b0; --b) { S it; sum_load<S>(totals, bstride, b, it); S r; SumOps<S>::combine(it, acc, r); acc = r; }
  }
  S excl, total;
  block_scan_exclusive<S, REVERSE, KJX_SCAN_BLOCK>(acc, excl, total, smem);
  S run = excl;
  if (!REVERSE) {
    for (int64_t b = b0; b < b1; ++b) {
      sum_store<S>(prefix, bstride, b, run);
      S it; sum_load<S>(totals, bstride, b, it); S r; SumOps<S>::combine(run, it, r); run = r;
    }
  } else {
    for (int64_t b = b1 - 1; b >= b0; --b) {
      sum_store<S>(prefix, bstride, b, run);
      S it; sum_load<S>(totals, bstride, b, it); S r; SumOps<S>::combine(it, run, r); run = r;
    }
  }
  if (threadIdx.x == 0) sum_store<S>(shard_total, 1, 0, total);
}

template <typename T, int D>
__global__ void __launch_bounds__(KJX_SCAN_BLOCK) smoother_blockscan_kernel(const SmallArgs<T, D> a) {
  using S = SSum<T, D>;
  __shared__ T smem[S::NS * (KJX_SCAN_BLOCK / 32 + 1)];
  __shared__ T red[KJX_SCAN_BLOCK / 32];
  blockscan_body<S, true>(a.s_block_total, a.s_block_suffix, a.s_shard_total, a.nblocks, a.bstride, smem);
  if (a.nlml_out != nullptr && a.block_nlml != nullptr) {
    // deterministic reduction of the per-block log-likelihood sums (fixed order)
    T s = T(0);
    for (int64_t b = threadIdx.x; b < a.nblocks; b += KJX_SCAN_BLOCK) s += a.block_nlml[b];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      T t = T(0);
      for (int w = 0; w < KJX_SCAN_BLOCK / 32; ++w) t += red[w];
      a.nlml_out[0] = -t;                            // sde_gp.py:301  nlml -= sum(log_lik_n)
    }
  }
}

// ------------------------------------------------------------------------------------------------
// one step of the reference filter given the predicted state in x; returns log_lik_n
// GAUSS_EP: compile-time fast path for Gaussian likelihood + EP (closed-form moments, utils.py:219-244)
// Red: SerialRed (one thread owns the step) or WarpRed (the 32 lanes of a warp run the same chain and share the
// cubature evaluations of every step)
template <typename T, int D, bool GAUSS_EP, class Red = SerialRed>
__device__ __forceinline__ T filter_step_body(const SmallArgs<T, D>& a, int64_t k, FiltState<T, D>& x,
                                              bool init_sites, T& smean, T& svar) {
  const bool masked = a.mask ? (a.mask[k] != 0) : false;
  const T pm = x.m[0], pv = x.P[0][0];               // sde_gp.py:283-284 with H = e0
  T lml = T(0);
  if (init_sites) {
    // sde_gp.py:287 — site from the predictive (site_params=None); masked: y := predict_mean (:286)
    const T yk = masked ? pm : a.y[k];
    SiteOut<T> o = site_update<T, Red>(a.site, yk, pm, pv, false, T(0), T(0));
    smean = o.mean; svar = o.var; lml = o.lml;
  } else {
    step_site<T, D>(a, k, smean, svar);
    // log Z is always the true likelihood under the predictive, even when stored sites do the update
    if (!masked) lml = GAUSS_EP ? gaussian_moment_match<T>(a.y[k], pm, pv, T(a.site.lik_hyp)).lml
                                : filter_loglik<T, Red>(a.site, a.y[k], pm, pv);
  }
  if (!masked) filter_update<T, D>(x, smean, svar);  // :292-296 ; masked: keep the prediction (:297-299)
  else lml = T(0);                                   // :300
  return lml;
}

// S1: smoother fold from the stored filter output (stand-alone kjx_smoother)
template <typename T, int KIND>
__global__ void __launch_bounds__(KJX_BLOCK) smoother_reduce_kernel(const SmallArgs<T, BlockDim<KIND>::D> a) {
  constexpr int D = BlockDim<KIND>::D;
  using SS = SSum<T, D>;
  __shared__ T smem[SS::NS * (KJX_BLOCK / 32 + 1)];
  const int64_t c = (int64_t)blockIdx.x * KJX_BLOCK + threadIdx.x;
  const int64_t s0 = c * a.L;
  const int64_t s1 = (s0 + a.L < a.N) ? s0 + a.L : a.N;
  T Pinf[D][D]; load_pinf<T, D>(a.prior, Pinf);
  SS acc; ssum_identity<T, D>(acc);
  for (int64_t k = s0; k < s1; ++k) {
    FiltState<T, D> xf; load_state_at(a.filt_mean_in, a.filt_cov_in, k, xf);
    SS e;
    const bool has_next = (k + 1 < a.N) || !a.is_last_shard;
    if (has_next) {
      T F[D][D];
      Transition<T, KIND>::eval((k + 1 < a.N) ? a.dt[k + 1] : a.dt_next, a.prior.lam, F);
      SmoothGain<T, D> g;
      smoother_gain<T, D>(F, Pinf, xf.m, xf.P, g);
      ssum_element<T, D>(g, xf.m, xf.P, e);
    } else {
      ssum_terminal<T, D>(xf.m, xf.P, e);
    }
    SS r; ssum_combine<T, D>(acc, e, r); acc = r;
  }
  SS excl, total;
  block_scan_exclusive<SS, true, KJX_BLOCK>(acc, excl, total, smem);
  if (c < a.nchunks) sum_store<SS>(a.s_thread_suffix, a.cstride, c, excl);
  if (threadIdx.x == 0) sum_store<SS>(a.s_block_total, a.bstride, blockIdx.x, total);
}

// one reference RTS step at k (sde_gp.py:351-379); x = smoothed state at k+1 on entry, at k on exit
template <typename T, int KIND, bool GAUSS_EP, class Red = SerialRed>
__device__ __forceinline__ void smoother_step_body(const SmallArgs<T, BlockDim<KIND>::D>& a, int64_t k,
                                                   const T (&Pinf)[BlockDim<KIND>::D][BlockDim<KIND>::D],
                                                   FiltState<T, BlockDim<KIND>::D>& x) {
  constexpr int D = BlockDim<KIND>::D;
  // sde_gp.py:337: dt shifted by one, trailing 0 at the end of the series
  const T dtn = (k + 1 < a.N) ? a.dt[k + 1] : (a.is_last_shard ? T(0) : a.dt_next);
  T F[D][D];
  Transition<T, KIND>::eval(dtn, a.prior.lam, F);
  FiltState<T, D> xf; load_state_at(a.filt_mean_in, a.filt_cov_in, k, xf);
  SmoothGain<T, D> g;
  smoother_gain<T, D>(F, Pinf, xf.m, xf.P, g);
  smoother_step<T, D>(g, xf.m, xf.P, x);
  const T pm = x.m[0], pv = x.P[0][0];               // :368-369 / :373 with H = e0
  if (a.smooth_mean) {
    if (a.return_full) store_state<T, D>(a.smooth_mean, a.smooth_cov, k, x);
    else { a.smooth_mean[k] = pm; a.smooth_cov[k] = pv; }
  }
  if (a.update_sites) {
    T pmean, pvar; step_site<T, D>(a, k, pmean, pvar);     // previous site of this step
    SiteOut<T> o;
    if (GAUSS_EP) {
      // EP with the closed-form Gaussian moment match: the new site is (y, sigma^2) whatever the cavity
      // (likelihoods.py:385-404); only damping mixes in the previous site (approximate_inference.py:68-72)
      o.mean = a.y[k]; o.var = T(a.site.lik_hyp);
      const T damping = T(a.site.damping);
      if (damping != T(1)) damp<T>(inv1(o.var), o.mean, inv1(pvar), pmean, damping, T(0), o.mean, o.var);
    } else {
      o = site_update<T, Red>(a.site, a.y[k], pm, pv, true, pmean, pvar);                          // :375
    }
    a.new_site_mean[k] = o.mean; a.new_site_var[k] = o.var;
  }
}

// S3
template <typename T, int KIND, bool GAUSS_EP>
__global__ void __launch_bounds__(KJX_BLOCK) smoother_apply_kernel(const SmallArgs<T, BlockDim<KIND>::D> a) {
  constexpr int D = BlockDim<KIND>::D;
  using SS = SSum<T, D>;
  const int64_t c = (int64_t)blockIdx.x * KJX_BLOCK + threadIdx.x;
  const int64_t s0 = c * a.L;
  const int64_t s1 = (s0 + a.L < a.N) ? s0 + a.L : a.N;
  if (s0 >= s1) return;
  T Pinf[D][D]; load_pinf<T, D>(a.prior, Pinf);
  FiltState<T, D> x;
  if (s1 == a.N && a.is_last_shard) {
    load_state_at(a.filt_mean_in, a.filt_cov_in, a.N - 1, x);      // sde_gp.py:339
  } else {
    load_state<T, D>(a.s_state_in, x);
    { SS bs; sum_load<SS>(a.s_block_suffix, a.bstride, blockIdx.x, bs); ssum_apply<T, D>(bs, x); }
    { SS ts; sum_load<SS>(a.s_thread_suffix, a.cstride, c, ts); ssum_apply<T, D>(ts, x); }
  }
  for (int64_t k = s1 - 1; k >= s0; --k) smoother_step_body<T, KIND, GAUSS_EP>(a, k, Pinf, x);
}

// ------------------------------------------------------------------------------------------------
// Sequential chains in reference order (one thread).  Used when the sites must be initialised from
// the predictive with a non-Gaussian likelihood (the recursion is then not associative), and as the
// forced "reference order" path for validation.
template <typename T, int KIND>
__global__ void filter_sequential_kernel(const SmallArgs<T, BlockDim<KIND>::D> a, int init_sites) {
  constexpr int D = BlockDim<KIND>::D;
  if (threadIdx.x >= 32 || blockIdx.x != 0) return;
  // the 32 lanes run the same chain on the same values; they split the cubature of every step, lane 0 stores
  const bool lead = threadIdx.x == 0;
  T Pinf[D][D]; load_pinf<T, D>(a.prior, Pinf);
  FiltState<T, D> x; load_state<T, D>(a.f_state_in, x);
  T nl = T(0);
  for (int64_t k = 0; k < a.N; ++k) {
    T F[D][D];
    Transition<T, KIND>::eval(a.dt[k], a.prior.lam, F);
    filter_predict<T, D>(F, Pinf, x);
    T smean, svar;
    nl += filter_step_body<T, D, false, WarpRed>(a, k, x, init_sites != 0, smean, svar);
    if (lead && a.filt_mean) store_state<T, D>(a.filt_mean, a.filt_cov, k, x);
    if (lead && a.out_site_mean) { a.out_site_mean[k] = smean; a.out_site_var[k] = svar; }
  }
  if (lead && a.nlml_out) a.nlml_out[0] = -nl;
}

template <typename T, int KIND>
__global__ void smoother_sequential_kernel(const SmallArgs<T, BlockDim<KIND>::D> a) {
  constexpr int D = BlockDim<KIND>::D;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  T Pinf[D][D]; load_pinf<T, D>(a.prior, Pinf);
  FiltState<T, D> x;
  if (a.is_last_shard) load_state_at(a.filt_mean_in, a.filt_cov_in, a.N - 1, x);
  else load_state<T, D>(a.s_state_in, x);
  for (int64_t k = a.N - 1; k >= 0; --k) smoother_step_body<T, KIND, false>(a, k, Pinf, x);
}

}  // namespace kjx
