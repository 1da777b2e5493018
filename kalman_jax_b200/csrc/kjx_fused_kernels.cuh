// kjx_fused_kernels.cuh — the fused run() iteration for small state dimension (d = 2, 3) on sm_100a:
// filter (stored sites) -> RTS smoother -> per-step site update, as ONE parallel-in-time scan.
//
//   R1 fused_reduce     every thread folds its L steps into a filter element (A,b,C,eta,J); a block-wide
//                       PREFIX scan gives the element of everything before the thread's chunk and a SUFFIX
//                       scan the element of everything after it (kjx_small_math.cuh, "two-filter identity")
//   R2 fused_blockscan  one CTA: prefix + suffix scans of the block totals, shard total
//   R3 fused_filter     rebuild the true incoming (m,P), run the reference filter recursion verbatim
//                       (sde_gp.py:276-301) over the chunk, accumulate log Z, park the filtered moments in a
//                       private warp-transposed, symmetric-packed scratch (9 instead of 12 scalars per step at
//                       d=3, every access a fully coalesced 256-byte warp transaction)
//   R4 fused_smoother   smoothed state of the chunk's last step from the suffix information, then the
//                       reference RTS recursion (sde_gp.py:351-361) backwards + site update (sde_gp.py:373-379)
//
// HBM streaming: the per-step scalar arrays keep the reference layout (dt[N], y[N,1], site_mean[N,1,1],
// site_cov[N,1,1], outputs alike); a thread owns L consecutive steps, so it moves them as 256-bit vectors
// (4 steps per LDG.256/STG.256 — exactly one 32-byte sector per access, nothing fetched twice).
#pragma once
#include <type_traits>
#include "kjx_small_kernels.cuh"

// minimum resident blocks per SM requested from ptxas (register caps 168 / 128): measured best on B200
#ifndef KJX_R4_MINB
#define KJX_R4_MINB 3
#endif
#ifndef KJX_R3_MINB
#define KJX_R3_MINB 4
#endif

namespace kjx {

// ------------------------------------------------------------------------------------------------
// 4 consecutive steps of a per-step scalar array
template <typename T> struct Vec4 { T x[4]; };

__device__ __forceinline__ Vec4<double> load4(const double* p) {
  Vec4<double> v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x[0]), "=d"(v.x[1]), "=d"(v.x[2]), "=d"(v.x[3]) : "l"(p));
  return v;
}
__device__ __forceinline__ Vec4<float> load4(const float* p) {
  const float4 f = __ldg(reinterpret_cast<const float4*>(p));
  Vec4<float> v; v.x[0] = f.x; v.x[1] = f.y; v.x[2] = f.z; v.x[3] = f.w;
  return v;
}
__device__ __forceinline__ void store4(double* p, const Vec4<double>& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x[0]), "d"(v.x[1]), "d"(v.x[2]), "d"(v.x[3]) : "memory");
}
__device__ __forceinline__ void store4(float* p, const Vec4<float>& v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.x[0], v.x[1], v.x[2], v.x[3]);
}
template <typename T, bool VEC> __device__ __forceinline__ Vec4<T> load4v(const T* p) {
  if (VEC) return load4(p);
  Vec4<T> v;
#pragma unroll
  for (int j = 0; j < 4; ++j) v.x[j] = p[j];
  return v;
}
template <typename T, bool VEC> __device__ __forceinline__ void store4v(T* p, const Vec4<T>& v) {
  if (VEC) { store4(p, v); return; }
#pragma unroll
  for (int j = 0; j < 4; ++j) p[j] = v.x[j];
}

// ------------------------------------------------------------------------------------------------
// private scratch for the filtered moments: [warp][step][element][lane], element = m (D) then upper(P)
template <int D> struct Packed { static constexpr int NP = D + D * (D + 1) / 2; };

template <typename T, int D>
__device__ __forceinline__ void xf_store(T* base, const FiltState<T, D>& x) {   // base already points at [warp][step][0][lane]
  int e = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) base[32 * e++] = x.m[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) base[32 * e++] = x.P[i][j];
}
template <typename T, int D>
__device__ __forceinline__ void xf_load(const T* base, T (&m)[D], T (&P)[D][D]) {
  int e = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) m[i] = __ldcs(base + 32 * e++);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) { const T v = __ldcs(base + 32 * e++); P[i][j] = v; P[j][i] = v; }
}

template <typename T, int D> struct ScanLevel {
  T* items;      // [NS][stride] elements of this level (slots, see seg_slot)
  T* prefix;     // [NS][stride] exclusive prefix of every element within its segment
  T* suffix;     // [NS][stride] exclusive suffix of every element within its segment
  size_t stride;
  int64_t n;     // number of elements
};

// ------------------------------------------------------------------------------------------------
// Time-sharded iteration (one shard per GPU): the exchange of the shards' scan elements and of their partial nlml
// values lives INSIDE the scan / filter / smoother kernels, over NVLink / NVSwitch peer memory — no collective call and
// no extra launch.  Every rank owns a receive buffer that all ranks can write (symmetric memory):
//   doubles: summaries [2][world][stride] | flags [2][world] (u64) | nlml [2][world] | nlml flags [2][world] (u64)
// (two parities; a slot of parity p is rewritten two iterations later, and the nlml exchange — every rank waits for
// every rank's partial at the end of its smoother — keeps the ranks within one iteration of each other).
struct ShardLink {
  double* const* peer_bufs;      // DEVICE array [world]: receive buffer of every rank (own one included)
  double* my_buf;                // == peer_bufs[rank], by value (no pointer chase on the kernels' critical path)
  unsigned long long* ctrl;      // DEVICE [4], zero before first use: epoch | top-scan ticket | filter ticket | smoother ticket
  double* nlml_out;              // [1] negative log marginal likelihood of the WHOLE series (every rank gets it)
  long long stride;              // doubles per summary slot (>= packed element size)
  int world, rank;
};
__device__ __forceinline__ double* link_summary(double* buf, const ShardLink& l, int par, int r) {
  return buf + ((size_t)par * l.world + r) * l.stride;
}
__device__ __forceinline__ unsigned long long* link_flag(double* buf, const ShardLink& l, int par, int r) {
  return reinterpret_cast<unsigned long long*>(buf + (size_t)2 * l.world * l.stride) + (size_t)par * l.world + r;
}
__device__ __forceinline__ double* link_nlml(double* buf, const ShardLink& l, int par, int r) {
  return buf + (size_t)2 * l.world * l.stride + (size_t)2 * l.world + (size_t)par * l.world + r;
}
__device__ __forceinline__ unsigned long long* link_nlml_flag(double* buf, const ShardLink& l, int par, int r) {
  return reinterpret_cast<unsigned long long*>(buf + (size_t)2 * l.world * l.stride + (size_t)4 * l.world) + (size_t)par * l.world + r;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void wait_flag_sys(const unsigned long long* p, unsigned long long e) {
  unsigned long long seen;
  do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(p) : "memory"); } while (seen < e);
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// element of shard r of this epoch, out of the own receive buffer (waits for it to arrive)
template <int D>
__device__ __forceinline__ void link_wait_element(const ShardLink& l, double* mybuf, int par, unsigned long long e, int r, FSum<double, D>& s) {
  wait_flag_sys(link_flag(mybuf, l, par, r), e);
  const double* src = link_summary(mybuf, l, par, r);
  double v[FSum<double, D>::NS];
#pragma unroll
  for (int k = 0; k < FSum<double, D>::NS; ++k) v[k] = ld_relaxed_sys(src + k);
  fsum_unpack<double, D>(v, s);
}

// Elements of the shards [r0, r1) of this epoch -> stage[(r - r0) * NS ..], one LANE per element (warp 0 calls): every
// lane waits for its flag and issues its 27 loads, so the NVLink / HBM latencies of all elements overlap (walking them
// one after the other cost ~2 us per element on the critical path of an 8-GPU iteration).  At most 32 elements.
template <int D>
__device__ __forceinline__ void link_stage_elements(const ShardLink& l, int r0, int r1, double* stage) {
  constexpr int NS = FSum<double, D>::NS;
  const int lane = threadIdx.x & 31;
  if (r1 > r0) {
    const unsigned long long e = l.ctrl[0] + 1ull;
    const int par = (int)(e & 1ull);
    for (int r = r0 + lane; r < r1; r += 32) {
      wait_flag_sys(link_flag(l.my_buf, l, par, r), e);
      const double* src = link_summary(l.my_buf, l, par, r);
#pragma unroll
      for (int k = 0; k < NS; ++k) stage[(r - r0) * NS + k] = ld_relaxed_sys(src + k);
    }
  }
  __syncwarp();
}

template <typename T, int D> struct FusedArgs {
  SmallArgs<T, D> a;        // prior, site model, data pointers, scan workspace (see kjx_small_kernels.cuh)
  ShardLink link;           // time-sharded iteration only (kernels instantiated with SHARDED)
  ScanLevel<T, D> lvl[3];   // chunk elements / segment totals / super-segment totals with their scans
  T* xf;                    // filtered moments, private layout
  const T* s_info_in;       // [D + D*D] information (eta, J) of all later shards about this shard's last state
  T* info_out;              // sharded iteration: where the filter kernel's extra block leaves that information (same buffer)
  int vec_ok;               // every per-step array is 32-byte (fp64) / 16-byte (fp32) aligned
  int L2; int64_t nbig;     // chunks [0, nbig) have a.L steps, chunks [nbig, nchunks) have L2 (= a.L - 4 or a.L) steps:
                            // two lengths make the chunk count whatever fills the SMs evenly (see plan_chunks)
  int seg0, seg1;           // elements per scan segment: level 0 (64 / 128: throughput) and levels 1, 2 (32 / 64 / 128: latency)
  int store_xf;             // keep the filtered moments (0: energy-only filter)
  int jacobi;               // site-initialisation sweep (see fused_filter_kernel): out_site_* receive the sites initialised
                            // from THIS pass's predictive moments while the update uses the given sites
  T* block_diff;            // [blocks] largest relative change between the given and the newly initialised sites
};

template <typename T, int D>
__device__ __forceinline__ void load_info(const T* p, Info<T, D>& f) {
#pragma unroll
  for (int i = 0; i < D; ++i) f.eta[i] = p[i];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) f.J[i][j] = p[D + i * D + j];
}

// steps [s0, s1) of thread chunk c
template <typename T, int D>
__device__ __forceinline__ void chunk_range(const FusedArgs<T, D>& fa, int64_t c, int64_t& s0, int64_t& s1) {
  const int64_t L1 = fa.a.L;
  const bool big = c < fa.nbig;
  s0 = big ? c * L1 : fa.nbig * L1 + (c - fa.nbig) * (int64_t)fa.L2;
  s1 = s0 + (big ? L1 : (int64_t)fa.L2);
  if (s1 > fa.a.N) s1 = fa.a.N;
  if (s0 > fa.a.N) s0 = fa.a.N;
}

constexpr int KJX_SEG_MAX = 128;         // most elements per scan segment (= one block of the segment-scan kernel).  Level 0
                                         // wants long segments (fewer combines per element), the two upper levels short ones
                                         // (they are pure latency: 32 = one element per lane, no sequential part)

// Slot of element i in a per-level workspace array.  A scan warp owns KJX_SEG_PER consecutive elements per
// lane; storing element q = PER*lane + j of a segment at j*32 + lane makes the segment a [PER][32] tile that
// stages into shared memory with coalesced copies and is read bank-conflict-free.  (The fold / filter /
// smoother threads touch these arrays once per chunk, scattered, which is negligible.)
__host__ __device__ __forceinline__ int64_t seg_slot(int64_t i, int seg) {
  const int per = seg >> 5;
  const int64_t q = i % seg;
  return (i - q) + (q % per) * 32 + q / per;
}

// ------------------------------------------------------------------------------------------------
// R1a: fold.  One thread per chunk; the chunk's element goes to the workspace (struct-of-arrays, coalesced).
template <typename T, int KIND, bool VEC>
__global__ void __launch_bounds__(KJX_BLOCK) fused_fold_kernel(const FusedArgs<T, BlockDim<KIND>::D> fa) {
  constexpr int D = BlockDim<KIND>::D;
  using S = FSum<T, D>;
  const SmallArgs<T, D>& a = fa.a;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nchunks) return;
  int64_t s0, s1; chunk_range<T, D>(fa, c, s0, s1);
  T Pinf[D][D]; load_pinf<T, D>(a.prior, Pinf);
  S mine; fsum_identity<T, D>(mine);
  const T* ysrc = a.gaussian_init ? a.y : a.site_mean;
  const T lik_hyp = T(a.site.lik_hyp);
  int64_t k = s0;
  for (; k + 4 <= s1; k += 4) {
    const Vec4<T> vdt = load4v<T, VEC>(a.dt + k);
    const Vec4<T> vy = load4v<T, VEC>(ysrc + k);
    Vec4<T> vr;
    if (!a.gaussian_init) vr = load4v<T, VEC>(a.site_var + k);
    uint32_t mk = 0;
    if (a.mask) mk = (uint32_t)a.mask[k] | ((uint32_t)a.mask[k + 1] << 8) | ((uint32_t)a.mask[k + 2] << 16) | ((uint32_t)a.mask[k + 3] << 24);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T F[D][D];
      Transition<T, KIND>::eval(vdt.x[j], a.prior.lam, F);
      fsum_fold_step<T, D>(mine, F, Pinf, vy.x[j], a.gaussian_init ? lik_hyp : vr.x[j], ((mk >> (8 * j)) & 0xff) != 0);
    }
  }
  for (; k < s1; ++k) {
    T F[D][D];
    Transition<T, KIND>::eval(a.dt[k], a.prior.lam, F);
    T yt, R; step_site<T, D>(a, k, yt, R);
    fsum_fold_step<T, D>(mine, F, Pinf, yt, R, a.mask ? (a.mask[k] != 0) : false);
  }
  sum_store<S>(fa.lvl[0].items, fa.lvl[0].stride, seg_slot(c, fa.seg0), mine);
}

// ------------------------------------------------------------------------------------------------
// R2: segment scan, used at every level of the hierarchy (chunks -> 128-chunk segments -> 128-segment
// super-segments -> shard).  One block of two warps per segment: the 128 elements are staged into shared memory
// with cp.async (27 KB at d=3), warp 0 produces the exclusive PREFIX of every element, warp 1 the exclusive SUFFIX,
// concurrently.  Per warp: 3 sequential combines for the lane totals, a 5-level shuffle scan, 3 more to expand —
// a dependent chain of 11 combines whose operands all live in shared memory or registers.
template <typename S, int SEG> __device__ __forceinline__ void smem_item(const typename SumOps<S>::Scalar* sm, int slot, S& it) {
  typename SumOps<S>::Scalar v[SumOps<S>::NS];
#pragma unroll
  for (int k = 0; k < SumOps<S>::NS; ++k) v[k] = sm[k * SEG + slot];
  SumOps<S>::unpack(v, it);
}

// nlanes: lanes that hold real items (the others hold the identity): offsets >= nlanes would only combine with
// identities, so a short top level (a handful of segment totals) takes 3 steps instead of 5
template <typename S, bool REVERSE> __device__ __forceinline__ S warp_scan_inclusive_rolled(const S& mine, int lane, int nlanes = 32) {
  S inc = mine;
#pragma unroll 1
  for (int off = 1; off < nlanes; off <<= 1) {
    const int src = REVERSE ? lane + off : lane - off;
    S other = sum_shfl<S>(inc, src & 31);
    const bool ok = REVERSE ? (src < 32) : (src >= 0);
    if (!ok) SumOps<S>::identity(other);       // combining with the identity is exact: no divergence needed
    S r;
    if (REVERSE) SumOps<S>::combine(inc, other, r); else SumOps<S>::combine(other, inc, r);
    inc = r;
  }
  return inc;
}

template <typename T, int D, bool REVERSE, int SEG>
__device__ __forceinline__ void segscan_pass(const T* sm, const ScanLevel<T, D>& lv, int64_t first, int lane, T* total_out,
                                             size_t total_stride, int64_t total_slot) {
  using S = FSum<T, D>;
  constexpr int KJX_SEG_PER = SEG / 32;
  auto live = [&](int j) { return first + (int64_t)lane * KJX_SEG_PER + j < lv.n; };
  auto item = [&](int j, S& it) { if (live(j)) smem_item<S, SEG>(sm, j * 32 + lane, it); else SumOps<S>::identity(it); };
  S tot; item(0, tot);
#pragma unroll 1
  for (int j = 1; j < KJX_SEG_PER; ++j) { S it; item(j, it); S r; SumOps<S>::combine(tot, it, r); tot = r; }
  const int64_t left = lv.n - first;
  const int nlanes = (left >= (int64_t)SEG) ? 32 : (int)((left + KJX_SEG_PER - 1) / KJX_SEG_PER);
  S inc = warp_scan_inclusive_rolled<S, REVERSE>(tot, lane, nlanes);
  if (!REVERSE && total_out != nullptr) {
    S whole = sum_shfl<S>(inc, nlanes - 1);     // last lane that holds real items (the scan stops at nlanes)
    if (lane == 0) sum_store<S>(total_out, total_stride, total_slot, whole);
  }
  S run = sum_shfl<S>(inc, (REVERSE ? lane + 1 : lane - 1) & 31);
  if (REVERSE ? (lane == 31) : (lane == 0)) SumOps<S>::identity(run);
  T* out = REVERSE ? lv.suffix : lv.prefix;
  if (!REVERSE) {
#pragma unroll 1
    for (int j = 0; j < KJX_SEG_PER; ++j) {
      if (live(j)) sum_store<S>(out, lv.stride, first + j * 32 + lane, run);
      if (j + 1 < KJX_SEG_PER) { S it; item(j, it); S r; SumOps<S>::combine(run, it, r); run = r; }
    }
  } else {
#pragma unroll 1
    for (int j = KJX_SEG_PER - 1; j >= 0; --j) {
      if (live(j)) sum_store<S>(out, lv.stride, first + j * 32 + lane, run);
      if (j > 0) { S it; item(j, it); S r; SumOps<S>::combine(it, run, r); run = r; }
    }
  }
}

template <typename T, int D, int SEG> __device__ __forceinline__ void segscan_stage(T* sm, const ScanLevel<T, D>& lv, int64_t first) {
  using S = FSum<T, D>;
  constexpr int PER16 = 16 / sizeof(T);
  for (int idx = threadIdx.x; idx < S::NS * SEG / PER16; idx += 64) {
    const int k = (idx * PER16) / SEG, sl = (idx * PER16) % SEG;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + k * SEG + sl);
    const T* src = lv.items + k * lv.stride + first + sl;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

template <typename T, int D, int SEG>
__global__ void __launch_bounds__(64) fused_segscan_kernel(const ScanLevel<T, D> lv, T* total_out, size_t total_stride, int next_seg) {
  using S = FSum<T, D>;
  constexpr int KJX_SEG = SEG;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);                  // [NS][KJX_SEG]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t first = (int64_t)blockIdx.x * KJX_SEG;
  segscan_stage<T, D, SEG>(sm, lv, first);   // 16-byte asynchronous copies, coalesced (stride is a multiple of the segment)
  // the total goes where the NEXT level (segments of next_seg) will stage it from
  if (warp == 0) segscan_pass<T, D, false, SEG>(sm, lv, first, lane, total_out, total_stride, seg_slot(blockIdx.x, next_seg));
  else segscan_pass<T, D, true, SEG>(sm, lv, first, lane, nullptr, 0, 0);
}

// The two upper levels in ONE launch: every block scans one level-1 segment exactly like fused_segscan_kernel; the block
// that finishes last (ticket) then scans the single level-2 segment, whose total is the shard's element.  SHARDED: that
// block also PUBLISHES the element — thread r stores it into rank r's receive buffer over NVLink and releases the epoch
// flag —, so the exchange costs no launch of its own and starts the instant the element exists.
template <typename T, int D, int SEG, bool SHARDED>
__global__ void __launch_bounds__(64) fused_segscan_top_kernel(const ScanLevel<T, D> lv1, const ScanLevel<T, D> lv2, T* total_out,
                                                               unsigned long long* ticket, const ShardLink link) {
  using S = FSum<T, D>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);                  // [NS][SEG]
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ONE copy of the scan code serves both levels (rolled loop): the level-2 pass is executed once, by one block, and
  // would otherwise run from cold instruction memory (measured: 32 us instead of 14 for the two levels)
#pragma unroll 1
  for (int stage = 0; stage < 2; ++stage) {
    const ScanLevel<T, D>& lv = stage ? lv2 : lv1;
    const int64_t first = stage ? 0 : (int64_t)blockIdx.x * SEG;
    T* tot = stage ? total_out : lv2.items;
    const size_t tstride = stage ? 1 : lv2.stride;
    const int64_t tslot = stage ? 0 : seg_slot(blockIdx.x, SEG);
    segscan_stage<T, D, SEG>(sm, lv, first);               // cp.async.cg reads through L2: sees every block's total
    if (warp == 0) segscan_pass<T, D, false, SEG>(sm, lv, first, lane, tot, tstride, tslot);
    else segscan_pass<T, D, true, SEG>(sm, lv, first, lane, nullptr, 0, 0);
    if (stage == 0) {                                      // last block done?
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(ticket, 1ull);
        s_last = (t == (unsigned long long)gridDim.x - 1ull) ? 1 : 0;
        if (s_last) *ticket = 0ull;                        // self-resetting: the next launch starts from zero again
      }
      __syncthreads();
      if (!s_last) return;
      __threadfence();
    }
  }
  if constexpr (SHARDED && std::is_same<T, double>::value) {
    __syncthreads();                                       // total_out (written by lane 0 of warp 0) is visible to the block
    const int r = threadIdx.x;
    if (r < link.world) {
      const unsigned long long e = link.ctrl[0] + 1ull;
      const int par = (int)(e & 1ull);
      double* dst = link_summary(link.peer_bufs[r], link, par, link.rank);
#pragma unroll
      for (int k = 0; k < S::NS; ++k) dst[k] = total_out[k];
      __threadfence_system();
      st_release_sys(link_flag(link.peer_bufs[r], link, par, link.rank), e);
    }
  }
}

// largest relative change of a site, for the convergence test of the initialisation sweeps (a NaN counts as infinite)
template <typename T> __device__ __forceinline__ T site_change(T dmax, T new_mean, T new_var, T old_mean, T old_var) {
  const T a = fabs(new_mean - old_mean) / (T(1) + fabs(old_mean)), b = fabs(new_var - old_var) / (T(1) + fabs(old_var));
  T d = (a > b) ? a : b;
  if (!(d <= T(1e30))) d = T(1e30);          // NaN or overflow: "not converged"
  return (d > dmax) ? d : dmax;
}

// ------------------------------------------------------------------------------------------------
// R3
// PLAIN: the run() iteration (no mask, filtered moments parked in the private layout, sites not written back).
// !PLAIN adds, with runtime checks: masked (predict-only) steps, energy-only runs (no filtered moments kept) and
// writing the initialised sites — the stand-alone kjx_filter and the prediction path.
// Site-initialisation sweep (fa.jacobi, !PLAIN, non-Gaussian): the first iteration of a non-Gaussian model initialises
// every site from the predictive moments of its own step (sde_gp.py:287 with site_params=None), which makes the
// reference's loop a NONLINEAR recursion — one dependent chain, 0.4 M steps/s as a single-warp kernel.  It is the fixed
// point of  sites -> init(predictive(sites)):  the predictive of step k only depends on the sites of steps < k, so a sweep
// that filters with the current sites (parallel in time, like every other iteration) and re-initialises all sites from
// the predictive moments it sees is exact for one more step at least, and in practice converges in a few dozen sweeps
// because the filter forgets.  The host iterates until no site changes by more than 1e-13 (kjx_api.cu), then runs the
// ordinary pass; if it does not converge it falls back to the one-chain kernel.
// SHARDED (fp64): the filtered state entering the shard is not read from f_state_in but folded, by warp 0 of every
// block, from the elements the earlier shards published into this rank's receive buffer (waiting for them if need be);
// ONE EXTRA block folds the later shards' elements into the information the smoother kernel will need.
// The per-block log-likelihood sums are picked up by the smoother kernel's extra block.
template <typename T, int KIND, bool GAUSS_EP, bool VEC, bool PLAIN, bool SHARDED = false>
__global__ void __launch_bounds__(KJX_BLOCK, KJX_R3_MINB) fused_filter_kernel(const FusedArgs<T, BlockDim<KIND>::D> fa) {
  constexpr int D = BlockDim<KIND>::D;
  constexpr int NP = Packed<D>::NP;
  using FS = FSum<T, D>;
  const SmallArgs<T, D>& a = fa.a;
  __shared__ T red[KJX_BLOCK / 32];
  __shared__ T st_in[D + D * D];
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s0, s1; chunk_range<T, D>(fa, c, s0, s1);
  T Pinf[D][D]; load_pinf<T, D>(a.prior, Pinf);
  T nl = T(0), dmax = T(0);
  if constexpr (SHARDED && std::is_same<T, double>::value) {
    const ShardLink& l = fa.link;
    __shared__ double stage[31 * FSum<double, D>::NS];       // world <= 32
    if (threadIdx.x < 32) {
      if (blockIdx.x == gridDim.x - 1) {
        // The EXTRA block (no chunks): information of the LATER shards about this shard's last state, for the smoother
        // kernel — folded here, off the critical path, while the other blocks filter.
        link_stage_elements<D>(l, l.rank + 1, l.world, stage);
        if (threadIdx.x == 0) {
          Info<double, D> f; info_zero<double, D>(f);
#pragma unroll 1
          for (int r = l.world - 1; r > l.rank; --r) {
            FSum<double, D> sr; fsum_unpack<double, D>(stage + (r - l.rank - 1) * FSum<double, D>::NS, sr);
            fsum_combine_info<double, D>(sr, f, f);
          }
#pragma unroll
          for (int i = 0; i < D; ++i) { fa.info_out[i] = f.eta[i]; for (int j = 0; j < D; ++j) fa.info_out[D + i * D + j] = f.J[i][j]; }
        }
      } else {
        link_stage_elements<D>(l, 0, l.rank, stage);         // shard 0 starts from the prior: nothing to wait for
        if (threadIdx.x == 0) {
          FiltState<double, D> x0;
#pragma unroll
          for (int i = 0; i < D; ++i) { x0.m[i] = 0.0; for (int j = 0; j < D; ++j) x0.P[i][j] = Pinf[i][j]; }    // sde_gp.py:265
#pragma unroll 1
          for (int r = 0; r < l.rank; ++r) {
            FSum<double, D> sr; fsum_unpack<double, D>(stage + r * FSum<double, D>::NS, sr);
            fsum_apply<double, D>(sr, x0);
          }
#pragma unroll
          for (int i = 0; i < D; ++i) { st_in[i] = x0.m[i]; for (int j = 0; j < D; ++j) st_in[D + i * D + j] = x0.P[i][j]; }
        }
      }
    }
    __syncthreads();
  }
  if (s0 < s1) {
    FiltState<T, D> x; load_state<T, D>(SHARDED ? st_in : a.f_state_in, x);
    {   // everything before this chunk: super-segments, segments, chunks (outermost first)
      int64_t idx[3] = {c, c / fa.seg0, c / ((int64_t)fa.seg0 * fa.seg1)};
#pragma unroll 1
      for (int l = 2; l >= 0; --l) { FS p; sum_load<FS>(fa.lvl[l].prefix, fa.lvl[l].stride, seg_slot(idx[l], l ? fa.seg1 : fa.seg0), p); fsum_apply<T, D>(p, x); }
    }
    T* xf = fa.xf + ((c >> 5) * (int64_t)a.L * NP) * 32 + (c & 31);
    const T* ysrc = a.gaussian_init ? a.y : a.site_mean;
    const T lik_hyp = T(a.site.lik_hyp);
    int64_t k = s0;
    for (; k + 4 <= s1; k += 4) {
      const Vec4<T> vdt = load4v<T, VEC>(a.dt + k);
      const Vec4<T> vy = load4v<T, VEC>(a.y + k);
      Vec4<T> vm, vr;
      if (!a.gaussian_init) { vm = load4v<T, VEC>(a.site_mean + k); vr = load4v<T, VEC>(a.site_var + k); }
      uint32_t mk = 0;
      if (!PLAIN && a.mask) mk = (uint32_t)a.mask[k] | ((uint32_t)a.mask[k + 1] << 8) | ((uint32_t)a.mask[k + 2] << 16) | ((uint32_t)a.mask[k + 3] << 24);
      T logarg = T(1);                      // Gaussian log Z: the four log terms of the group as one log of a product
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        T F[D][D];
        Transition<T, KIND>::eval(vdt.x[j], a.prior.lam, F);
        filter_predict<T, D>(F, Pinf, x);                                   // sde_gp.py:276-278
        const T pm = x.m[0], pv = x.P[0][0];
        const T smean = a.gaussian_init ? vy.x[j] : vm.x[j], svar = a.gaussian_init ? lik_hyp : vr.x[j];
        const bool masked = !PLAIN && ((mk >> (8 * j)) & 0xff) != 0;
        if (!masked) {
          if (GAUSS_EP) {
            // utils.py:237-240: -(y-m)^2/(s)/2 - log(max(2 pi s, 1e-10))/2 with s = sigma^2 + v.  One reciprocal serves
            // both this term and the gain when the stored site variance equals sigma^2 (always, after the first EP sweep).
            const T sv = lik_hyp + pv, r = vy.x[j] - pm;
            const T rs = T(1) / sv;
            nl -= T(0.5) * (r * r) * rs;
            logarg *= max_nan(T(2 * KJX_PI) * sv, T(1e-10));               // each factor >= 1e-10 (or NaN): no sign issue
            const T S = pv + svar;
            const T Sinv = (svar == lik_hyp) ? ((S < T(0)) ? qnan<T>() : rs) : inv1(S);          // sde_gp.py:292-294
            const T innov = smean - pm;
            T HP[D];
#pragma unroll
            for (int i = 0; i < D; ++i) HP[i] = x.P[0][i];
#pragma unroll
            for (int i = 0; i < D; ++i) x.m[i] += (HP[i] * Sinv) * innov;                     // :295
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
              for (int q = i; q < D; ++q) { const T v = x.P[i][q] - (HP[i] * Sinv) * HP[q]; x.P[i][q] = v; x.P[q][i] = v; }   // :296
          } else if (!PLAIN && fa.jacobi) {
            const SiteOut<T> o = site_update<T>(a.site, vy.x[j], pm, pv, false, T(0), T(0));   // :287, from the predictive
            nl += o.lml;
            filter_update<T, D>(x, smean, svar);                            // with the CURRENT sites
            a.out_site_mean[k + j] = o.mean; a.out_site_var[k + j] = o.var;
            dmax = site_change<T>(dmax, o.mean, o.var, smean, svar);
          } else {
            nl += filter_loglik<T>(a.site, vy.x[j], pm, pv);                // sde_gp.py:287 (log Z only)
            filter_update<T, D>(x, smean, svar);                            // sde_gp.py:292-296
          }
        } else if (!PLAIN && !GAUSS_EP && fa.jacobi) {                      // masked: y := predictive mean (:286), predict-only
          const SiteOut<T> o = site_update<T>(a.site, pm, pm, pv, false, T(0), T(0));
          a.out_site_mean[k + j] = o.mean; a.out_site_var[k + j] = o.var;
          dmax = site_change<T>(dmax, o.mean, o.var, smean, svar);
        }                                                                   // masked: predict-only, :297-300
        if (PLAIN || fa.store_xf) xf_store<T, D>(xf + (k - s0 + j) * NP * 32, x);
        if (!PLAIN && a.out_site_mean && !fa.jacobi) { a.out_site_mean[k + j] = smean; a.out_site_var[k + j] = svar; }
      }
      if (GAUSS_EP) nl -= T(0.5) * log(logarg);
    }
    (void)ysrc;
    for (; k < s1; ++k) {
      T F[D][D];
      Transition<T, KIND>::eval(a.dt[k], a.prior.lam, F);
      filter_predict<T, D>(F, Pinf, x);
      T smean, svar;
      if (!PLAIN && !GAUSS_EP && fa.jacobi) {
        const bool masked = a.mask ? (a.mask[k] != 0) : false;
        const T pm = x.m[0], pv = x.P[0][0];
        step_site<T, D>(a, k, smean, svar);
        const SiteOut<T> o = site_update<T>(a.site, masked ? pm : a.y[k], pm, pv, false, T(0), T(0));
        if (!masked) { nl += o.lml; filter_update<T, D>(x, smean, svar); }
        a.out_site_mean[k] = o.mean; a.out_site_var[k] = o.var;
        dmax = site_change<T>(dmax, o.mean, o.var, smean, svar);
      } else {
        nl += filter_step_body<T, D, GAUSS_EP>(a, k, x, false, smean, svar);   // handles the mask itself
        if (!PLAIN && a.out_site_mean) { a.out_site_mean[k] = smean; a.out_site_var[k] = svar; }
      }
      if (PLAIN || fa.store_xf) xf_store<T, D>(xf + (k - s0) * NP * 32, x);
    }
  }
  if (!PLAIN && !GAUSS_EP && fa.jacobi) {                    // block maximum of the sites' change
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const T o = __shfl_down_sync(0xffffffffu, dmax, off); dmax = (o > dmax) ? o : dmax; }
    __shared__ T dred[KJX_BLOCK / 32];
    if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = dmax;
    __syncthreads();
    if (threadIdx.x == 0) {
      T t = T(0);
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = (dred[w] > t) ? dred[w] : t;
      fa.block_diff[blockIdx.x] = t;
    }
  }
  T s = nl;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    T t = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    a.block_nlml[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// R4: outputs of one smoothed step (sde_gp.py:368-379)
template <typename T, int D, bool GAUSS_EP>
__device__ __forceinline__ void smoothed_outputs(const SmallArgs<T, D>& a, const FiltState<T, D>& x, T y, T prev_mean,
                                                 T prev_var, T& pm, T& pv, T& nsm, T& nsv) {
  pm = x.m[0]; pv = x.P[0][0];
  if (GAUSS_EP) {
    nsm = y; nsv = T(a.site.lik_hyp);
    const T damping = T(a.site.damping);
    if (damping != T(1)) damp<T>(inv1(nsv), nsm, inv1(prev_var), prev_mean, damping, T(0), nsm, nsv);
  } else {
    const SiteOut<T> o = site_update<T>(a.site, y, pm, pv, true, prev_mean, prev_var);
    nsm = o.mean; nsv = o.var;
  }
}

// SHARDED (fp64): the information of the later shards about this shard's last state was left in s_info_in by the filter
// kernel's extra block.  The grid has ONE EXTRA block that exchanges the shards' partial nlml values while the others
// work; the block that finishes last closes the epoch.
template <typename T, int KIND, bool GAUSS_EP, bool VEC, bool SHARDED = false>
__global__ void __launch_bounds__(KJX_BLOCK, KJX_R4_MINB) fused_smoother_kernel(const FusedArgs<T, BlockDim<KIND>::D> fa) {
  constexpr int D = BlockDim<KIND>::D;
  constexpr int NP = Packed<D>::NP;
  using FS = FSum<T, D>;
  const SmallArgs<T, D>& a = fa.a;
  __shared__ int s_last;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s0, s1; chunk_range<T, D>(fa, c, s0, s1);
  if constexpr (SHARDED && std::is_same<T, double>::value) {
    const ShardLink& l = fa.link;
    if (blockIdx.x == gridDim.x - 1) {
      // The EXTRA block (no chunks): the nlml exchange, off everybody's critical path.  Fixed-order sum of the filter
      // kernel's per-block values -> this shard's partial, stored into every rank's buffer; then the partials of all
      // shards are summed in rank order (every rank gets the same bits).
      __shared__ double red[KJX_BLOCK / 32];
      const int64_t nb = (int64_t)gridDim.x - 1;             // blocks of the filter kernel
      double p = 0.0;
      for (int64_t b = threadIdx.x; b < nb; b += blockDim.x) p += a.block_nlml[b];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) p += __shfl_down_sync(0xffffffffu, p, off);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = p;
      __syncthreads();
      const unsigned long long e = l.ctrl[0] + 1ull;
      const int par = (int)(e & 1ull);
      if ((int)threadIdx.x < l.world) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        const int r = threadIdx.x;
        *link_nlml(l.peer_bufs[r], l, par, l.rank) = -t;      // sde_gp.py:301  nlml -= sum(log_lik_n)
        __threadfence_system();
        st_release_sys(link_nlml_flag(l.peer_bufs[r], l, par, l.rank), e);
      }
      if (threadIdx.x == 0) {
        double total = 0.0;
        for (int r = 0; r < l.world; ++r) {
          wait_flag_sys(link_nlml_flag(l.my_buf, l, par, r), e);
          total += ld_relaxed_sys(link_nlml(l.my_buf, l, par, r));
        }
        l.nlml_out[0] = total;
      }
    }
    __syncthreads();
  }
  if (s0 < s1) {
  T Pinf[D][D]; load_pinf<T, D>(a.prior, Pinf);
  const T* xf = fa.xf + ((c >> 5) * (int64_t)a.L * NP) * 32 + (c & 31);
  const T lik_hyp = T(a.site.lik_hyp);
  const bool upd = a.update_sites != 0, post = a.smooth_mean != nullptr;

  // smoothed state of the chunk's last step: filtered moments x information of everything after it
  FiltState<T, D> x;
  int64_t k = s1 - 1;
  {
    Info<T, D> info; load_info<T, D>(fa.s_info_in, info);     // sharded: left there by the filter kernel's extra block
    {   // everything after this chunk
      int64_t idx[3] = {c, c / fa.seg0, c / ((int64_t)fa.seg0 * fa.seg1)};
#pragma unroll 1
      for (int l = 2; l >= 0; --l) { FS q; sum_load<FS>(fa.lvl[l].suffix, fa.lvl[l].stride, seg_slot(idx[l], l ? fa.seg1 : fa.seg0), q); fsum_combine_info<T, D>(q, info, info); }
    }
    xf_load<T, D>(xf + (k - s0) * NP * 32, x.m, x.P);
    info_smooth<T, D>(info, x);
    T pm, pv, nsm, nsv, pmean, pvar;
    step_site<T, D>(a, k, pmean, pvar);
    smoothed_outputs<T, D, GAUSS_EP>(a, x, a.y[k], pmean, pvar, pm, pv, nsm, nsv);
    if (post) { a.smooth_mean[k] = pm; a.smooth_cov[k] = pv; }
    if (upd) { a.new_site_mean[k] = nsm; a.new_site_var[k] = nsv; }
  }
  T dt_after = a.dt[k];           // dt of the step just done = transition INTO it = what step k-1 needs (sde_gp.py:337)
  // the remaining len-1 steps: scalar I/O down to a multiple of 4, then 256-bit groups
  const int64_t ngroups = (k - s0) / 4;
  for (--k; k >= s0 + 4 * ngroups; --k) {
    T F[D][D], mf[D], Pf[D][D];
    Transition<T, KIND>::eval(dt_after, a.prior.lam, F);
    xf_load<T, D>(xf + (k - s0) * NP * 32, mf, Pf);
    SmoothGain<T, D> g;
    smoother_gain<T, D>(F, Pinf, mf, Pf, g);
    smoother_step<T, D>(g, mf, Pf, x);
    T pm, pv, nsm, nsv, pmean, pvar;
    step_site<T, D>(a, k, pmean, pvar);
    smoothed_outputs<T, D, GAUSS_EP>(a, x, a.y[k], pmean, pvar, pm, pv, nsm, nsv);
    if (post) { a.smooth_mean[k] = pm; a.smooth_cov[k] = pv; }
    if (upd) { a.new_site_mean[k] = nsm; a.new_site_var[k] = nsv; }
    dt_after = a.dt[k];
  }
  // software pipeline: the packed filtered moments of the NEXT step to process are requested one step ahead
  T nm[D], nP[D][D];
  if (ngroups > 0) xf_load<T, D>(xf + (4 * ngroups - 1) * NP * 32, nm, nP);
  for (int64_t g4 = ngroups - 1; g4 >= 0; --g4) {
    const int64_t k0 = s0 + 4 * g4;
    const Vec4<T> vdt = load4v<T, VEC>(a.dt + k0);
    const Vec4<T> vy = load4v<T, VEC>(a.y + k0);
    Vec4<T> vm, vr, opm, opv, onm, onv;
    if (!a.gaussian_init) { vm = load4v<T, VEC>(a.site_mean + k0); vr = load4v<T, VEC>(a.site_var + k0); }
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      T F[D][D], mf[D], Pf[D][D];
#pragma unroll
      for (int p = 0; p < D; ++p) {
        mf[p] = nm[p];
#pragma unroll
        for (int q = 0; q < D; ++q) Pf[p][q] = nP[p][q];
      }
      if (j > 0 || g4 > 0) xf_load<T, D>(xf + (k0 + j - 1 - s0) * NP * 32, nm, nP);      // prefetch step k-1
      Transition<T, KIND>::eval(dt_after, a.prior.lam, F);
      SmoothGain<T, D> g;
      smoother_gain<T, D>(F, Pinf, mf, Pf, g);                              // sde_gp.py:351-359
      smoother_step<T, D>(g, mf, Pf, x);                                    // sde_gp.py:360-361
      smoothed_outputs<T, D, GAUSS_EP>(a, x, vy.x[j], a.gaussian_init ? vy.x[j] : vm.x[j],
                                       a.gaussian_init ? lik_hyp : vr.x[j], opm.x[j], opv.x[j], onm.x[j], onv.x[j]);
      dt_after = vdt.x[j];
    }
    if (post) { store4v<T, VEC>(a.smooth_mean + k0, opm); store4v<T, VEC>(a.smooth_cov + k0, opv); }
    if (upd) { store4v<T, VEC>(a.new_site_mean + k0, onm); store4v<T, VEC>(a.new_site_var + k0, onv); }
  }
  }   // s0 < s1
  if constexpr (SHARDED && std::is_same<T, double>::value) {
    const ShardLink& l = fa.link;
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long t = atomicAdd(l.ctrl + 3, 1ull);
      s_last = (t == (unsigned long long)gridDim.x - 1ull) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {                        // every block (the extra one included) is through
      l.ctrl[3] = 0ull;
      __threadfence();
      l.ctrl[0] = l.ctrl[0] + 1ull;                          // epoch closed
    }
  }
}

}  // namespace kjx
