// kjx_api.cu — the C ABI of libkjx.so (see include/kjx.h) and the dispatch onto the sm_100a kernels.
// KJX_PART: 64 = everything but the _f32 entry points, 32 = only the _f32 entry points (the two halves compile side by
// side and instantiate disjoint sets of kernels); undefined = the whole interface in one object.
#ifndef KJX_PART
#define KJX_PART 0
#endif
#include <vector>

#include "kjx_internal.cuh"
#include "kjx_small_kernels.cuh"
#include "kjx_fused_kernels.cuh"
#include "kjx_multi_kernels.cuh"

using namespace kjx;
using namespace kjx_impl;

namespace {

// the register-resident path: one Matern-3/2 or Matern-5/2 block, scalar site, H = e0
int small_kind(const kjx_model_t* m) {
  if (m->func_dim != 1 || m->n_blocks != 1 || m->blocks[0].count != 1 || m->H_steps) return 0;
  const int kind = m->blocks[0].kind;
  if (kind != KJX_BLOCK_MATERN32 && kind != KJX_BLOCK_MATERN52) return 0;
  if (!m->H) return 0;
  const int d = m->state_dim;
  if (m->H[0] != 1.0) return 0;
  for (int i = 1; i < d; ++i) if (m->H[i] != 0.0) return 0;
  return kind;
}

// number of SMs of the current device (queried once per device; 148 on a B200)
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cached[dev];
}
// tuning knobs, read from the environment ONCE per process (0 = not set): KJX_CHUNK steps per thread chunk,
// KJX_THREADS threads per block of the fold / filter / smoother kernels
int env_knob(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}
int knob_chunk() { static const int v = env_knob("KJX_CHUNK"); return v; }
int knob_threads() { static const int v = env_knob("KJX_THREADS"); return v; }
int knob_seg() { static const int v = env_knob("KJX_SEG"); return v; }
int knob_balance() { static const int v = env_knob("KJX_BALANCE"); return v; }
int knob_seg1() { static const int v = env_knob("KJX_SEG1"); return v; }

// steps per thread chunk: long chunks amortise the O(d^3) scan combines, short ones keep every SM
// busy on short series.  A multiple of 4 so a chunk starts on a 32-byte boundary of every array.
int pick_chunk(int64_t N) {
  // the fused scan hierarchy has three levels of at most 128: at most 128^3 chunks
  const int64_t max_chunks = (int64_t)KJX_SEG_MAX * KJX_SEG_MAX * KJX_SEG_MAX;
  // Long series: 64 steps per thread amortise the O(d^3) scan combines (several waves of blocks anyway).
  // Short series / small shards: the kernels are one wave of latency-bound threads, so the time is (steps per
  // thread) x (latency of a step): spread the series over 95 % of the resident thread slots (3 blocks of 128 per SM).
  const int64_t slots = (int64_t)(sm_count() * 3 * KJX_BLOCK * 0.95);
  int64_t L = (N + slots - 1) / slots;
  L = std::max<int64_t>(4, std::min<int64_t>(64, L));
  const int v = knob_chunk();
  if (v >= 4 && v % 4 == 0) L = v;
  if ((N + L - 1) / L > max_chunks) L = (N + max_chunks - 1) / max_chunks;      // also bounds the knob
  return (int)round_up((size_t)L, 4);
}
// threads per block of the fold / filter / smoother kernels: 128 when there are several waves of blocks; for a single
// wave smaller blocks spread the chunks more evenly over the SMs (the kernel ends with its fullest SM)
int pick_threads(int64_t nchunks) {
  const int v = knob_threads();
  if (v == 32 || v == 64 || v == 128) return v;
  const int64_t one_wave = (int64_t)sm_count() * 3 * KJX_BLOCK;
  (void)one_wave;
  return KJX_BLOCK;
}

// Chunking of the FUSED kernels (fold / filter / smoother: one thread per chunk, 128-thread blocks).  The smoother runs 3
// blocks per SM (168 registers), the fold and the filter 4; a kernel ends with its fullest SM, so the chunk count is made
// a whole number of waves: a multiple of SMs x 3 blocks (and of SMs x 12 when that leaves 32..96 steps per chunk, which
// is whole waves for all three kernels).  An exact count needs two chunk lengths, L and L - 4 (both multiples of 4: every
// chunk starts on a 32-byte boundary of the per-step arrays).  Measured on a 2^21-step shard: 12 warps on every SM instead
// of 12 on some and 8 on others.  Short series fall back to equal chunks of pick_chunk().
struct ChunkPlan { int L1, L2; int64_t nbig, nchunks; };
ChunkPlan plan_chunks(int64_t N) {
  ChunkPlan p;
  const int64_t max_chunks = (int64_t)KJX_SEG_MAX * KJX_SEG_MAX * KJX_SEG_MAX;
  const int64_t wave3 = (int64_t)sm_count() * 3, wave12 = (int64_t)sm_count() * 12;      // blocks
  int64_t blocks = 0;
  // (off by default: on B200 it measured no faster than equal chunks — 2^21 steps: 192 vs 192.5 us per iteration — and
  // slower at 2^24, where it picks 76 / 72-step chunks: 1.185 vs 1.138 ms; KJX_BALANCE=1 turns it on for experiments)
  if (knob_balance() == 1 && knob_chunk() == 0 && N >= wave3 * KJX_BLOCK * 8) {
    // whole waves with about 64 steps per chunk; prefer the common multiple when the chunks stay reasonable
    const double want = (double)N / (64.0 * KJX_BLOCK);
    const int64_t k12 = std::max<int64_t>(1, (int64_t)(want / wave12 + 0.5));
    const double l12 = (double)N / (double)(k12 * wave12 * KJX_BLOCK);
    if (l12 >= 32.0 && l12 <= 96.0) blocks = k12 * wave12;
    else blocks = std::max<int64_t>(1, (int64_t)(want / wave3 + 0.5)) * wave3;
    if (blocks * KJX_BLOCK > max_chunks) blocks = 0;
  }
  if (blocks == 0) {                                          // equal chunks
    p.L1 = p.L2 = pick_chunk(N);
    p.nchunks = std::max<int64_t>(1, (N + p.L1 - 1) / p.L1);
    p.nbig = p.nchunks;
    return p;
  }
  p.nchunks = blocks * KJX_BLOCK;
  const int64_t q = (N + 3) / 4;                              // groups of 4 steps to hand out
  const int64_t base = q / p.nchunks, extra = q % p.nchunks; // `extra` chunks get one group more
  p.L1 = (int)(4 * (base + (extra ? 1 : 0)));
  p.L2 = (int)(4 * base);
  p.nbig = extra ? extra : p.nchunks;
  if (p.L2 < 4) { p.L1 = p.L2 = pick_chunk(N); p.nchunks = std::max<int64_t>(1, (N + p.L1 - 1) / p.L1); p.nbig = p.nchunks; }
  return p;
}

template <typename T, int D> struct SmallLayout {
  int L; int64_t nchunks, nblocks; size_t cstride, bstride;   // equal chunks: the stand-alone smoother's kernels
  ChunkPlan plan;                   // fused kernels
  int threads; int64_t fblocks;     // fused kernels: threads per block (32 / 64 / 128) and blocks
  size_t off_ftp, off_fbt, off_fbp, off_fst, off_sts, off_sbt, off_sbs, off_sst, off_nl, off_fin, off_sin;
  size_t off_info, off_xf, off_jac, off_jdiff, off_lvl[3][3], lvl_stride[3], total;
  int64_t lvl_n[3];
  int seg0, seg1;
  // jac: reserve the buffers of the site-initialisation sweeps (only models that need them: anything but Gaussian + EP)
  explicit SmallLayout(int64_t N, bool jac = true) {
    constexpr size_t FNS = FSum<T, D>::NS, SNS = SSum<T, D>::NS;
    L = pick_chunk(N);
    nchunks = std::max<int64_t>(1, (N + L - 1) / L);
    nblocks = (nchunks + KJX_BLOCK - 1) / KJX_BLOCK;
    plan = plan_chunks(N);
    threads = pick_threads(plan.nchunks);
    fblocks = (plan.nchunks + threads - 1) / threads;
    const int64_t fchunks = plan.nchunks;
    cstride = round_up((size_t)nchunks, 128);
    bstride = round_up((size_t)nblocks, 32);
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += round_up(n * sizeof(T), 256); return r; };
    off_ftp = take(FNS * cstride); off_fbt = take(FNS * bstride); off_fbp = take(FNS * bstride); off_fst = take(FNS);
    off_sts = take(SNS * cstride); off_sbt = take(SNS * bstride); off_sbs = take(SNS * bstride); off_sst = take(SNS);
    off_nl = take(round_up((size_t)((std::max(nchunks, fchunks) + 31) / 32) + 2, 32)); off_fin = take(D + D * D); off_sin = take(D + D * D);
    off_info = take(D + D * D);
    // three scan levels of the fused path: elements, exclusive prefix, exclusive suffix (strides = whole segments)
    // segments of 64 while three levels of 64 hold every chunk (shorter dependent chains), else 128
    // level 0: long segments (3 + 5 + 3 combines per 128 elements instead of 1 + 5 + 1 per 64: less FP64 work per
    // element; measured 16.5 vs 20.5 us at 52 429 chunks) unless the series is short; levels 1 and 2 are one or a few
    // blocks each, i.e. pure latency: the shortest segments that still reach the top in three levels
    seg0 = (fchunks >= 8192) ? 128 : 64;
    if (knob_seg() == 64 || knob_seg() == 128) seg0 = knob_seg();
    seg1 = 128;
    for (int c : {64, 32}) if ((int64_t)seg0 * c * c >= fchunks) seg1 = c;
    if (knob_seg1() == 32 || knob_seg1() == 64 || knob_seg1() == 128) seg1 = std::max(seg1, knob_seg1());
    if ((int64_t)seg0 * seg1 * seg1 < fchunks) { seg0 = 128; seg1 = 128; }
    int64_t n = fchunks;
    for (int l = 0; l < 3; ++l) {
      lvl_n[l] = n; lvl_stride[l] = round_up((size_t)n, KJX_SEG_MAX);
      for (int q = 0; q < 3; ++q) off_lvl[l][q] = take(FNS * lvl_stride[l]);
      n = (n + (l ? seg1 : seg0) - 1) / (l ? seg1 : seg0);
    }
    // filtered moments between the fused filter and smoother passes: [warp][step][element][lane]
    off_xf = take(round_up((size_t)fchunks, 128) * (size_t)plan.L1 * Packed<D>::NP);
    // site-initialisation sweeps (first iteration of a non-Gaussian model): two pairs of site arrays + per-block changes
    off_jac = take(jac ? 4 * (size_t)std::max<int64_t>(N, 1) : 0); off_jdiff = take(round_up((size_t)fblocks + 2, 32) + 32);
    total = o;
  }
};

template <typename T, int D>
void bind_workspace(SmallArgs<T, D>& a, const SmallLayout<T, D>& lay, void* ws) {
  char* w = (char*)ws;
  a.L = lay.L; a.nchunks = lay.nchunks; a.nblocks = lay.nblocks; a.cstride = lay.cstride; a.bstride = lay.bstride;
  a.f_thread_prefix = (T*)(w + lay.off_ftp); a.f_block_total = (T*)(w + lay.off_fbt);
  a.f_block_prefix = (T*)(w + lay.off_fbp); a.f_shard_total = (T*)(w + lay.off_fst);
  a.s_thread_suffix = (T*)(w + lay.off_sts); a.s_block_total = (T*)(w + lay.off_sbt);
  a.s_block_suffix = (T*)(w + lay.off_sbs); a.s_shard_total = (T*)(w + lay.off_sst);
  a.block_nlml = (T*)(w + lay.off_nl);
  a.f_state_in = (T*)(w + lay.off_fin); a.s_state_in = (T*)(w + lay.off_sin);
}

template <typename T, int D>
void bind_fused(FusedArgs<T, D>& fa, const SmallLayout<T, D>& lay, void* ws) {
  char* w = (char*)ws;
  bind_workspace<T, D>(fa.a, lay, ws);
  fa.s_info_in = (T*)(w + lay.off_info); fa.info_out = (T*)(w + lay.off_info); fa.xf = (T*)(w + lay.off_xf);
  fa.seg0 = lay.seg0; fa.seg1 = lay.seg1;
  fa.a.L = lay.plan.L1; fa.L2 = lay.plan.L2; fa.nbig = lay.plan.nbig; fa.a.nchunks = lay.plan.nchunks;
  for (int l = 0; l < 3; ++l) {
    fa.lvl[l].items = (T*)(w + lay.off_lvl[l][0]); fa.lvl[l].prefix = (T*)(w + lay.off_lvl[l][1]);
    fa.lvl[l].suffix = (T*)(w + lay.off_lvl[l][2]); fa.lvl[l].stride = lay.lvl_stride[l]; fa.lvl[l].n = lay.lvl_n[l];
  }
}

inline bool aligned_to(const void* p, size_t a) { return p == nullptr || ((uintptr_t)p % a) == 0; }

template <typename T, int D>
void fill_prior(SmallArgs<T, D>& a, const kjx_model_t* m) {
  a.prior.lam = (T)m->blocks[0].lam;
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) a.prior.Pinf[i][j] = (T)m->Pinf[i * D + j];
  a.site = make_site(m);
}

// writes a (mean, cov) boundary state into device memory from kernel arguments
template <typename T, int D> struct StateArg { T v[D + D * D]; };
template <typename T, int D> __global__ void set_state_kernel(T* dst, StateArg<T, D> s) {
  if (threadIdx.x < D + D * D) dst[threadIdx.x] = s.v[threadIdx.x];
}
template <typename T, int D> __global__ void copy_state_kernel(T* dst, const T* src) {
  if (threadIdx.x < D + D * D) dst[threadIdx.x] = src[threadIdx.x];
}
template <typename T> __global__ void copy_scalars_kernel(T* dst, const T* src, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

template <typename T, int D>
void launch_set_prior_state(T* dst, const kjx_model_t* m, cudaStream_t st) {
  StateArg<T, D> s;
  for (int i = 0; i < D; ++i) s.v[i] = T(0);                    // sde_gp.py:265  m0 = 0, P0 = Pinf
  for (int i = 0; i < D * D; ++i) s.v[D + i] = (T)m->Pinf[i];
  set_state_kernel<T, D><<<1, 32, 0, st>>>(dst, s);
}


struct SmallCall {
  // what the caller asked for (device pointers, type-erased)
  const void *dt, *y; const uint8_t* mask; const void *site_mean, *site_var;
  void *filt_mean, *filt_cov, *out_site_mean, *out_site_var;
  const void *filt_mean_in, *filt_cov_in;
  void *smooth_mean, *smooth_cov, *new_site_mean, *new_site_var;
  void* nlml;
};

// ---- the fused iteration (kjx_run and the sharded calls) ----------------------------------------------
template <typename T, int KIND>
struct Fused {
  static constexpr int D = BlockDim<KIND>::D;
  FusedArgs<T, D> fa;
  SmallLayout<T, D> lay;
  const kjx_model_t* m;
  int64_t ngroups, nthread_blocks;
  Fused(const kjx_model_t* m_, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var, void* ws)
      : lay(N, !gauss_ep(m_)), m(m_) {
    memset(&fa, 0, sizeof(fa));
    fill_prior<T, D>(fa.a, m);
    bind_fused<T, D>(fa, lay, ws);
    fa.a.N = N; fa.a.dt = dt; fa.a.y = y; fa.a.site_mean = site_mean; fa.a.site_var = site_var;
    fa.a.gaussian_init = (site_mean == nullptr) ? 1 : 0;
    fa.a.is_last_shard = 1;
    nthread_blocks = lay.fblocks;
    ngroups = 0;
    const size_t al = sizeof(T) * 4;
    fa.vec_ok = (lay.plan.L1 % 4 == 0) && (lay.plan.L2 % 4 == 0) && aligned_to(dt, al) && aligned_to(y, al) && aligned_to(site_mean, al) && aligned_to(site_var, al);
  }
  unsigned grid() const { return (unsigned)lay.fblocks; }
  unsigned threads() const { return (unsigned)lay.threads; }
  // total_out: where the top-level scan writes the shard's element (default: the workspace slot)
  void reduce(cudaStream_t st, T* total_out = nullptr) {
    if (fa.vec_ok) fused_fold_kernel<T, KIND, true><<<grid(), threads(), 0, st>>>(fa);
    else fused_fold_kernel<T, KIND, false><<<grid(), threads(), 0, st>>>(fa);
    // three levels of segment scans; the top level's single total is the shard's element
    for (int l = 0; l < 3; ++l) {
      const int seg = l ? fa.seg1 : fa.seg0;
      const unsigned nseg = (unsigned)((fa.lvl[l].n + seg - 1) / seg);
      T* tot = (l < 2) ? fa.lvl[l + 1].items : (total_out ? total_out : fa.a.f_shard_total);
      const size_t tstride = (l < 2) ? fa.lvl[l + 1].stride : 1;
      launch_segscan(seg, nseg, fa.lvl[l], tot, tstride, st);
    }
  }
  void launch_segscan(int seg, unsigned nseg, const ScanLevel<T, D>& lv, T* tot, size_t tstride, cudaStream_t st) {
    const size_t smem = sizeof(T) * FSum<T, D>::NS * seg;
    if (seg == 32) fused_segscan_kernel<T, D, 32><<<nseg, 64, smem, st>>>(lv, tot, tstride, fa.seg1);
    else if (seg == 64) fused_segscan_kernel<T, D, 64><<<nseg, 64, smem, st>>>(lv, tot, tstride, fa.seg1);
    else fused_segscan_kernel<T, D, 128><<<nseg, 64, smem, st>>>(lv, tot, tstride, fa.seg1);
  }
  // sharded iteration: fold, level-0 scan, then levels 1 + 2 and the publication of the shard's element in one launch
  void reduce_linked(cudaStream_t st) {
    if (fa.vec_ok) fused_fold_kernel<T, KIND, true><<<grid(), threads(), 0, st>>>(fa);
    else fused_fold_kernel<T, KIND, false><<<grid(), threads(), 0, st>>>(fa);
    const unsigned nseg0 = (unsigned)((fa.lvl[0].n + fa.seg0 - 1) / fa.seg0), nseg1 = (unsigned)((fa.lvl[1].n + fa.seg1 - 1) / fa.seg1);
    launch_segscan(fa.seg0, nseg0, fa.lvl[0], fa.lvl[1].items, fa.lvl[1].stride, st);
    const size_t smem = sizeof(T) * FSum<T, D>::NS * fa.seg1;
    if (fa.seg1 == 32) fused_segscan_top_kernel<T, D, 32, true><<<nseg1, 64, smem, st>>>(fa.lvl[1], fa.lvl[2], fa.a.f_shard_total, fa.link.ctrl + 1, fa.link);
    else if (fa.seg1 == 64) fused_segscan_top_kernel<T, D, 64, true><<<nseg1, 64, smem, st>>>(fa.lvl[1], fa.lvl[2], fa.a.f_shard_total, fa.link.ctrl + 1, fa.link);
    else fused_segscan_top_kernel<T, D, 128, true><<<nseg1, 64, smem, st>>>(fa.lvl[1], fa.lvl[2], fa.a.f_shard_total, fa.link.ctrl + 1, fa.link);
  }
  void filter_linked(cudaStream_t st) {
    const bool fast = gauss_ep(m);
    fa.store_xf = 1;
    const unsigned g1 = grid() + 1;                         // one extra block: folds the later shards' elements for the smoother
    if (fast && fa.vec_ok) fused_filter_kernel<T, KIND, true, true, true, true><<<g1, threads(), 0, st>>>(fa);
    else if (fast) fused_filter_kernel<T, KIND, true, false, true, true><<<g1, threads(), 0, st>>>(fa);
    else if (fa.vec_ok) fused_filter_kernel<T, KIND, false, true, true, true><<<g1, threads(), 0, st>>>(fa);
    else fused_filter_kernel<T, KIND, false, false, true, true><<<g1, threads(), 0, st>>>(fa);
  }
  void smoother_linked(T* post_mean, T* post_var, T* new_site_mean, T* new_site_var, bool update, cudaStream_t st) {
    fa.a.smooth_mean = post_mean; fa.a.smooth_cov = post_var; fa.a.new_site_mean = new_site_mean; fa.a.new_site_var = new_site_var;
    fa.a.update_sites = (update && new_site_mean != nullptr) ? 1 : 0;
    const size_t al = sizeof(T) * 4;
    const bool vec = fa.vec_ok && aligned_to(post_mean, al) && aligned_to(post_var, al) && aligned_to(new_site_mean, al) && aligned_to(new_site_var, al);
    const bool fast = gauss_ep(m);
    const unsigned g1 = grid() + 1;                         // one extra block: the nlml exchange
    if (fast && vec) fused_smoother_kernel<T, KIND, true, true, true><<<g1, threads(), 0, st>>>(fa);
    else if (fast) fused_smoother_kernel<T, KIND, true, false, true><<<g1, threads(), 0, st>>>(fa);
    else if (vec) fused_smoother_kernel<T, KIND, false, true, true><<<g1, threads(), 0, st>>>(fa);
    else fused_smoother_kernel<T, KIND, false, false, true><<<g1, threads(), 0, st>>>(fa);
  }
  // plain: the run() iteration; otherwise mask / energy-only / site write-back are honoured
  void filter(cudaStream_t st, bool plain = true) {
    const bool fast = gauss_ep(m);
    if (plain) {
      fa.store_xf = 1;
      if (fast && fa.vec_ok) fused_filter_kernel<T, KIND, true, true, true><<<grid(), threads(), 0, st>>>(fa);
      else if (fast) fused_filter_kernel<T, KIND, true, false, true><<<grid(), threads(), 0, st>>>(fa);
      else if (fa.vec_ok) fused_filter_kernel<T, KIND, false, true, true><<<grid(), threads(), 0, st>>>(fa);
      else fused_filter_kernel<T, KIND, false, false, true><<<grid(), threads(), 0, st>>>(fa);
    } else {
      if (fast && fa.vec_ok) fused_filter_kernel<T, KIND, true, true, false><<<grid(), threads(), 0, st>>>(fa);
      else if (fast) fused_filter_kernel<T, KIND, true, false, false><<<grid(), threads(), 0, st>>>(fa);
      else if (fa.vec_ok) fused_filter_kernel<T, KIND, false, true, false><<<grid(), threads(), 0, st>>>(fa);
      else fused_filter_kernel<T, KIND, false, false, false><<<grid(), threads(), 0, st>>>(fa);
    }
  }
  void smoother(T* post_mean, T* post_var, T* new_site_mean, T* new_site_var, bool update, cudaStream_t st) {
    fa.a.smooth_mean = post_mean; fa.a.smooth_cov = post_var; fa.a.new_site_mean = new_site_mean; fa.a.new_site_var = new_site_var;
    fa.a.update_sites = (update && new_site_mean != nullptr) ? 1 : 0;
    const size_t al = sizeof(T) * 4;
    const bool vec = fa.vec_ok && aligned_to(post_mean, al) && aligned_to(post_var, al) && aligned_to(new_site_mean, al) && aligned_to(new_site_var, al);
    const bool fast = gauss_ep(m);
    if (fast && vec) fused_smoother_kernel<T, KIND, true, true><<<grid(), threads(), 0, st>>>(fa);
    else if (fast) fused_smoother_kernel<T, KIND, true, false><<<grid(), threads(), 0, st>>>(fa);
    else if (vec) fused_smoother_kernel<T, KIND, false, true><<<grid(), threads(), 0, st>>>(fa);
    else fused_smoother_kernel<T, KIND, false, false><<<grid(), threads(), 0, st>>>(fa);
  }
};

// private packed filtered moments -> the reference's dense [N,d,1] / [N,d,d] layout (optional output of kjx_run)
template <typename T, int D>
__global__ void unpack_filtered_kernel(const T* __restrict__ xf, int64_t N, int L, int L2, int64_t nbig, T* __restrict__ fm, T* __restrict__ fP) {
  constexpr int NP = Packed<D>::NP;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  const int64_t kb = nbig * L;                               // steps covered by the chunks of length L
  const int64_t c = (k < kb) ? k / L : nbig + (k - kb) / L2, i = (k < kb) ? k % L : (k - kb) % L2;
  const T* base = xf + (((c >> 5) * (int64_t)L + i) * NP) * 32 + (c & 31);
  T m[D], P[D][D];
  int e = 0;
  for (int a = 0; a < D; ++a) m[a] = base[32 * e++];
  for (int a = 0; a < D; ++a) for (int b = a; b < D; ++b) { P[a][b] = base[32 * e]; P[b][a] = base[32 * e]; ++e; }
  for (int a = 0; a < D; ++a) fm[k * D + a] = m[a];
  for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) fP[k * D * D + a * D + b] = P[a][b];
}

template <typename T, int KIND>
int small_smoother_apply(const kjx_model_t* m, SmallArgs<T, BlockDim<KIND>::D>& a, cudaStream_t st) {
  if (gauss_ep(m)) smoother_apply_kernel<T, KIND, true><<<(unsigned)a.nblocks, KJX_BLOCK, 0, st>>>(a);
  else smoother_apply_kernel<T, KIND, false><<<(unsigned)a.nblocks, KJX_BLOCK, 0, st>>>(a);
  return KJX_OK;
}

template <typename T> __global__ void max_reduce_kernel(const T* __restrict__ v, int64_t n, T* __restrict__ out) {
  __shared__ T red[8];
  T m = T(0);
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const T x = v[i]; m = (x > m) ? x : m; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { const T o = __shfl_down_sync(0xffffffffu, m, off); m = (o > m) ? o : m; }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { T t = T(0); for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = (red[w] > t) ? red[w] : t; out[0] = t; }
}
template <typename T> __global__ void fill_kernel(T* a, T va, T* b, T vb, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = va; b[i] = vb; }
}

// Sites initialised from the predictive (sde_gp.py:287 with site_params=None) for a non-Gaussian model, as the fixed
// point of parallel-in-time sweeps (see fused_filter_kernel).  On success site_mean / site_var point at the converged
// sites (in the workspace) and the return value is the number of sweeps; 0 = did not converge (caller falls back to the
// one-chain kernel).  Synchronises the stream once per sweep (a scalar comes back to decide on the next one).
template <typename T, int KIND>
int jacobi_init_sites(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask, void* ws,
                      const T** site_mean, const T** site_var, cudaStream_t st) {
  constexpr int D = BlockDim<KIND>::D;
  static const int max_sweeps = [] { const char* e = getenv("KJX_INIT_SWEEPS"); return e ? atoi(e) : 400; }();
  if (max_sweeps <= 0) return 0;
  SmallLayout<T, D> lay(N, !gauss_ep(m));
  char* w = (char*)ws;
  T* buf = (T*)(w + lay.off_jac);
  T* sm[2] = {buf, buf + 2 * N}; T* sv[2] = {buf + N, buf + 3 * N};
  T* diff = (T*)(w + lay.off_jdiff);                         // [0]: the maximum; [32..]: per block
  // start: sites that say nothing (the first sweep then initialises every site from the prior's predictive)
  fill_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(sm[0], T(0), sv[0], T(1e12), N);
  const T tol = std::is_same<T, double>::value ? T(1e-13) : T(2e-6);
  T prev = T(1e30);
  for (int it = 0; it < max_sweeps; ++it) {
    const int cur = it & 1, nxt = cur ^ 1;
    Fused<T, KIND> f(m, N, dt, y, sm[cur], sv[cur], ws);
    f.fa.a.mask = mask; f.fa.a.out_site_mean = sm[nxt]; f.fa.a.out_site_var = sv[nxt];
    f.fa.store_xf = 0; f.fa.jacobi = 1; f.fa.block_diff = diff + 32;
    f.reduce(st);
    launch_set_prior_state<T, D>((T*)f.fa.a.f_state_in, m, st);
    f.filter(st, /*plain=*/false);
    max_reduce_kernel<T><<<1, 256, 0, st>>>(diff + 32, f.nthread_blocks, diff);
    T h = T(0);
    if (cudaMemcpyAsync(&h, diff, sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 0;
    if (h <= tol) { *site_mean = sm[nxt]; *site_var = sv[nxt]; return it + 1; }
    // the change cannot shrink below rounding noise: accept a plateau at that level, give up on a real stall
    if (it > 8 && h >= prev && h <= T(64) * tol) { *site_mean = sm[nxt]; *site_var = sv[nxt]; return it + 1; }
    prev = h;
  }
  return 0;
}

// the filter must run as one chain when sites have to be initialised from the predictive and that
// initialisation depends on the state (anything but Gaussian + EP, utils.py:241-244)
bool needs_sequential_filter(const kjx_model_t* m, bool have_sites) { return !have_sites && !gauss_ep(m); }

template <typename T, int KIND>
int small_filter(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask,
                 const T* site_mean, const T* site_var, T* filt_mean, T* filt_cov, T* out_site_mean,
                 T* out_site_var, T* nlml, void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  constexpr int D = BlockDim<KIND>::D;
  SmallLayout<T, D> lay(N, !gauss_ep(m));
  if (ws_bytes < lay.total || !ws) return KJX_ERR_WORKSPACE;
  SmallArgs<T, D> a; memset(&a, 0, sizeof(a));
  fill_prior<T, D>(a, m); bind_workspace<T, D>(a, lay, ws);
  a.N = N; a.dt = dt; a.y = y; a.mask = mask; a.site_mean = site_mean; a.site_var = site_var;
  a.filt_mean = filt_mean; a.filt_cov = filt_cov; a.out_site_mean = out_site_mean; a.out_site_var = out_site_var;
  a.is_last_shard = 1; a.dt_next = T(0); a.nlml_out = nlml;
  const bool have_sites = site_mean != nullptr;
  a.gaussian_init = (!have_sites && gauss_ep(m)) ? 1 : 0;
  if (N == 0) { cudaMemsetAsync(nlml, 0, sizeof(T), st); return KJX_OK; }
  bool chain = (flags & KJX_FLAG_SEQUENTIAL) || needs_sequential_filter(m, have_sites);
  if (chain && !(flags & KJX_FLAG_SEQUENTIAL)) {
    // sites to be initialised from the predictive: parallel-in-time sweeps to their fixed point, then the ordinary pass
    const T *jm = nullptr, *jv = nullptr;
    if (jacobi_init_sites<T, KIND>(m, N, dt, y, mask, ws, &jm, &jv, st) > 0) { site_mean = jm; site_var = jv; chain = false; }
  }
  if (chain) {
    launch_set_prior_state<T, D>((T*)a.f_state_in, m, st);
    const int init = (!have_sites && !a.gaussian_init) ? 1 : 0;
    filter_sequential_kernel<T, KIND><<<1, 32, 0, st>>>(a, init);
  } else {
    // parallel-in-time: the fused scan's fold / segment scans / filter (energy-only when no output is asked for)
    Fused<T, KIND> f(m, N, dt, y, site_mean, site_var, ws);
    f.fa.a.mask = mask; f.fa.a.out_site_mean = out_site_mean; f.fa.a.out_site_var = out_site_var;
    f.fa.store_xf = (filt_mean != nullptr) ? 1 : 0;
    f.reduce(st);
    launch_set_prior_state<T, D>((T*)f.fa.a.f_state_in, m, st);
    f.filter(st, /*plain=*/false);
    nlml_reduce_kernel<T><<<1, 256, 0, st>>>(f.fa.a.block_nlml, f.nthread_blocks, nlml);
    if (filt_mean) unpack_filtered_kernel<T, D><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(f.fa.xf, N, f.lay.plan.L1, f.lay.plan.L2, f.lay.plan.nbig, filt_mean, filt_cov);
  }
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

template <typename T, int KIND>
int small_smoother(const kjx_model_t* m, int64_t N, const T* dt, const T* filt_mean, const T* filt_cov,
                   const T* y, const T* site_mean, const T* site_var, T* smooth_mean, T* smooth_cov,
                   T* new_site_mean, T* new_site_var, void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  constexpr int D = BlockDim<KIND>::D;
  SmallLayout<T, D> lay(N, !gauss_ep(m));
  if (ws_bytes < lay.total || !ws) return KJX_ERR_WORKSPACE;
  SmallArgs<T, D> a; memset(&a, 0, sizeof(a));
  fill_prior<T, D>(a, m); bind_workspace<T, D>(a, lay, ws);
  a.N = N; a.dt = dt; a.y = y; a.site_mean = site_mean; a.site_var = site_var;
  a.filt_mean_in = filt_mean; a.filt_cov_in = filt_cov;
  a.smooth_mean = smooth_mean; a.smooth_cov = smooth_cov; a.new_site_mean = new_site_mean; a.new_site_var = new_site_var;
  a.return_full = (flags & KJX_FLAG_RETURN_FULL) ? 1 : 0;
  a.update_sites = (new_site_mean != nullptr && site_mean != nullptr && y != nullptr) ? 1 : 0;
  a.is_last_shard = 1; a.dt_next = T(0);
  if (N == 0) return KJX_OK;
  if (flags & KJX_FLAG_SEQUENTIAL) {
    smoother_sequential_kernel<T, KIND><<<1, 32, 0, st>>>(a);
  } else {
    smoother_reduce_kernel<T, KIND><<<(unsigned)a.nblocks, KJX_BLOCK, 0, st>>>(a);
    a.nlml_out = nullptr;
    smoother_blockscan_kernel<T, D><<<1, KJX_SCAN_BLOCK, 0, st>>>(a);
    cudaMemsetAsync((void*)a.s_state_in, 0, (D + D * D) * sizeof(T), st);
    small_smoother_apply<T, KIND>(m, a, st);
  }
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

template <typename T, int KIND>
int small_run(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var,
              T* filt_mean, T* filt_cov, T* new_site_mean, T* new_site_var, T* post_mean, T* post_var,
              T* nlml, void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st, const uint8_t* mask = nullptr) {
  constexpr int D = BlockDim<KIND>::D;
  SmallLayout<T, D> lay(N, !gauss_ep(m));
  if (ws_bytes < lay.total || !ws) return KJX_ERR_WORKSPACE;
  if (N == 0) { if (nlml) cudaMemsetAsync(nlml, 0, sizeof(T), st); return KJX_OK; }
  const bool have_sites = site_mean != nullptr;
  const bool seq = (flags & KJX_FLAG_SEQUENTIAL) != 0;
  const bool update = !(flags & KJX_FLAG_NO_SITE_UPDATE);
  bool chain = seq || needs_sequential_filter(m, have_sites);
  if (chain && !seq) {
    // first iteration of a non-Gaussian model: the sites initialised from the predictive, by parallel-in-time sweeps to
    // their fixed point; they are then the sites of an ordinary fused iteration (sde_gp.py:183-187)
    const T *jm = nullptr, *jv = nullptr;
    if (jacobi_init_sites<T, KIND>(m, N, dt, y, mask, ws, &jm, &jv, st) > 0) { site_mean = jm; site_var = jv; chain = false; }
  }
  if (chain) {
    // One dependent chain (reference order).  Needs dense filter output: use the caller's or scratch in ws is
    // not provisioned for it, so the caller must pass filt_mean/filt_cov on this path.
    if (!filt_mean || !filt_cov) return KJX_ERR_INVALID_ARG;
    SmallArgs<T, D> a; memset(&a, 0, sizeof(a));
    fill_prior<T, D>(a, m); bind_workspace<T, D>(a, lay, ws);
    a.N = N; a.dt = dt; a.y = y; a.mask = mask; a.site_mean = site_mean; a.site_var = site_var;
    a.filt_mean = filt_mean; a.filt_cov = filt_cov; a.filt_mean_in = filt_mean; a.filt_cov_in = filt_cov;
    a.smooth_mean = post_mean; a.smooth_cov = post_var; a.new_site_mean = new_site_mean; a.new_site_var = new_site_var;
    a.update_sites = update ? 1 : 0; a.is_last_shard = 1; a.nlml_out = nlml;
    a.gaussian_init = (!have_sites && gauss_ep(m)) ? 1 : 0;
    launch_set_prior_state<T, D>((T*)a.f_state_in, m, st);
    const int init = (!have_sites && !a.gaussian_init) ? 1 : 0;
    // freshly initialised sites land in new_site_* and are the smoother's "previous" sites (sde_gp.py:183-187)
    if (init) { a.out_site_mean = new_site_mean; a.out_site_var = new_site_var; }
    filter_sequential_kernel<T, KIND><<<1, 32, 0, st>>>(a, init);
    if (init) { a.site_mean = new_site_mean; a.site_var = new_site_var; a.out_site_mean = nullptr; a.out_site_var = nullptr; }
    if (seq) {
      smoother_sequential_kernel<T, KIND><<<1, 32, 0, st>>>(a);
    } else {
      smoother_reduce_kernel<T, KIND><<<(unsigned)a.nblocks, KJX_BLOCK, 0, st>>>(a);
      a.nlml_out = nullptr;
      smoother_blockscan_kernel<T, D><<<1, KJX_SCAN_BLOCK, 0, st>>>(a);
      cudaMemsetAsync((void*)a.s_state_in, 0, (D + D * D) * sizeof(T), st);
      small_smoother_apply<T, KIND>(m, a, st);
    }
  } else {
    Fused<T, KIND> f(m, N, dt, y, site_mean, site_var, ws);
    f.fa.a.mask = mask;
    f.reduce(st);
    launch_set_prior_state<T, D>((T*)f.fa.a.f_state_in, m, st);
    cudaMemsetAsync((void*)f.fa.s_info_in, 0, (D + D * D) * sizeof(T), st);
    if (mask) { f.fa.store_xf = 1; f.filter(st, /*plain=*/false); } else f.filter(st);
    if (nlml) nlml_reduce_kernel<T><<<1, 256, 0, st>>>(f.fa.a.block_nlml, f.nthread_blocks, nlml);
    f.smoother(post_mean, post_var, new_site_mean, new_site_var, update, st);
    if (filt_mean && filt_cov)
      unpack_filtered_kernel<T, D><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(f.fa.xf, N, f.lay.plan.L1, f.lay.plan.L2, f.lay.plan.nbig, filt_mean, filt_cov);
  }
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

// ---- sharded calls -----------------------------------------------------------------------------------
template <typename T, int KIND>
int shard_summary(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var,
                  T* summary_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int D = BlockDim<KIND>::D;
  if (ws_bytes < SmallLayout<T, D>(N, !gauss_ep(m)).total || !ws) return KJX_ERR_WORKSPACE;
  if (needs_sequential_filter(m, site_mean != nullptr)) return KJX_ERR_UNSUPPORTED;
  Fused<T, KIND> f(m, N, dt, y, site_mean, site_var, ws);
  f.reduce(st, summary_out);              // the top-level scan writes the shard's element straight to the caller
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
template <typename T, int KIND>
int shard_filter(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var,
                 const T* state_in, T* nlml_partial, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int D = BlockDim<KIND>::D;
  if (ws_bytes < SmallLayout<T, D>(N, !gauss_ep(m)).total || !ws) return KJX_ERR_WORKSPACE;
  Fused<T, KIND> f(m, N, dt, y, site_mean, site_var, ws);
  f.fa.a.f_state_in = state_in;           // read in place
  f.filter(st);
  nlml_reduce_kernel<T><<<1, 256, 0, st>>>(f.fa.a.block_nlml, f.nthread_blocks, nlml_partial);
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
template <typename T, int KIND>
int shard_smoother(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var,
                   const T* info_in, T* post_mean, T* post_var, T* new_site_mean, T* new_site_var, void* ws,
                   size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  constexpr int D = BlockDim<KIND>::D;
  if (ws_bytes < SmallLayout<T, D>(N, !gauss_ep(m)).total || !ws) return KJX_ERR_WORKSPACE;
  Fused<T, KIND> f(m, N, dt, y, site_mean, site_var, ws);
  f.fa.s_info_in = info_in;               // read in place
  f.smoother(post_mean, post_var, new_site_mean, new_site_var, !(flags & KJX_FLAG_NO_SITE_UPDATE), st);
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

// one whole iteration of a shard with the exchange inside the kernels (kjx_sharded_iteration_f64)
template <int KIND>
int shard_iteration(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const double* site_mean, const double* site_var,
                    const ShardLink& link, double* post_mean, double* post_var, double* new_site_mean, double* new_site_var,
                    void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  constexpr int D = BlockDim<KIND>::D;
  if (ws_bytes < SmallLayout<double, D>(N, !gauss_ep(m)).total || !ws) return KJX_ERR_WORKSPACE;
  if (needs_sequential_filter(m, site_mean != nullptr)) return KJX_ERR_UNSUPPORTED;
  Fused<double, KIND> f(m, N, dt, y, site_mean, site_var, ws);
  f.fa.link = link;
  f.reduce_linked(st);
  f.filter_linked(st);
  f.smoother_linked(post_mean, post_var, new_site_mean, new_site_var, !(flags & KJX_FLAG_NO_SITE_UPDATE), st);
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

// the fold: filtered state entering shard `rank` and information of the shards after it (host and device)
template <int D>
__host__ __device__ void fold_shards(const double* Pinf, const double* summaries, int64_t stride, int world, int rank,
                                     double* state_out, double* info_out) {
  FiltState<double, D> x;
  for (int i = 0; i < D; ++i) { x.m[i] = 0.0; for (int j = 0; j < D; ++j) x.P[i][j] = Pinf[i * D + j]; }   // sde_gp.py:265
  for (int r = 0; r < rank; ++r) {
    FSum<double, D> s; fsum_unpack<double, D>(summaries + (size_t)r * stride, s);
    fsum_apply<double, D>(s, x);
  }
  Info<double, D> f; info_zero<double, D>(f);
  for (int r = world - 1; r > rank; --r) {
    FSum<double, D> s; fsum_unpack<double, D>(summaries + (size_t)r * stride, s);
    fsum_combine_info<double, D>(s, f, f);
  }
  for (int i = 0; i < D; ++i) { state_out[i] = x.m[i]; info_out[i] = f.eta[i]; }
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) { state_out[D + i * D + j] = x.P[i][j]; info_out[D + i * D + j] = f.J[i][j]; }
}
template <int D> struct PinfArg { double v[D * D]; };

// Exchange + fold over peer memory.  One block; thread r < world talks to peer r.
template <int D>
__global__ void exchange_fold_kernel(PinfArg<D> pinf, const double* __restrict__ mine, double* const* __restrict__ peer_bufs,
                                     int64_t stride, int world, int rank, unsigned long long* epoch, double* state_out,
                                     double* info_out) {
  constexpr int NS = FSum<double, D>::NS;
  __shared__ double gathered[8 * 64];       // world <= 8 here, stride <= 64
  const unsigned long long e = *epoch + 1ull;
  const int par = (int)(e & 1ull);
  const int r = threadIdx.x;
  if (r < world) {
    // 1. my element -> peer r's receive slot [par][rank]   (remote stores over NVLink; local for r == rank)
    double* dst = peer_bufs[r] + ((size_t)par * world + rank) * stride;
    for (int k = 0; k < NS; ++k) dst[k] = mine[k];
    __threadfence_system();
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(peer_bufs[r] + (size_t)2 * world * stride) + (size_t)par * world + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(e) : "memory");
    // 2. wait for peer r's element in MY buffer
    const unsigned long long* myflag = reinterpret_cast<const unsigned long long*>(peer_bufs[rank] + (size_t)2 * world * stride) + (size_t)par * world + r;
    unsigned long long seen;
    do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(myflag) : "memory"); } while (seen < e);
    const double* src = peer_bufs[rank] + ((size_t)par * world + r) * stride;
    for (int k = 0; k < NS; ++k) {
      double v;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src + k) : "memory");
      gathered[r * 64 + k] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    fold_shards<D>(pinf.v, gathered, 64, world, rank, state_out, info_out);
    *epoch = e;
  }
}
template <int D>
__global__ void fold_shards_kernel(PinfArg<D> pinf, const double* summaries, int64_t stride, int world, int rank,
                                   double* state_out, double* info_out) {
  if (threadIdx.x == 0) fold_shards<D>(pinf.v, summaries, stride, world, rank, state_out, info_out);
}

// ---- multi-latent path (func_dim = 2: Independent prior + HeteroscedasticNoise + EP), kjx_multi_kernels.cuh ---------
bool is_hetero(const kjx_model_t* m) { return m->likelihood == KJX_LIK_HETERO_SOFTPLUS || m->likelihood == KJX_LIK_HETERO_EXP; }
bool multi_ok(const kjx_model_t* m) {
  if (m->method == KJX_INF_SLEP && (m->cub_rule != 0 && m->cub_rule != 3 && m->cub_rule != 5)) return false;
  return m->func_dim == 2 && m->state_dim <= KJX_MULTI_DMAX && m->H && !m->H_steps && is_hetero(m) && m->method >= KJX_INF_EP && m->method <= KJX_INF_VI;
}
template <typename T> void multi_fill(MultiArgs<T>& g, const kjx_model_t* m, int64_t N) {
  memset(&g, 0, sizeof(g));
  g.disc.d = m->state_dim; g.disc.n_blocks = m->n_blocks;
  for (int b = 0; b < m->n_blocks; ++b) g.disc.blocks[b] = m->blocks[b];
  g.site = make_site(m); g.d = m->state_dim; g.N = N;
  for (int i = 0; i < m->state_dim * m->state_dim; ++i) g.Pinf[i] = m->Pinf[i];
  for (int i = 0; i < 2 * m->state_dim; ++i) g.H[i] = m->H[i];
}
template <typename T>
int multi_filter(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask, const T* site_mean, const T* site_cov,
                 T* filt_mean, T* filt_cov, T* out_site_mean, T* out_site_cov, T* nlml, cudaStream_t st) {
  if (N == 0) { cudaMemsetAsync(nlml, 0, sizeof(T), st); return KJX_OK; }
  MultiArgs<T> g; multi_fill<T>(g, m, N);
  g.dt = dt; g.y = y; g.mask = mask; g.site_mean = site_mean; g.site_cov = site_cov; g.init_sites = site_mean ? 0 : 1;
  g.filt_mean = filt_mean; g.filt_cov = filt_cov; g.out_site_mean = out_site_mean; g.out_site_cov = out_site_cov; g.nlml = nlml;
  multi_filter_kernel<T><<<1, 32, 0, st>>>(g);
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
template <typename T>
int multi_smoother(const kjx_model_t* m, int64_t N, const T* dt, const T* filt_mean, const T* filt_cov, const T* y, const T* site_mean,
                   const T* site_cov, T* smooth_mean, T* smooth_cov, T* new_site_mean, T* new_site_cov, uint32_t flags, cudaStream_t st) {
  if (N == 0) return KJX_OK;
  MultiArgs<T> g; multi_fill<T>(g, m, N);
  g.dt = dt; g.y = y; g.site_mean = site_mean; g.site_cov = site_cov; g.filt_mean_in = filt_mean; g.filt_cov_in = filt_cov;
  g.smooth_mean = smooth_mean; g.smooth_cov = smooth_cov; g.new_site_mean = new_site_mean; g.new_site_cov = new_site_cov;
  g.return_full = (flags & KJX_FLAG_RETURN_FULL) ? 1 : 0;
  g.update_sites = (new_site_mean != nullptr && site_mean != nullptr && y != nullptr) ? 1 : 0;
  multi_smoother_kernel<T><<<1, 32, 0, st>>>(g);
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
// one run() iteration: filter (dense moments from the caller) then smoother + site update; the sites the filter
// initialised (first iteration) land in new_site_* and are updated in place by the smoother (sde_gp.py:183-187)
template <typename T>
int multi_run(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_cov, T* filt_mean, T* filt_cov,
              T* new_site_mean, T* new_site_cov, T* post_mean, T* post_var, T* nlml, uint32_t flags, cudaStream_t st) {
  if (!filt_mean || !filt_cov) return KJX_ERR_INVALID_ARG;    // this path keeps the filtered moments dense
  int rc = multi_filter<T>(m, N, dt, y, nullptr, site_mean, site_cov, filt_mean, filt_cov, site_mean ? nullptr : new_site_mean,
                           site_mean ? nullptr : new_site_cov, nlml, st);
  if (rc != KJX_OK) return rc;
  const T* prev_m = site_mean ? site_mean : new_site_mean; const T* prev_c = site_mean ? site_cov : new_site_cov;
  const bool update = !(flags & KJX_FLAG_NO_SITE_UPDATE);
  return multi_smoother<T>(m, N, dt, filt_mean, filt_cov, y, prev_m, prev_c, post_mean, post_var, update ? new_site_mean : nullptr,
                           update ? new_site_cov : nullptr, 0, st);
}

// ---- typed front ends shared by the _f64 / _f32 entry points ------------------------------------------
template <typename T>
int filter_T(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask, const T* site_mean,
             const T* site_cov, T* filt_mean, T* filt_cov, T* out_site_mean, T* out_site_cov, T* nlml, void* ws,
             size_t ws_bytes, uint32_t flags, kjx_stream_t stream) {
  if (!model_run_ok(m) || N < 0 || !dt || !y || !nlml) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((filt_mean == nullptr) != (filt_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((out_site_mean == nullptr) != (out_site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->func_dim != 1 || is_hetero(m))
    return multi_ok(m) ? multi_filter<T>(m, N, dt, y, mask, site_mean, site_cov, filt_mean, filt_cov, out_site_mean, out_site_cov, nlml, st)
                       : KJX_ERR_UNSUPPORTED;
  switch (small_kind(m)) {
    case KJX_BLOCK_MATERN32: return small_filter<T, KJX_BLOCK_MATERN32>(m, N, dt, y, mask, site_mean, site_cov, filt_mean, filt_cov, out_site_mean, out_site_cov, nlml, ws, ws_bytes, flags, st);
    case KJX_BLOCK_MATERN52: return small_filter<T, KJX_BLOCK_MATERN52>(m, N, dt, y, mask, site_mean, site_cov, filt_mean, filt_cov, out_site_mean, out_site_cov, nlml, ws, ws_bytes, flags, st);
    default:
      if (!generic_ok(m)) return KJX_ERR_UNSUPPORTED;
      return generic_filter<T>(m, N, dt, y, mask, site_mean, site_cov, filt_mean, filt_cov, out_site_mean, out_site_cov, nlml, ws, ws_bytes, flags, st);
  }
}

template <typename T>
int smoother_T(const kjx_model_t* m, int64_t N, const T* dt, const T* filt_mean, const T* filt_cov, const T* y,
               const T* site_mean, const T* site_cov, T* smooth_mean, T* smooth_cov, T* new_site_mean,
               T* new_site_cov, void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream) {
  if (!model_run_ok(m) || N < 0 || !dt || !filt_mean || !filt_cov) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((smooth_mean == nullptr) != (smooth_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((new_site_mean == nullptr) != (new_site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->func_dim != 1 || is_hetero(m))
    return multi_ok(m) ? multi_smoother<T>(m, N, dt, filt_mean, filt_cov, y, site_mean, site_cov, smooth_mean, smooth_cov, new_site_mean, new_site_cov, flags, st)
                       : KJX_ERR_UNSUPPORTED;
  switch (small_kind(m)) {
    case KJX_BLOCK_MATERN32: return small_smoother<T, KJX_BLOCK_MATERN32>(m, N, dt, filt_mean, filt_cov, y, site_mean, site_cov, smooth_mean, smooth_cov, new_site_mean, new_site_cov, ws, ws_bytes, flags, st);
    case KJX_BLOCK_MATERN52: return small_smoother<T, KJX_BLOCK_MATERN52>(m, N, dt, filt_mean, filt_cov, y, site_mean, site_cov, smooth_mean, smooth_cov, new_site_mean, new_site_cov, ws, ws_bytes, flags, st);
    default:
      if (!generic_ok(m)) return KJX_ERR_UNSUPPORTED;
      return generic_smoother<T>(m, N, dt, filt_mean, filt_cov, y, site_mean, site_cov, smooth_mean, smooth_cov, new_site_mean, new_site_cov, ws, ws_bytes, flags, st);
  }
}

template <typename T>
int run_T(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_cov,
          T* filt_mean, T* filt_cov, T* new_site_mean, T* new_site_cov, T* post_mean, T* post_var, T* nlml,
          void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream) {
  if (!model_run_ok(m) || N < 0 || !dt || !y || !nlml) return KJX_ERR_INVALID_ARG;
  if ((filt_mean == nullptr) != (filt_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if (!new_site_mean || !new_site_cov) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((post_mean == nullptr) != (post_var == nullptr)) return KJX_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->func_dim != 1 || is_hetero(m))
    return multi_ok(m) ? multi_run<T>(m, N, dt, y, site_mean, site_cov, filt_mean, filt_cov, new_site_mean, new_site_cov, post_mean, post_var, nlml, flags, st)
                       : KJX_ERR_UNSUPPORTED;
  switch (small_kind(m)) {
    case KJX_BLOCK_MATERN32: return small_run<T, KJX_BLOCK_MATERN32>(m, N, dt, y, site_mean, site_cov, filt_mean, filt_cov, new_site_mean, new_site_cov, post_mean, post_var, nlml, ws, ws_bytes, flags, st);
    case KJX_BLOCK_MATERN52: return small_run<T, KJX_BLOCK_MATERN52>(m, N, dt, y, site_mean, site_cov, filt_mean, filt_cov, new_site_mean, new_site_cov, post_mean, post_var, nlml, ws, ws_bytes, flags, st);
    default:
      if (!generic_ok(m)) return KJX_ERR_UNSUPPORTED;
      return generic_run<T>(m, N, dt, y, site_mean, site_cov, filt_mean, filt_cov, new_site_mean, new_site_cov, post_mean, post_var, nlml, ws, ws_bytes, flags, st);
  }
}

template <typename T>
int posterior_T(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask, const T* site_mean,
                const T* site_cov, T* post_mean, T* post_var, T* nlml, void* ws, size_t ws_bytes, kjx_stream_t stream) {
  if (!model_run_ok(m) || N < 0 || !dt || !y || !post_mean || !post_var) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if (needs_sequential_filter(m, site_mean != nullptr)) return KJX_ERR_UNSUPPORTED;   // use kjx_filter + kjx_smoother
  if (m->func_dim != 1 || is_hetero(m)) return KJX_ERR_UNSUPPORTED;                   // likewise (multi-latent path)
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t flags = KJX_FLAG_NO_SITE_UPDATE;
  switch (small_kind(m)) {
    case KJX_BLOCK_MATERN32: return small_run<T, KJX_BLOCK_MATERN32>(m, N, dt, y, site_mean, site_cov, nullptr, nullptr, nullptr, nullptr, post_mean, post_var, nlml, ws, ws_bytes, flags, st, mask);
    case KJX_BLOCK_MATERN52: return small_run<T, KJX_BLOCK_MATERN52>(m, N, dt, y, site_mean, site_cov, nullptr, nullptr, nullptr, nullptr, post_mean, post_var, nlml, ws, ws_bytes, flags, st, mask);
    default:
      if (!generic_ok(m)) return KJX_ERR_UNSUPPORTED;
      return generic_run<T>(m, N, dt, y, site_mean, site_cov, nullptr, nullptr, nullptr, nullptr, post_mean, post_var, nlml, ws, ws_bytes, flags, st, mask);
  }
}

template <typename T>
int site_update_T(const kjx_model_t* m, int64_t N, const T* y, const T* post_mean, const T* post_var,
                  const T* sprev_mean, const T* sprev_var, T* log_lik, T* site_mean, T* site_var, kjx_stream_t stream) {
  if (!model_basic_ok(m) || N < 0 || !y || !post_mean || !post_var) return KJX_ERR_INVALID_ARG;
  if ((sprev_mean == nullptr) != (sprev_var == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_var == nullptr)) return KJX_ERR_INVALID_ARG;
  if (m->func_dim != 1 || is_hetero(m)) {
    if (m->func_dim != 2 || !is_hetero(m) || (m->method == KJX_INF_SLEP && m->cub_rule != 0 && m->cub_rule != 3 && m->cub_rule != 5)) return KJX_ERR_UNSUPPORTED;
    if (N == 0) return KJX_OK;
    multi_site_update_kernel<T><<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(make_site(m), N, y, post_mean, post_var, sprev_mean,
                                                                                              sprev_var, log_lik, site_mean, site_var);
    KJX_CUDA_CHECK(cudaGetLastError());
    return KJX_OK;
  }
  if ((site_mean == nullptr) != (site_var == nullptr)) return KJX_ERR_INVALID_ARG;
  if (N == 0) return KJX_OK;
  const int threads = 128;
  const int64_t blocks = (N + threads - 1) / threads;
  site_update_kernel<T><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(make_site(m), N, y, post_mean, post_var,
                                                                                 sprev_mean, sprev_var, log_lik, site_mean, site_var);
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

template <typename T>
int discretise_T(const kjx_model_t* m, int64_t N, const T* dt, T* A, T* Q, void* /*unused*/, kjx_stream_t stream) {
  if (!model_basic_ok(m) || N < 0 || !dt || !A) return KJX_ERR_INVALID_ARG;
  if (N == 0) return KJX_OK;
  DiscArgs da; memset(&da, 0, sizeof(da));
  da.d = m->state_dim; da.n_blocks = m->n_blocks;
  for (int b = 0; b < m->n_blocks; ++b) da.blocks[b] = m->blocks[b];
  const int threads = 128, nw = threads / 32;
  const size_t smem = (size_t)nw * m->state_dim * (KJX_ROW_W * sizeof(T) + sizeof(int));
  if (smem > 48 * 1024) return KJX_ERR_UNSUPPORTED;          // before anything is allocated
  // Pinf is a host pointer: stage it into a small device buffer owned by this call's stream order
  double* pinf_dev = nullptr;
  if (Q) {
    KJX_CUDA_CHECK(cudaMallocAsync((void**)&pinf_dev, sizeof(double) * m->state_dim * m->state_dim, (cudaStream_t)stream));
    KJX_CUDA_CHECK(cudaMemcpyAsync(pinf_dev, m->Pinf, sizeof(double) * m->state_dim * m->state_dim, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  }
  const int64_t blocks = (N + nw - 1) / nw;
  discretise_kernel<T><<<(unsigned)std::min<int64_t>(blocks, (int64_t)sm_count() * 16), threads, smem, (cudaStream_t)stream>>>(da, N, dt, A, Q, pinf_dev);
  if (pinf_dev) KJX_CUDA_CHECK(cudaFreeAsync(pinf_dev, (cudaStream_t)stream));
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

#if KJX_PART != 32
int kjx_version(void) { return 100; }
#endif

#if KJX_PART != 32
const char* kjx_status_string(int status) {
  switch (status) {
    case KJX_OK: return "ok";
    case KJX_ERR_INVALID_ARG: return "invalid argument";
    case KJX_ERR_UNSUPPORTED: return "unsupported model for this build";
    case KJX_ERR_WORKSPACE: return "workspace too small";
    case KJX_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}
#endif

#if KJX_PART != 32
int kjx_workspace_bytes(const kjx_model_t* m, int64_t N, int dtype_bytes, size_t* bytes) {
  if (!model_basic_ok(m) || N < 0 || !bytes || (dtype_bytes != 4 && dtype_bytes != 8)) return KJX_ERR_INVALID_ARG;
  if (m->func_dim != 1 || is_hetero(m)) { if (!multi_ok(m) && !(m->func_dim == 2 && is_hetero(m))) return KJX_ERR_UNSUPPORTED; *bytes = 256; return KJX_OK; }
  const int kind = small_kind(m);
  size_t total = 0;
  if (kind == KJX_BLOCK_MATERN32) total = dtype_bytes == 8 ? SmallLayout<double, 2>(N, !gauss_ep(m)).total : SmallLayout<float, 2>(N, !gauss_ep(m)).total;
  else if (kind == KJX_BLOCK_MATERN52) total = dtype_bytes == 8 ? SmallLayout<double, 3>(N, !gauss_ep(m)).total : SmallLayout<float, 3>(N, !gauss_ep(m)).total;
  else if (generic_ok(m)) total = dtype_bytes == 8 ? generic_ws_bytes<double>(m, N) : generic_ws_bytes<float>(m, N);
  else return KJX_ERR_UNSUPPORTED;
  *bytes = total;
  return KJX_OK;
}
#endif

#if KJX_PART != 32
int kjx_filter_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const uint8_t* mask,
                   const double* site_mean, const double* site_cov, double* filt_mean, double* filt_cov,
                   double* out_site_mean, double* out_site_cov, double* nlml, void* ws, size_t ws_bytes,
                   uint32_t flags, kjx_stream_t stream) {
  return filter_T<double>(m, N, dt, y, mask, site_mean, site_cov, filt_mean, filt_cov, out_site_mean, out_site_cov, nlml, ws, ws_bytes, flags, stream);
}
#endif
#if KJX_PART != 64
int kjx_filter_f32(const kjx_model_t* m, int64_t N, const float* dt, const float* y, const uint8_t* mask,
                   const float* site_mean, const float* site_cov, float* filt_mean, float* filt_cov,
                   float* out_site_mean, float* out_site_cov, float* nlml, void* ws, size_t ws_bytes,
                   uint32_t flags, kjx_stream_t stream) {
  return filter_T<float>(m, N, dt, y, mask, site_mean, site_cov, filt_mean, filt_cov, out_site_mean, out_site_cov, nlml, ws, ws_bytes, flags, stream);
}
#endif

#if KJX_PART != 32
int kjx_smoother_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* filt_mean, const double* filt_cov,
                     const double* y, const double* site_mean, const double* site_cov, double* smooth_mean,
                     double* smooth_cov, double* new_site_mean, double* new_site_cov, void* ws, size_t ws_bytes,
                     uint32_t flags, kjx_stream_t stream) {
  return smoother_T<double>(m, N, dt, filt_mean, filt_cov, y, site_mean, site_cov, smooth_mean, smooth_cov, new_site_mean, new_site_cov, ws, ws_bytes, flags, stream);
}
#endif
#if KJX_PART != 64
int kjx_smoother_f32(const kjx_model_t* m, int64_t N, const float* dt, const float* filt_mean, const float* filt_cov,
                     const float* y, const float* site_mean, const float* site_cov, float* smooth_mean,
                     float* smooth_cov, float* new_site_mean, float* new_site_cov, void* ws, size_t ws_bytes,
                     uint32_t flags, kjx_stream_t stream) {
  return smoother_T<float>(m, N, dt, filt_mean, filt_cov, y, site_mean, site_cov, smooth_mean, smooth_cov, new_site_mean, new_site_cov, ws, ws_bytes, flags, stream);
}
#endif

#if KJX_PART != 32
int kjx_run_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const double* site_mean,
                const double* site_cov, double* filt_mean, double* filt_cov, double* new_site_mean,
                double* new_site_cov, double* post_mean, double* post_var, double* nlml, void* ws, size_t ws_bytes,
                uint32_t flags, kjx_stream_t stream) {
  return run_T<double>(m, N, dt, y, site_mean, site_cov, filt_mean, filt_cov, new_site_mean, new_site_cov, post_mean, post_var, nlml, ws, ws_bytes, flags, stream);
}
#endif
#if KJX_PART != 64
int kjx_run_f32(const kjx_model_t* m, int64_t N, const float* dt, const float* y, const float* site_mean,
                const float* site_cov, float* filt_mean, float* filt_cov, float* new_site_mean, float* new_site_cov,
                float* post_mean, float* post_var, float* nlml, void* ws, size_t ws_bytes, uint32_t flags,
                kjx_stream_t stream) {
  return run_T<float>(m, N, dt, y, site_mean, site_cov, filt_mean, filt_cov, new_site_mean, new_site_cov, post_mean, post_var, nlml, ws, ws_bytes, flags, stream);
}
#endif

#if KJX_PART != 32
int kjx_posterior_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const uint8_t* mask,
                      const double* site_mean, const double* site_cov, double* post_mean, double* post_var, double* nlml,
                      void* ws, size_t ws_bytes, kjx_stream_t stream) {
  return posterior_T<double>(m, N, dt, y, mask, site_mean, site_cov, post_mean, post_var, nlml, ws, ws_bytes, stream);
}
#endif

#if KJX_PART != 32
int kjx_site_update_f64(const kjx_model_t* m, int64_t N, const double* y, const double* post_mean, const double* post_var,
                        const double* site_mean_prev, const double* site_var_prev, double* log_lik, double* site_mean,
                        double* site_var, kjx_stream_t stream) {
  return site_update_T<double>(m, N, y, post_mean, post_var, site_mean_prev, site_var_prev, log_lik, site_mean, site_var, stream);
}
#endif
#if KJX_PART != 64
int kjx_site_update_f32(const kjx_model_t* m, int64_t N, const float* y, const float* post_mean, const float* post_var,
                        const float* site_mean_prev, const float* site_var_prev, float* log_lik, float* site_mean,
                        float* site_var, kjx_stream_t stream) {
  return site_update_T<float>(m, N, y, post_mean, post_var, site_mean_prev, site_var_prev, log_lik, site_mean, site_var, stream);
}
#endif

#if KJX_PART != 32
int kjx_discretise_f64(const kjx_model_t* m, int64_t N, const double* dt, double* A, double* Q, kjx_stream_t stream) {
  return discretise_T<double>(m, N, dt, A, Q, nullptr, stream);
}
#endif
#if KJX_PART != 64
int kjx_discretise_f32(const kjx_model_t* m, int64_t N, const float* dt, float* A, float* Q, kjx_stream_t stream) {
  return discretise_T<float>(m, N, dt, A, Q, nullptr, stream);
}
#endif

#if KJX_PART != 32
size_t kjx_filter_summary_doubles(int32_t d) { return (size_t)(d * d + d + d * (d + 1) / 2 + d + d * (d + 1) / 2); }
#endif

#define KJX_SMALL_DISPATCH(call32, call52)                     \
  switch (small_kind(m)) {                                     \
    case KJX_BLOCK_MATERN32: return call32;                    \
    case KJX_BLOCK_MATERN52: return call52;                    \
    default: return KJX_ERR_UNSUPPORTED;                       \
  }

#if KJX_PART != 32
int kjx_filter_summary_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const double* site_mean,
                           const double* site_cov, double* summary_out, void* ws, size_t ws_bytes, kjx_stream_t stream) {
  if (!model_run_ok(m) || N <= 0 || !dt || !y || !summary_out) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  KJX_SMALL_DISPATCH((shard_summary<double, KJX_BLOCK_MATERN32>(m, N, dt, y, site_mean, site_cov, summary_out, ws, ws_bytes, st)),
                     (shard_summary<double, KJX_BLOCK_MATERN52>(m, N, dt, y, site_mean, site_cov, summary_out, ws, ws_bytes, st)))
}
#endif

#if KJX_PART != 32
int kjx_filter_apply_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const double* site_mean,
                         const double* site_cov, const double* state_in, double* nlml_partial, void* ws, size_t ws_bytes,
                         kjx_stream_t stream) {
  if (!model_run_ok(m) || N <= 0 || !dt || !y || !state_in || !nlml_partial) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  KJX_SMALL_DISPATCH((shard_filter<double, KJX_BLOCK_MATERN32>(m, N, dt, y, site_mean, site_cov, state_in, nlml_partial, ws, ws_bytes, st)),
                     (shard_filter<double, KJX_BLOCK_MATERN52>(m, N, dt, y, site_mean, site_cov, state_in, nlml_partial, ws, ws_bytes, st)))
}
#endif

#if KJX_PART != 32
int kjx_smoother_apply_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const double* site_mean,
                           const double* site_cov, const double* info_in, double* post_mean, double* post_var,
                           double* new_site_mean, double* new_site_cov, void* ws, size_t ws_bytes, uint32_t flags,
                           kjx_stream_t stream) {
  if (!model_run_ok(m) || N <= 0 || !dt || !y || !info_in) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((post_mean == nullptr) != (post_var == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((new_site_mean == nullptr) != (new_site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  KJX_SMALL_DISPATCH((shard_smoother<double, KJX_BLOCK_MATERN32>(m, N, dt, y, site_mean, site_cov, info_in, post_mean, post_var, new_site_mean, new_site_cov, ws, ws_bytes, flags, st)),
                     (shard_smoother<double, KJX_BLOCK_MATERN52>(m, N, dt, y, site_mean, site_cov, info_in, post_mean, post_var, new_site_mean, new_site_cov, ws, ws_bytes, flags, st)))
}
#endif

#if KJX_PART != 32
int kjx_fold_shard_summaries_host(const kjx_model_t* m, const double* summaries_host, int64_t stride, int32_t world,
                                  int32_t rank, double* state_out_host, double* info_out_host) {
  if (!model_run_ok(m) || !summaries_host || !state_out_host || !info_out_host || rank < 0 || world <= rank) return KJX_ERR_INVALID_ARG;
  if (stride < (int64_t)kjx_filter_summary_doubles(m->state_dim)) return KJX_ERR_INVALID_ARG;
  if (!small_kind(m)) return KJX_ERR_UNSUPPORTED;
  if (m->state_dim == 2) fold_shards<2>(m->Pinf, summaries_host, stride, world, rank, state_out_host, info_out_host);
  else fold_shards<3>(m->Pinf, summaries_host, stride, world, rank, state_out_host, info_out_host);
  return KJX_OK;
}
#endif

#if KJX_PART != 32
int kjx_fold_shard_summaries_f64(const kjx_model_t* m, const double* summaries, int64_t stride, int32_t world,
                                 int32_t rank, double* state_out, double* info_out, kjx_stream_t stream) {
  if (!model_run_ok(m) || !summaries || !state_out || !info_out || rank < 0 || world <= rank) return KJX_ERR_INVALID_ARG;
  if (stride < (int64_t)kjx_filter_summary_doubles(m->state_dim)) return KJX_ERR_INVALID_ARG;
  if (!small_kind(m)) return KJX_ERR_UNSUPPORTED;
  if (m->state_dim == 2) {
    PinfArg<2> p; for (int i = 0; i < 4; ++i) p.v[i] = m->Pinf[i];
    fold_shards_kernel<2><<<1, 32, 0, (cudaStream_t)stream>>>(p, summaries, stride, world, rank, state_out, info_out);
  } else {
    PinfArg<3> p; for (int i = 0; i < 9; ++i) p.v[i] = m->Pinf[i];
    fold_shards_kernel<3><<<1, 32, 0, (cudaStream_t)stream>>>(p, summaries, stride, world, rank, state_out, info_out);
  }
  return cudaGetLastError() == cudaSuccess ? KJX_OK : KJX_ERR_CUDA;
}
#endif

#if KJX_PART != 32
int kjx_exchange_fold_f64(const kjx_model_t* m, const double* my_summary, double* const* peer_bufs, int64_t stride,
                          int32_t world, int32_t rank, uint64_t* epoch, double* state_out, double* info_out,
                          kjx_stream_t stream) {
  if (!model_run_ok(m) || !my_summary || !peer_bufs || !epoch || !state_out || !info_out) return KJX_ERR_INVALID_ARG;
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return KJX_ERR_INVALID_ARG;
  if (stride < (int64_t)kjx_filter_summary_doubles(m->state_dim) || stride > 64) return KJX_ERR_INVALID_ARG;
  if (!small_kind(m)) return KJX_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->state_dim == 2) {
    PinfArg<2> p; for (int i = 0; i < 4; ++i) p.v[i] = m->Pinf[i];
    exchange_fold_kernel<2><<<1, 32, 0, st>>>(p, my_summary, peer_bufs, stride, world, rank, (unsigned long long*)epoch, state_out, info_out);
  } else {
    PinfArg<3> p; for (int i = 0; i < 9; ++i) p.v[i] = m->Pinf[i];
    exchange_fold_kernel<3><<<1, 32, 0, st>>>(p, my_summary, peer_bufs, stride, world, rank, (unsigned long long*)epoch, state_out, info_out);
  }
  return cudaGetLastError() == cudaSuccess ? KJX_OK : KJX_ERR_CUDA;
}
#endif

#if KJX_PART != 32
size_t kjx_shard_link_doubles(int32_t world, int64_t stride) {
  return (world < 1 || stride < 1) ? 0 : (size_t)2 * world * (size_t)stride + (size_t)6 * world;
}
int kjx_sharded_iteration_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const double* site_mean,
                              const double* site_cov, double* const* peer_bufs, double* my_buf, int64_t stride, int32_t world,
                              int32_t rank, uint64_t* ctrl, double* post_mean, double* post_var, double* new_site_mean, double* new_site_cov,
                              double* nlml, void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream) {
  if (!model_run_ok(m) || N <= 0 || !dt || !y || !peer_bufs || !my_buf || !ctrl || !nlml) return KJX_ERR_INVALID_ARG;
  if ((site_mean == nullptr) != (site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((post_mean == nullptr) != (post_var == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((new_site_mean == nullptr) != (new_site_cov == nullptr)) return KJX_ERR_INVALID_ARG;
  if (world < 1 || world > 32 || rank < 0 || rank >= world) return KJX_ERR_INVALID_ARG;
  if (stride < (int64_t)kjx_filter_summary_doubles(m->state_dim)) return KJX_ERR_INVALID_ARG;
  ShardLink link;
  link.peer_bufs = peer_bufs; link.my_buf = my_buf; link.ctrl = (unsigned long long*)ctrl; link.nlml_out = nlml; link.stride = stride;
  link.world = world; link.rank = rank;
  cudaStream_t st = (cudaStream_t)stream;
  KJX_SMALL_DISPATCH((shard_iteration<KJX_BLOCK_MATERN32>(m, N, dt, y, site_mean, site_cov, link, post_mean, post_var, new_site_mean, new_site_cov, ws, ws_bytes, flags, st)),
                     (shard_iteration<KJX_BLOCK_MATERN52>(m, N, dt, y, site_mean, site_cov, link, post_mean, post_var, new_site_mean, new_site_cov, ws, ws_bytes, flags, st)))
}
#endif

#if KJX_PART != 32
int kjx_run_host_scratch_bytes(const kjx_model_t* m, int64_t N, int dtype_bytes, size_t* bytes) {
  size_t ws = 0;
  int rc = kjx_workspace_bytes(m, N, dtype_bytes, &ws);
  if (rc != KJX_OK) return rc;
  const size_t d = (size_t)m->state_dim, n = (size_t)std::max<int64_t>(N, 1);
  // dt, y, site_mean, site_var, new_site_mean, new_site_var, post_mean, post_var (8 N) + nlml; plus a dense
  // filter output only when the model needs the one-chain filter (first iteration of non-Gaussian models)
  const bool dense = !gauss_ep(m);
  *bytes = round_up(ws, 256) + 10 * round_up(n * dtype_bytes, 256) +
           (dense ? round_up(n * d * dtype_bytes, 256) + round_up(n * d * d * dtype_bytes, 256) : 0) + 256;
  return KJX_OK;
}
#endif

#if KJX_PART != 32
int kjx_run_host_f64(const kjx_model_t* m, int64_t N, const double* dt_host, const double* y_host,
                     const double* site_mean_host, const double* site_cov_host, double* new_site_mean_host,
                     double* new_site_cov_host, double* post_mean_host, double* post_var_host, double* nlml_host,
                     void* dev_scratch, size_t dev_scratch_bytes, uint32_t flags, kjx_stream_t stream) {
  const bool resident = (flags & KJX_FLAG_RESIDENT_SITES) != 0;
  if (!model_run_ok(m) || N <= 0 || !dt_host || !y_host || !nlml_host || !dev_scratch) return KJX_ERR_INVALID_ARG;
  if ((site_mean_host == nullptr) != (site_cov_host == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((new_site_mean_host == nullptr) != (new_site_cov_host == nullptr)) return KJX_ERR_INVALID_ARG;
  if ((post_mean_host == nullptr) != (post_var_host == nullptr)) return KJX_ERR_INVALID_ARG;
  const bool keep = (flags & (KJX_FLAG_RESIDENT_SITES | KJX_FLAG_RESIDENT_INIT)) != 0;
  if (!new_site_mean_host && !keep) return KJX_ERR_INVALID_ARG;         // the new sites must go somewhere
  size_t need = 0, ws_bytes = 0;
  int rc = kjx_run_host_scratch_bytes(m, N, 8, &need);
  if (rc != KJX_OK) return rc;
  if (dev_scratch_bytes < need) return KJX_ERR_WORKSPACE;
  kjx_workspace_bytes(m, N, 8, &ws_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t d = (size_t)m->state_dim, n = (size_t)N, nb = round_up(n * 8, 256);
  char* p = (char*)dev_scratch;
  auto take = [&](size_t b) { char* r = p; p += round_up(b, 256); return r; };
  void* ws = take(ws_bytes);
  double *dt = (double*)take(nb), *y = (double*)take(nb), *sm = (double*)take(nb), *sv = (double*)take(nb);
  double *nsm = (double*)take(nb), *nsv = (double*)take(nb), *pm = (double*)take(nb), *pv = (double*)take(nb);
  const bool dense = !gauss_ep(m);
  double *fm = dense ? (double*)take(n * d * 8) : nullptr, *fP = dense ? (double*)take(n * d * d * 8) : nullptr;
  double* nlml = (double*)take(8);
  KJX_CUDA_CHECK(cudaMemcpyAsync(dt, dt_host, n * 8, cudaMemcpyHostToDevice, st));
  KJX_CUDA_CHECK(cudaMemcpyAsync(y, y_host, n * 8, cudaMemcpyHostToDevice, st));
  if (site_mean_host) {
    KJX_CUDA_CHECK(cudaMemcpyAsync(sm, site_mean_host, n * 8, cudaMemcpyHostToDevice, st));
    KJX_CUDA_CHECK(cudaMemcpyAsync(sv, site_cov_host, n * 8, cudaMemcpyHostToDevice, st));
  }
  // sites on the device: the ones just uploaded, or (resident) the ones the previous call left in the scratch
  const bool have_sites = site_mean_host != nullptr || resident;
  rc = kjx_run_f64(m, N, dt, y, have_sites ? sm : nullptr, have_sites ? sv : nullptr, fm, fP, nsm, nsv,
                   post_mean_host ? pm : nullptr, post_mean_host ? pv : nullptr, nlml, ws, ws_bytes,
                   flags & ~(uint32_t)(KJX_FLAG_RESIDENT_SITES | KJX_FLAG_RESIDENT_INIT), stream);
  if (rc != KJX_OK) return rc;
  if (new_site_mean_host) {
    KJX_CUDA_CHECK(cudaMemcpyAsync(new_site_mean_host, nsm, n * 8, cudaMemcpyDeviceToHost, st));
    KJX_CUDA_CHECK(cudaMemcpyAsync(new_site_cov_host, nsv, n * 8, cudaMemcpyDeviceToHost, st));
  }
  if (post_mean_host) {
    KJX_CUDA_CHECK(cudaMemcpyAsync(post_mean_host, pm, n * 8, cudaMemcpyDeviceToHost, st));
    KJX_CUDA_CHECK(cudaMemcpyAsync(post_var_host, pv, n * 8, cudaMemcpyDeviceToHost, st));
  }
  KJX_CUDA_CHECK(cudaMemcpyAsync(nlml_host, nlml, 8, cudaMemcpyDeviceToHost, st));
  if (keep) {                                                           // the new sites are the next call's sites
    KJX_CUDA_CHECK(cudaMemcpyAsync(sm, nsm, n * 8, cudaMemcpyDeviceToDevice, st));
    KJX_CUDA_CHECK(cudaMemcpyAsync(sv, nsv, n * 8, cudaMemcpyDeviceToDevice, st));
  }
  KJX_CUDA_CHECK(cudaStreamSynchronize(st));
  return KJX_OK;
}
#endif

#if KJX_PART != 32
int kjx_run_host_submit_f64(const kjx_model_t* m, int64_t N, const double* dt_host, const double* y_host, double* nlml_host,
                            void* dev_scratch, size_t dev_scratch_bytes, int32_t slot, kjx_stream_t copy_stream, kjx_stream_t stream) {
  if (!model_run_ok(m) || N <= 0 || !dt_host || !y_host || !nlml_host || !dev_scratch || (slot != 0 && slot != 1)) return KJX_ERR_INVALID_ARG;
  size_t need = 0, ws_bytes = 0;
  int rc = kjx_run_host_scratch_bytes(m, N, 8, &need);
  if (rc != KJX_OK) return rc;
  if (dev_scratch_bytes < need) return KJX_ERR_WORKSPACE;
  kjx_workspace_bytes(m, N, 8, &ws_bytes);
  cudaStream_t st = (cudaStream_t)stream, cs = (cudaStream_t)copy_stream;
  const size_t d = (size_t)m->state_dim, n = (size_t)N, nb = round_up(n * 8, 256);
  char* p = (char*)dev_scratch;
  auto take = [&](size_t b) { char* r = p; p += round_up(b, 256); return r; };
  void* ws = take(ws_bytes);                                             // same carve-up as kjx_run_host_f64
  double *dt0 = (double*)take(nb), *y0 = (double*)take(nb), *sm = (double*)take(nb), *sv = (double*)take(nb);
  double *nsm = (double*)take(nb), *nsv = (double*)take(nb);
  take(nb); take(nb);                                                    // post_mean / post_var slots (unused here)
  const bool dense = !gauss_ep(m);
  double *fm = dense ? (double*)take(n * d * 8) : nullptr, *fP = dense ? (double*)take(n * d * d * 8) : nullptr;
  double* nlml = (double*)take(8);
  double *dt1 = (double*)take(nb), *y1 = (double*)take(nb);              // the second input slot
  double *dt = slot ? dt1 : dt0, *y = slot ? y1 : y0;
  // upload on the copy stream; the iteration waits for it on the compute stream
  cudaEvent_t up;
  KJX_CUDA_CHECK(cudaEventCreateWithFlags(&up, cudaEventDisableTiming));
  KJX_CUDA_CHECK(cudaMemcpyAsync(dt, dt_host, n * 8, cudaMemcpyHostToDevice, cs));
  KJX_CUDA_CHECK(cudaMemcpyAsync(y, y_host, n * 8, cudaMemcpyHostToDevice, cs));
  KJX_CUDA_CHECK(cudaEventRecord(up, cs));
  KJX_CUDA_CHECK(cudaStreamWaitEvent(st, up, 0));
  KJX_CUDA_CHECK(cudaEventDestroy(up));                                  // released once the wait has been satisfied
  rc = kjx_run_f64(m, N, dt, y, sm, sv, fm, fP, nsm, nsv, nullptr, nullptr, nlml + slot, ws, ws_bytes, 0, stream);
  if (rc != KJX_OK) return rc;
  KJX_CUDA_CHECK(cudaMemcpyAsync(nlml_host, nlml + slot, 8, cudaMemcpyDeviceToHost, st));
  KJX_CUDA_CHECK(cudaMemcpyAsync(sm, nsm, n * 8, cudaMemcpyDeviceToDevice, st));   // the new sites are the next call's sites
  KJX_CUDA_CHECK(cudaMemcpyAsync(sv, nsv, n * 8, cudaMemcpyDeviceToDevice, st));
  return KJX_OK;
}
#endif

}  // extern "C"
