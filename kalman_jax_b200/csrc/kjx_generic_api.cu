// kjx_generic_api.cu — host side of the generic (block-diagonal prior) path: workspace layout, chunking and the
// launch sequences over kjx_generic_kernels.cuh.  Declared in kjx_internal.cuh, used by kjx_api.cu.
#include "kjx_internal.cuh"
#include "kjx_generic_kernels.cuh"

namespace kjx_impl {

// ---- generic block-diagonal prior, d <= 64: CTA-level chunked scan (kjx_generic_kernels.cuh) ------------------
bool generic_ok(const kjx_model_t* m) {
  return m->func_dim == 1 && m->state_dim <= KJX_GEN_DMAX_GLOBAL && (m->H != nullptr || m->H_steps != nullptr);
}
// number of chunks: every chunk is walked concurrently by its own CTA (~55 us per step at d = 59 over fold, filter and
// smoother) while the two-level boundary walk costs ~22 us per chunk, so about sqrt(2.5 N) chunks balance the two; one
// wave of CTAs (the fold's and the smoother's working sets fill an SM's shared memory), so at most one chunk per SM
constexpr int KJX_GEN_MAXCHUNKS = 148 * 2;   // the workspace holds this many chunk elements / boundary states
int generic_chunks(int64_t N) {
  static const int knob = [] { const char* e = getenv("KJX_GEN_CHUNKS"); return e ? atoi(e) : 0; }();   // tuning knob, read once
  if (knob >= 1) return (int)std::min<int64_t>(std::min<int64_t>(knob, KJX_GEN_MAXCHUNKS), std::max<int64_t>(N, 1));
  const int64_t p = (int64_t)(sqrt(2.5 * (double)N) + 0.5);
  return (int)std::max<int64_t>(1, std::min<int64_t>(p, 148));
}
constexpr int KJX_GEN_MAXGROUPS = 64;      // two-level boundary walk: groups of ~sqrt(0.46 nchunks) chunks
constexpr int KJX_GEN_GAIN_GRID = 148 * 3;
// per-CTA working set when it lives in global memory (d > KJX_GEN_DMAX): the largest of the kernels' carve-ups
template <typename T> size_t generic_scratch_bytes(int d) {
  if (d <= KJX_GEN_DMAX) return 0;
  return round_up(std::max(std::max(std::max(generic_smoother_smem<T>(d), generic_fold_smem<T>(d)),
                                    std::max(generic_filter_smem<T>(d), generic_boundary_smem<T>(d))), generic_group_smem<T>(d)), 256);
}
// everything a call carves out of the workspace, in one place (sizes in bytes, 256-byte aligned)
template <typename T> struct GenericCarve {
  size_t pinf, h, mean, cov, gain, lc, gc, rows, c0n, pred, post, llpart, sum, stm, ine, stp, inj, nlml, gsum, scratch, total;
  GenericCarve(const kjx_model_t* m, int64_t N) {
    const size_t d = (size_t)m->state_dim, n = (size_t)std::max<int64_t>(N, 1), P = KJX_GEN_MAXCHUNKS;
    const GenLayout L = gen_layout(m->state_dim);
    const size_t tsz = d * (size_t)L.ld, gs = (size_t)gen_vec_stride<T>(m->state_dim);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += round_up(bytes, 256); return r; };
    pinf = take(d * d * 8); h = take(d * 8);
    mean = take(n * d * sizeof(T)); cov = take(n * d * d * sizeof(T));           // dense filter output (kjx_run without caller buffers)
    gain = take(n * tsz * sizeof(T)); lc = take(n * tsz * sizeof(T)); gc = take(n * gs * sizeof(T));   // RTS gains and affine terms, padded tiles
    rows = take((n + 1) * d * KJX_ROW_W * sizeof(T)); c0n = take(d * sizeof(int));       // rows of A(dt[n]) for every step
    pred = take(2 * n * sizeof(T)); post = take(2 * n * sizeof(T));              // predictive / smoothed marginals of every step
    llpart = take(((n + 127) / 128) * sizeof(T));
    sum = take(P * (3 * d * d + 2 * d) * sizeof(T));                             // chunk elements, states, informations, nlml
    stm = take(P * d * sizeof(T)); ine = take(P * d * sizeof(T));
    stp = take(P * d * d * sizeof(T)); inj = take(P * d * d * sizeof(T));
    nlml = take(P * sizeof(T));
    gsum = take(KJX_GEN_MAXGROUPS * (3 * d * d + 2 * d) * sizeof(T));
    scratch = take(generic_scratch_bytes<T>(m->state_dim) * KJX_GEN_GAIN_GRID);
    total = o;
  }
};
template <typename T> size_t generic_ws_bytes(const kjx_model_t* m, int64_t N) { return GenericCarve<T>(m, N).total; }

template <typename T> struct GenericBuffers { T* mean; T* cov; T* gain; T* post_mean; T* post_var; T* llpart; };

template <typename T>
int generic_setup(const kjx_model_t* m, int64_t N, void* ws, size_t ws_bytes, GenArgs<T>& g, GenericBuffers<T>& b, cudaStream_t st) {
  const GenericCarve<T> c(m, N);
  if (!ws || ws_bytes < c.total) return KJX_ERR_WORKSPACE;
  const size_t d = (size_t)m->state_dim, n = (size_t)std::max<int64_t>(N, 1);
  memset(&g, 0, sizeof(g));
  g.disc.d = m->state_dim; g.disc.n_blocks = m->n_blocks;
  for (int k = 0; k < m->n_blocks; ++k) g.disc.blocks[k] = m->blocks[k];
  g.site = make_site(m); g.d = m->state_dim; g.N = N;
  char* w = (char*)ws;
  double* pinf_dev = (double*)(w + c.pinf);
  double* h_dev = (double*)(w + c.h);
  b.mean = (T*)(w + c.mean); b.cov = (T*)(w + c.cov); b.gain = (T*)(w + c.gain);
  g.Lc = (T*)(w + c.lc); g.gc = (T*)(w + c.gc);
  g.rows = (T*)(w + c.rows); g.c0n_g = (int*)(w + c.c0n);
  g.roww = 4;
  for (int k = 0; k < m->n_blocks; ++k) if (block_size(m->blocks[k].kind) > 4) g.roww = KJX_ROW_W;
  g.pred_mean = nullptr; g.pred_var = nullptr;               // set by the caller when the log-likelihood is batched
  b.post_mean = (T*)(w + c.post); b.post_var = b.post_mean + n;
  b.llpart = (T*)(w + c.llpart);
  g.post_mean = b.post_mean; g.post_var = b.post_var;
  g.sum = (T*)(w + c.sum);
  g.st_m = (T*)(w + c.stm); g.in_eta = (T*)(w + c.ine);
  g.st_P = (T*)(w + c.stp); g.in_J = (T*)(w + c.inj);
  g.chunk_nlml = (T*)(w + c.nlml);
  g.gsum = (T*)(w + c.gsum);
  g.gscratch_stride = generic_scratch_bytes<T>(m->state_dim);
  g.gscratch = g.gscratch_stride ? (unsigned char*)(w + c.scratch) : nullptr;
  g.nchunks = 1; g.chunk_len = std::max<int64_t>(N, 1);
  KJX_CUDA_CHECK(cudaMemcpyAsync(pinf_dev, m->Pinf, d * d * 8, cudaMemcpyHostToDevice, st));
  g.Pinf = pinf_dev;
  if (m->H_steps) { g.H_steps = m->H_steps; g.H = nullptr; }
  else { KJX_CUDA_CHECK(cudaMemcpyAsync(h_dev, m->H, d * 8, cudaMemcpyHostToDevice, st)); g.H = h_dev; }
  (void)pinf_dev;
  return KJX_OK;
}
// predictive marginals for the batched log-likelihood live behind the smoothed ones' slot pair in the workspace
template <typename T> void generic_use_batched_loglik(GenArgs<T>& g, const kjx_model_t* m, int64_t N, void* ws) {
  const GenericCarve<T> c(m, N);
  g.pred_mean = (T*)((char*)ws + c.pred); g.pred_var = g.pred_mean + (size_t)std::max<int64_t>(N, 1);
}
// rows of A(dt[n]) for every step of the call (g.dt must be set)
template <typename T> int generic_launch_rows(const GenArgs<T>& g, cudaStream_t st) {
  const int64_t total = (g.N + 1) * (int64_t)g.d;
  generic_rows_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g.disc, g.d, g.roww, g.N, g.dt, const_cast<T*>(g.rows), const_cast<int*>(g.c0n_g));
  return KJX_OK;
}
// shared-memory build (tiles addressed as __shared__) or global-scratch build (d > 64) of a kernel template
#define KJX_GEN_LAUNCH(kernel, grid, threads, smem_bytes, st, ...)                                                        \
  do {                                                                                                                    \
    if (g.gscratch) {                                                                                                     \
      kernel<T, true><<<(grid), (threads), 0, (st)>>>(__VA_ARGS__);                                                       \
    } else {                                                                                                              \
      KJX_CUDA_CHECK(cudaFuncSetAttribute(kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_bytes))); \
      kernel<T, false><<<(grid), (threads), (smem_bytes), (st)>>>(__VA_ARGS__);                                           \
    }                                                                                                                     \
  } while (0)

template <typename T> void generic_set_chunks(GenArgs<T>& g, bool chunked) {
  g.nchunks = chunked ? generic_chunks(g.N) : 1;
  g.chunk_len = (g.N + g.nchunks - 1) / g.nchunks;
  g.nchunks = (int)((g.N + g.chunk_len - 1) / g.chunk_len);
}
// elements of all chunks (when there are several), then the state / information at every chunk boundary
template <typename T> int generic_launch_scan(const GenArgs<T>& g, cudaStream_t st) {
  if (g.nchunks > 1) KJX_GEN_LAUNCH(generic_fold_kernel, g.nchunks, KJX_GEN_THREADS, generic_fold_smem<T>(g.d), st, g);
  const size_t smem = generic_boundary_smem<T>(g.d);
  // a walk over the chunk elements is a dependent chain (~90 us per element at d = 59, ~105 us per combine): beyond a
  // few dozen chunks do it in two levels — group elements, a walk over the groups, walks inside all groups at once
  static const int group_knob = [] { const char* e = getenv("KJX_GEN_GROUP"); return e ? atoi(e) : 0; }();   // tuning knob, read once
  int gsz = (int)(sqrt(0.46 * (double)g.nchunks) + 0.5);
  if (group_knob > 0) gsz = group_knob;
  if (g.nchunks >= 24 && gsz >= 2 && (g.nchunks + gsz - 1) / gsz <= KJX_GEN_MAXGROUPS) {
    const int ngroups = (g.nchunks + gsz - 1) / gsz;
    KJX_GEN_LAUNCH(generic_group_kernel, ngroups, KJX_GEN_THREADS, generic_group_smem<T>(g.d), st, g, gsz);
    KJX_GEN_LAUNCH(generic_boundary_kernel, 2, KJX_GEN_THREADS, smem, st, g, (int)KJX_WALK_TOP, gsz);
    KJX_GEN_LAUNCH(generic_boundary_kernel, 2 * ngroups, KJX_GEN_THREADS, smem, st, g, (int)KJX_WALK_INNER, gsz);
  } else {
    KJX_GEN_LAUNCH(generic_boundary_kernel, 2, KJX_GEN_THREADS, smem, st, g, (int)KJX_WALK_FLAT, 1);
  }
  return KJX_OK;
}
template <typename T> int generic_launch_filter(const GenArgs<T>& g, T* llpart, cudaStream_t st) {
  int nblk = 0;
  for (int b = 0; b < g.disc.n_blocks; ++b) nblk += g.disc.blocks[b].count;
  static const bool tile_knob = getenv("KJX_GEN_TILE_FILTER") != nullptr;                              // tuning knob, read once
  if (nblk <= 16 && !tile_knob && !g.gscratch) {
    // at most 16 diagonal blocks: P in registers, one thread per pair of blocks
    const size_t smem = generic_filter_blocks_smem<T>(g.d);
    if (g.roww > 4) {               // a 6 x 6 block (Subband Matern-5/2): larger register tiles, 6-wide rows
      KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_filter_blocks_kernel<T, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      generic_filter_blocks_kernel<T, 6><<<g.nchunks, KJX_GEN_THREADS, smem, st>>>(g);
    } else {
      KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_filter_blocks_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      generic_filter_blocks_kernel<T, 4><<<g.nchunks, KJX_GEN_THREADS, smem, st>>>(g);
    }
  } else {
    KJX_GEN_LAUNCH(generic_filter_kernel, g.nchunks, KJX_GEN_THREADS, generic_filter_smem<T>(g.d), st, g);
  }
  if (g.pred_mean) {
    // steady state: the chains only recorded the predictive moments; all log Z terms at once, then their fixed-order sum
    const int64_t nb = (g.N + 127) / 128;
    generic_loglik_kernel<T><<<(unsigned)nb, 128, 0, st>>>(g.site, g.N, g.y, g.mask, g.pred_mean, g.pred_var, llpart);
    if (g.nlml) nlml_reduce_kernel<T><<<1, 256, 0, st>>>(llpart, nb, g.nlml);
  } else if (g.nlml) {
    nlml_reduce_kernel<T><<<1, 256, 0, st>>>(g.chunk_nlml, g.nchunks, g.nlml);
  }
  return KJX_OK;
}
template <typename T> int generic_launch_smoother(const GenArgs<T>& g, T* gains, cudaStream_t st) {
  // 1. all RTS gains and the affine terms of the recursion, parallel over time (they depend on the filtered moments only)
  static const int grid_knob = [] { const char* e = getenv("KJX_GEN_GAIN_CTAS"); return e ? atoi(e) : 0; }();   // tuning knob, read once
  int64_t ctas = KJX_GEN_GAIN_GRID;
  if (grid_knob >= 1) ctas = std::min<int64_t>(grid_knob, KJX_GEN_GAIN_GRID);
  const unsigned grid = (unsigned)std::min<int64_t>(g.N, ctas);
#define KJX_GAIN_LAUNCH(RW)                                                                                              \
  do {                                                                                                                  \
    if (g.gscratch) {                                                                                                   \
      generic_gain_kernel<T, true, RW><<<grid, KJX_GEN_GAIN_THREADS, 0, st>>>(g, gains);                                \
    } else {                                                                                                            \
      const size_t smem = generic_gain_smem<T>(g.d);                                                                    \
      KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_gain_kernel<T, false, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      generic_gain_kernel<T, false, RW><<<grid, KJX_GEN_GAIN_THREADS, smem, st>>>(g, gains);                            \
    }                                                                                                                   \
  } while (0)
  if (g.roww == 4) KJX_GAIN_LAUNCH(4); else KJX_GAIN_LAUNCH(KJX_ROW_W);
#undef KJX_GAIN_LAUNCH
  // 2. the backward recursion (two d^3 products per step), every chunk concurrently
  KJX_GEN_LAUNCH(generic_smoother_kernel, g.nchunks, KJX_GEN_THREADS, generic_smoother_smem<T>(g.d), st, g, (const T*)gains);
  // 3. the site update of every step from the smoothed marginals, embarrassingly parallel (sde_gp.py:373-379)
  if (g.update_sites)
    generic_site_update_kernel<T><<<(unsigned)((g.N + 127) / 128), 128, 0, st>>>(g.site, g.N, g.y, g.post_mean, g.post_var, g.site_mean,
                                                                              g.site_var, g.new_site_mean, g.new_site_var);
  return KJX_OK;
}

template <typename T>
int generic_filter(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask, const T* site_mean,
                   const T* site_var, T* filt_mean, T* filt_cov, T* out_site_mean, T* out_site_var, T* nlml, void* ws,
                   size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  GenArgs<T> g; GenericBuffers<T> b;
  int rc = generic_setup<T>(m, N, ws, ws_bytes, g, b, st);
  if (rc != KJX_OK) return rc;
  if (N == 0) { cudaMemsetAsync(nlml, 0, sizeof(T), st); return KJX_OK; }
  g.dt = dt; g.y = y; g.mask = mask; g.site_mean = site_mean; g.site_var = site_var;
  g.init_sites = (site_mean == nullptr && !gauss_ep(m)) ? 1 : 0;
  g.filt_mean = filt_mean; g.filt_cov = filt_cov; g.out_site_mean = out_site_mean; g.out_site_var = out_site_var; g.nlml = nlml;
  if (!g.init_sites) generic_use_batched_loglik<T>(g, m, N, ws);
  generic_set_chunks<T>(g, !g.init_sites && !(flags & KJX_FLAG_SEQUENTIAL));
  generic_launch_rows<T>(g, st);
  if ((rc = generic_launch_scan<T>(g, st)) != KJX_OK) return rc;
  if ((rc = generic_launch_filter<T>(g, b.llpart, st)) != KJX_OK) return rc;
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
template <typename T>
int generic_smoother(const kjx_model_t* m, int64_t N, const T* dt, const T* filt_mean, const T* filt_cov, const T* y,
                     const T* site_mean, const T* site_var, T* smooth_mean, T* smooth_cov, T* new_site_mean, T* new_site_var,
                     void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  GenArgs<T> g; GenericBuffers<T> b;
  int rc = generic_setup<T>(m, N, ws, ws_bytes, g, b, st);
  if (rc != KJX_OK) return rc;
  if (N == 0) return KJX_OK;
  g.dt = dt; g.y = y; g.site_mean = site_mean; g.site_var = site_var;
  g.filt_mean_in = filt_mean; g.filt_cov_in = filt_cov;
  g.smooth_mean = smooth_mean; g.smooth_cov = smooth_cov; g.new_site_mean = new_site_mean; g.new_site_var = new_site_var;
  g.return_full = (flags & KJX_FLAG_RETURN_FULL) ? 1 : 0;
  g.update_sites = (new_site_mean != nullptr && site_mean != nullptr && y != nullptr) ? 1 : 0;
  if (smooth_mean && !g.return_full) { g.post_mean = smooth_mean; g.post_var = smooth_cov; }   // H m, H P H^T are the outputs themselves
  // the stand-alone smoother only sees filtered moments (which mask / sites produced them is unknown here), so it
  // cannot rebuild the filter elements: one chunk, gains still computed in parallel
  generic_set_chunks<T>(g, false);
  generic_launch_rows<T>(g, st);
  if ((rc = generic_launch_smoother<T>(g, b.gain, st)) != KJX_OK) return rc;
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
template <typename T>
int generic_run(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var,
                T* filt_mean, T* filt_cov, T* new_site_mean, T* new_site_var, T* post_mean, T* post_var, T* nlml,
                void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st, const uint8_t* mask) {
  GenArgs<T> g; GenericBuffers<T> b;
  int rc = generic_setup<T>(m, N, ws, ws_bytes, g, b, st);
  if (rc != KJX_OK) return rc;
  if (N == 0) { if (nlml) cudaMemsetAsync(nlml, 0, sizeof(T), st); return KJX_OK; }
  if (!filt_mean) { filt_mean = b.mean; filt_cov = b.cov; }       // run() uses the filtered moments only internally
  g.dt = dt; g.y = y; g.mask = mask; g.site_mean = site_mean; g.site_var = site_var;
  g.init_sites = (site_mean == nullptr && !gauss_ep(m)) ? 1 : 0;
  g.filt_mean = filt_mean; g.filt_cov = filt_cov; g.nlml = nlml;
  if (g.init_sites) { g.out_site_mean = new_site_mean; g.out_site_var = new_site_var; }   // sde_gp.py:183
  else generic_use_batched_loglik<T>(g, m, N, ws);
  generic_set_chunks<T>(g, !g.init_sites && !(flags & KJX_FLAG_SEQUENTIAL));
  generic_launch_rows<T>(g, st);
  if ((rc = generic_launch_scan<T>(g, st)) != KJX_OK) return rc;
  if ((rc = generic_launch_filter<T>(g, b.llpart, st)) != KJX_OK) return rc;
  if (g.init_sites) {
    // sites were just initialised by the one-chain filter: rebuild the scan elements from them so the smoother can
    // run chunk-parallel too (the filter elements only depend on dt and the sites)
    g.site_mean = new_site_mean; g.site_var = new_site_var; g.out_site_mean = nullptr; g.out_site_var = nullptr; g.init_sites = 0;
    generic_set_chunks<T>(g, !(flags & KJX_FLAG_SEQUENTIAL));
    if ((rc = generic_launch_scan<T>(g, st)) != KJX_OK) return rc;
  }
  g.filt_mean_in = filt_mean; g.filt_cov_in = filt_cov;
  g.smooth_mean = post_mean; g.smooth_cov = post_var; g.new_site_mean = new_site_mean; g.new_site_var = new_site_var;
  if (post_mean) { g.post_mean = post_mean; g.post_var = post_var; }
  g.return_full = 0; g.update_sites = (flags & KJX_FLAG_NO_SITE_UPDATE) ? 0 : 1;
  if ((rc = generic_launch_smoother<T>(g, b.gain, st)) != KJX_OK) return rc;
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}


// the two scalar types the C ABI exposes
#define KJX_INSTANTIATE_GENERIC(T)                                                                                              \
  template size_t generic_ws_bytes<T>(const kjx_model_t*, int64_t);                                                              \
  template int generic_filter<T>(const kjx_model_t*, int64_t, const T*, const T*, const uint8_t*, const T*, const T*, T*, T*, T*, \
                                 T*, T*, void*, size_t, uint32_t, cudaStream_t);                                                 \
  template int generic_smoother<T>(const kjx_model_t*, int64_t, const T*, const T*, const T*, const T*, const T*, const T*, T*,  \
                                   T*, T*, T*, void*, size_t, uint32_t, cudaStream_t);                                           \
  template int generic_run<T>(const kjx_model_t*, int64_t, const T*, const T*, const T*, const T*, T*, T*, T*, T*, T*, T*, T*,   \
                              void*, size_t, uint32_t, cudaStream_t, const uint8_t*);
KJX_INSTANTIATE_GENERIC(double)
KJX_INSTANTIATE_GENERIC(float)

}  // namespace kjx_impl
