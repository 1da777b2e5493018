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
int generic_chunks(int64_t N) {
  if (const char* e = getenv("KJX_GEN_CHUNKS")) { const int v = atoi(e); if (v >= 1) return (int)std::min<int64_t>(v, std::max<int64_t>(N, 1)); }
  const int64_t p = (int64_t)(sqrt(2.5 * (double)N) + 0.5);
  return (int)std::max<int64_t>(1, std::min<int64_t>(p, 148));
}
constexpr int KJX_GEN_MAXCHUNKS = 148 * 2;
constexpr int KJX_GEN_MAXGROUPS = 64;      // two-level boundary walk: groups of ~sqrt(0.46 nchunks) chunks
constexpr int KJX_GEN_GAIN_GRID = 148 * 3;
// per-CTA working set when it lives in global memory (d > KJX_GEN_DMAX): the largest of the kernels' carve-ups
template <typename T> size_t generic_scratch_bytes(int d) {
  if (d <= KJX_GEN_DMAX) return 0;
  return round_up(std::max(std::max(std::max(generic_smoother_smem<T>(d), generic_fold_smem<T>(d)),
                                    std::max(generic_filter_smem<T>(d), generic_boundary_smem<T>(d))), generic_group_smem<T>(d)), 256);
}
template <typename T> size_t generic_ws_bytes(const kjx_model_t* m, int64_t N) {
  const size_t d = (size_t)m->state_dim, n = (size_t)std::max<int64_t>(N, 1), P = KJX_GEN_MAXCHUNKS;
  const size_t scratch = generic_scratch_bytes<T>(m->state_dim) * KJX_GEN_GAIN_GRID;
  // Pinf, H | dense filter output (kjx_run without caller buffers) | RTS gains [N,d,d] | chunk elements, states, infos, nlml
  return round_up(d * d * 8, 256) + round_up(d * 8, 256) + round_up(n * d * sizeof(T), 256) + 2 * round_up(n * d * d * sizeof(T), 256) +
         round_up(P * (3 * d * d + 2 * d) * sizeof(T), 256) + 2 * round_up(P * d * sizeof(T), 256) + 2 * round_up(P * d * d * sizeof(T), 256) +
         round_up(P * sizeof(T), 256) + round_up(KJX_GEN_MAXGROUPS * (3 * d * d + 2 * d) * sizeof(T), 256) + scratch;
}
template <typename T>
int generic_setup(const kjx_model_t* m, int64_t N, void* ws, size_t ws_bytes, GenArgs<T>& g, T** scratch_mean, T** scratch_cov,
                  T** scratch_gain, cudaStream_t st) {
  if (!ws || ws_bytes < generic_ws_bytes<T>(m, N)) return KJX_ERR_WORKSPACE;
  const size_t d = (size_t)m->state_dim, n = (size_t)std::max<int64_t>(N, 1), P = KJX_GEN_MAXCHUNKS;
  memset(&g, 0, sizeof(g));
  g.disc.d = m->state_dim; g.disc.n_blocks = m->n_blocks;
  for (int b = 0; b < m->n_blocks; ++b) g.disc.blocks[b] = m->blocks[b];
  g.site = make_site(m); g.d = m->state_dim; g.N = N;
  char* w = (char*)ws;
  auto take = [&](size_t bytes) { char* r = w; w += round_up(bytes, 256); return r; };
  double* pinf_dev = (double*)take(d * d * 8);
  double* h_dev = (double*)take(d * 8);
  *scratch_mean = (T*)take(n * d * sizeof(T));
  *scratch_cov = (T*)take(n * d * d * sizeof(T));
  *scratch_gain = (T*)take(n * d * d * sizeof(T));
  g.sum = (T*)take(P * (3 * d * d + 2 * d) * sizeof(T));
  g.st_m = (T*)take(P * d * sizeof(T)); g.in_eta = (T*)take(P * d * sizeof(T));
  g.st_P = (T*)take(P * d * d * sizeof(T)); g.in_J = (T*)take(P * d * d * sizeof(T));
  g.chunk_nlml = (T*)take(P * sizeof(T));
  g.gsum = (T*)take(KJX_GEN_MAXGROUPS * (3 * d * d + 2 * d) * sizeof(T));
  g.gscratch_stride = generic_scratch_bytes<T>(m->state_dim);
  g.gscratch = g.gscratch_stride ? (unsigned char*)take(g.gscratch_stride * KJX_GEN_GAIN_GRID) : nullptr;
  g.nchunks = 1; g.chunk_len = std::max<int64_t>(N, 1);
  KJX_CUDA_CHECK(cudaMemcpyAsync(pinf_dev, m->Pinf, d * d * 8, cudaMemcpyHostToDevice, st));
  g.Pinf = pinf_dev;
  if (m->H_steps) { g.H_steps = m->H_steps; g.H = nullptr; }
  else { KJX_CUDA_CHECK(cudaMemcpyAsync(h_dev, m->H, d * 8, cudaMemcpyHostToDevice, st)); g.H = h_dev; }
  return KJX_OK;
}
template <typename T> void generic_set_chunks(GenArgs<T>& g, bool chunked) {
  g.nchunks = chunked ? generic_chunks(g.N) : 1;
  g.chunk_len = (g.N + g.nchunks - 1) / g.nchunks;
  g.nchunks = (int)((g.N + g.chunk_len - 1) / g.chunk_len);
}
// elements of all chunks (when there are several), then the state / information at every chunk boundary
template <typename T> int generic_launch_scan(const GenArgs<T>& g, cudaStream_t st) {
  if (g.nchunks > 1) {
    const size_t smem = g.gscratch ? 0 : generic_fold_smem<T>(g.d);
    KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_fold_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_fold_kernel<T><<<g.nchunks, KJX_GEN_THREADS, smem, st>>>(g);
  }
  const size_t smem = g.gscratch ? 0 : generic_boundary_smem<T>(g.d);
  KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_boundary_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // a walk over the chunk elements is a dependent chain (~90 us per element at d = 59, ~105 us per combine): beyond a
  // few dozen chunks do it in two levels — group elements, a walk over the groups, walks inside all groups at once
  int gsz = (int)(sqrt(0.46 * (double)g.nchunks) + 0.5);
  if (const char* e = getenv("KJX_GEN_GROUP")) gsz = atoi(e);
  if (g.nchunks >= 24 && gsz >= 2 && (g.nchunks + gsz - 1) / gsz <= KJX_GEN_MAXGROUPS) {
    const int ngroups = (g.nchunks + gsz - 1) / gsz;
    const size_t gs = g.gscratch ? 0 : generic_group_smem<T>(g.d);
    KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_group_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs));
    generic_group_kernel<T><<<ngroups, KJX_GEN_THREADS, gs, st>>>(g, gsz);
    generic_boundary_kernel<T><<<2, KJX_GEN_THREADS, smem, st>>>(g, KJX_WALK_TOP, gsz);
    generic_boundary_kernel<T><<<2 * ngroups, KJX_GEN_THREADS, smem, st>>>(g, KJX_WALK_INNER, gsz);
  } else {
    generic_boundary_kernel<T><<<2, KJX_GEN_THREADS, smem, st>>>(g, KJX_WALK_FLAT, 1);
  }
  return KJX_OK;
}
template <typename T> int generic_launch_filter(const GenArgs<T>& g, cudaStream_t st) {
  int nblk = 0;
  for (int b = 0; b < g.disc.n_blocks; ++b) nblk += g.disc.blocks[b].count;
  if (nblk <= 16 && !getenv("KJX_GEN_TILE_FILTER")) {
    // at most 16 diagonal blocks: P in registers, one thread per pair of blocks
    const size_t smem = generic_filter_blocks_smem<T>(g.d);
    KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_filter_blocks_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_filter_blocks_kernel<T><<<g.nchunks, KJX_GEN_THREADS, smem, st>>>(g);
  } else {
    const size_t smem = g.gscratch ? 0 : generic_filter_smem<T>(g.d);
    KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_filter_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    generic_filter_kernel<T><<<g.nchunks, KJX_GEN_THREADS, smem, st>>>(g);
  }
  if (g.nlml) nlml_reduce_kernel<T><<<1, 256, 0, st>>>(g.chunk_nlml, g.nchunks, g.nlml);
  return KJX_OK;
}
template <typename T> int generic_launch_smoother(const GenArgs<T>& g, T* gains, cudaStream_t st) {
  const size_t smem = g.gscratch ? 0 : generic_smoother_smem<T>(g.d);
  // 1. all RTS gains, parallel over time (they depend on the filtered moments only)
  const size_t gsmem = g.gscratch ? 0 : generic_gain_smem<T>(g.d);
  KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_gain_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
  const unsigned grid = (unsigned)std::min<int64_t>(g.N, KJX_GEN_GAIN_GRID);
  generic_gain_kernel<T><<<grid, KJX_GEN_GAIN_THREADS, gsmem, st>>>(g, gains);
  // 2. the backward recursion (two d^3 products per step), every chunk concurrently, reference order inside a chunk
  KJX_CUDA_CHECK(cudaFuncSetAttribute(generic_smoother_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  generic_smoother_kernel<T><<<g.nchunks, KJX_GEN_THREADS, smem, st>>>(g, gains);
  return KJX_OK;
}

template <typename T>
int generic_filter(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask, const T* site_mean,
                   const T* site_var, T* filt_mean, T* filt_cov, T* out_site_mean, T* out_site_var, T* nlml, void* ws,
                   size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  GenArgs<T> g; T *sm, *sc, *sg;
  int rc = generic_setup<T>(m, N, ws, ws_bytes, g, &sm, &sc, &sg, st);
  if (rc != KJX_OK) return rc;
  if (N == 0) { cudaMemsetAsync(nlml, 0, sizeof(T), st); return KJX_OK; }
  g.dt = dt; g.y = y; g.mask = mask; g.site_mean = site_mean; g.site_var = site_var;
  g.init_sites = (site_mean == nullptr && !gauss_ep(m)) ? 1 : 0;
  g.filt_mean = filt_mean; g.filt_cov = filt_cov; g.out_site_mean = out_site_mean; g.out_site_var = out_site_var; g.nlml = nlml;
  generic_set_chunks<T>(g, !g.init_sites && !(flags & KJX_FLAG_SEQUENTIAL));
  if ((rc = generic_launch_scan<T>(g, st)) != KJX_OK) return rc;
  if ((rc = generic_launch_filter<T>(g, st)) != KJX_OK) return rc;
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
template <typename T>
int generic_smoother(const kjx_model_t* m, int64_t N, const T* dt, const T* filt_mean, const T* filt_cov, const T* y,
                     const T* site_mean, const T* site_var, T* smooth_mean, T* smooth_cov, T* new_site_mean, T* new_site_var,
                     void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st) {
  GenArgs<T> g; T *sm, *sc, *sg;
  int rc = generic_setup<T>(m, N, ws, ws_bytes, g, &sm, &sc, &sg, st);
  if (rc != KJX_OK) return rc;
  if (N == 0) return KJX_OK;
  g.dt = dt; g.y = y; g.site_mean = site_mean; g.site_var = site_var;
  g.filt_mean_in = filt_mean; g.filt_cov_in = filt_cov;
  g.smooth_mean = smooth_mean; g.smooth_cov = smooth_cov; g.new_site_mean = new_site_mean; g.new_site_var = new_site_var;
  g.return_full = (flags & KJX_FLAG_RETURN_FULL) ? 1 : 0;
  g.update_sites = (new_site_mean != nullptr && site_mean != nullptr && y != nullptr) ? 1 : 0;
  // the stand-alone smoother only sees filtered moments (which mask / sites produced them is unknown here), so it
  // cannot rebuild the filter elements: one chunk, gains still computed in parallel
  generic_set_chunks<T>(g, false);
  if ((rc = generic_launch_scan<T>(g, st)) != KJX_OK) return rc;
  if ((rc = generic_launch_smoother<T>(g, sg, st)) != KJX_OK) return rc;
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}
template <typename T>
int generic_run(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var,
                T* filt_mean, T* filt_cov, T* new_site_mean, T* new_site_var, T* post_mean, T* post_var, T* nlml,
                void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st, const uint8_t* mask) {
  GenArgs<T> g; T *sm, *sc, *sg;
  int rc = generic_setup<T>(m, N, ws, ws_bytes, g, &sm, &sc, &sg, st);
  if (rc != KJX_OK) return rc;
  if (N == 0) { if (nlml) cudaMemsetAsync(nlml, 0, sizeof(T), st); return KJX_OK; }
  if (!filt_mean) { filt_mean = sm; filt_cov = sc; }       // run() uses the filtered moments only internally
  g.dt = dt; g.y = y; g.mask = mask; g.site_mean = site_mean; g.site_var = site_var;
  g.init_sites = (site_mean == nullptr && !gauss_ep(m)) ? 1 : 0;
  g.filt_mean = filt_mean; g.filt_cov = filt_cov; g.nlml = nlml;
  if (g.init_sites) { g.out_site_mean = new_site_mean; g.out_site_var = new_site_var; }   // sde_gp.py:183
  generic_set_chunks<T>(g, !g.init_sites && !(flags & KJX_FLAG_SEQUENTIAL));
  if ((rc = generic_launch_scan<T>(g, st)) != KJX_OK) return rc;
  if ((rc = generic_launch_filter<T>(g, st)) != KJX_OK) return rc;
  if (g.init_sites) {
    // sites were just initialised by the one-chain filter: rebuild the scan elements from them so the smoother can
    // run chunk-parallel too (the filter elements only depend on dt and the sites)
    g.site_mean = new_site_mean; g.site_var = new_site_var; g.out_site_mean = nullptr; g.out_site_var = nullptr; g.init_sites = 0;
    generic_set_chunks<T>(g, !(flags & KJX_FLAG_SEQUENTIAL));
    if ((rc = generic_launch_scan<T>(g, st)) != KJX_OK) return rc;
  }
  g.filt_mean_in = filt_mean; g.filt_cov_in = filt_cov;
  g.smooth_mean = post_mean; g.smooth_cov = post_var; g.new_site_mean = new_site_mean; g.new_site_var = new_site_var;
  g.return_full = 0; g.update_sites = (flags & KJX_FLAG_NO_SITE_UPDATE) ? 0 : 1;
  if ((rc = generic_launch_smoother<T>(g, sg, st)) != KJX_OK) return rc;
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
