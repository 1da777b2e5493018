// kjx_internal.cuh — what the translation units of libkjx.so share: status macro, model inspection helpers and the
// entry points of the generic (block-diagonal prior) path, which is compiled on its own (kjx_generic_api.cu) so the
// library builds as three parallel nvcc jobs instead of one five-minute one.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>

#include "../../include/kjx.h"
#include "kjx_elementwise_kernels.cuh"

#define KJX_CUDA_CHECK(expr)                              \
  do {                                                    \
    cudaError_t e__ = (expr);                             \
    if (e__ != cudaSuccess) return KJX_ERR_CUDA;          \
  } while (0)

namespace kjx_impl {
using namespace kjx;

inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// model inspection
inline bool model_basic_ok(const kjx_model_t* m) {
  if (!m) return false;
  if (m->state_dim <= 0 || m->func_dim <= 0 || m->n_blocks <= 0 || m->n_blocks > KJX_MAX_BLOCKS) return false;
  if (m->n_cub < 1 || m->n_cub > KJX_MAX_CUBATURE) return false;
  if (m->likelihood < KJX_LIK_GAUSSIAN || m->likelihood > KJX_LIK_HETERO_EXP) return false;
  if (m->method < KJX_INF_EP || m->method > KJX_INF_MOMENT_MATCH) return false;
  if (!m->Pinf) return false;
  int d = 0;
  for (int b = 0; b < m->n_blocks; ++b) {
    const kjx_block_t& k = m->blocks[b];
    if (k.count <= 0) return false;
    int bs = block_size(k.kind);
    if (bs < 0) return false;
    d += bs * k.count;
  }
  return d == m->state_dim;
}

// filter / smoother entry points: KJX_INF_MOMENT_MATCH is no update rule (kjx_site_update only)
inline bool model_run_ok(const kjx_model_t* m) { return model_basic_ok(m) && m->method != KJX_INF_MOMENT_MATCH; }

inline SiteModel make_site(const kjx_model_t* m) {
  SiteModel s;
  memset(&s, 0, sizeof(s));
  s.lik = m->likelihood; s.method = m->method; s.n_cub = m->n_cub; s.cub_rule = m->cub_rule;
  s.lik_hyp = m->lik_hyp; s.power = m->power; s.damping = m->damping;
  for (int i = 0; i < KJX_MAX_CUBATURE; ++i) { s.cub_x[i] = m->cub_x[i]; s.cub_w[i] = m->cub_w[i]; }
  return s;
}

inline bool gauss_ep(const kjx_model_t* m) { return m->likelihood == KJX_LIK_GAUSSIAN && m->method == KJX_INF_EP; }

// ---- generic path (kjx_generic_api.cu) ----
bool generic_ok(const kjx_model_t* m);
template <typename T> size_t generic_ws_bytes(const kjx_model_t* m, int64_t N);
template <typename T>
int generic_filter(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const uint8_t* mask, const T* site_mean,
                   const T* site_var, T* filt_mean, T* filt_cov, T* out_site_mean, T* out_site_var, T* nlml, void* ws,
                   size_t ws_bytes, uint32_t flags, cudaStream_t st);
template <typename T>
int generic_smoother(const kjx_model_t* m, int64_t N, const T* dt, const T* filt_mean, const T* filt_cov, const T* y,
                     const T* site_mean, const T* site_var, T* smooth_mean, T* smooth_cov, T* new_site_mean, T* new_site_var,
                     void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st);
template <typename T>
int generic_run(const kjx_model_t* m, int64_t N, const T* dt, const T* y, const T* site_mean, const T* site_var,
                T* filt_mean, T* filt_cov, T* new_site_mean, T* new_site_var, T* post_mean, T* post_var, T* nlml,
                void* ws, size_t ws_bytes, uint32_t flags, cudaStream_t st, const uint8_t* mask = nullptr);

}  // namespace kjx_impl
