// kjx_sampling_api.cu — C ABI of the prior sampler (kjx_sampling_kernels.cuh).
#include "kjx_internal.cuh"
#include "kjx_sampling_kernels.cuh"

using namespace kjx_impl;

extern "C" {

/* ---- prior samples (sde_gp.py:386-417), d <= 64, constant measurement row ---- */
int kjx_prior_sample_workspace_bytes(const kjx_model_t* m, int64_t N, int32_t num_samps, size_t* bytes) {
  if (!model_basic_ok(m) || N < 0 || num_samps < 0 || !bytes) return KJX_ERR_INVALID_ARG;
  const size_t d = (size_t)m->state_dim, S = (size_t)std::min<int32_t>(std::max<int32_t>(num_samps, 1), KJX_SAMPLE_MAXS);
  *bytes = round_up(d * d * 8, 256) + round_up(d * 8, 256) + round_up((size_t)(N + 1) * d * S * 8, 256);
  return KJX_OK;
}

int kjx_prior_sample_f64(const kjx_model_t* m, int64_t N, const double* dt, int32_t num_samps, uint64_t seed,
                         double* f_sample, void* ws, size_t ws_bytes, kjx_stream_t stream) {
  if (!model_basic_ok(m) || N < 0 || num_samps < 0 || (N > 0 && num_samps > 0 && (!dt || !f_sample))) return KJX_ERR_INVALID_ARG;
  if (m->func_dim != 1 || !m->H || m->state_dim > 64) return KJX_ERR_UNSUPPORTED;   // the reference has no spatio-temporal sampling either (:432)
  if (N == 0 || num_samps == 0) return KJX_OK;
  size_t need = 0;
  kjx_prior_sample_workspace_bytes(m, N, num_samps, &need);
  if (!ws || ws_bytes < need) return KJX_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int d = m->state_dim;
  DiscArgs da; memset(&da, 0, sizeof(da));
  da.d = d; da.n_blocks = m->n_blocks;
  for (int b = 0; b < m->n_blocks; ++b) da.blocks[b] = m->blocks[b];
  char* p = (char*)ws;
  double* pinf_dev = (double*)p; p += round_up((size_t)d * d * 8, 256);
  double* h_dev = (double*)p; p += round_up((size_t)d * 8, 256);
  double* W = (double*)p;
  KJX_CUDA_CHECK(cudaMemcpyAsync(pinf_dev, m->Pinf, (size_t)d * d * 8, cudaMemcpyHostToDevice, st));
  KJX_CUDA_CHECK(cudaMemcpyAsync(h_dev, m->H, (size_t)d * 8, cudaMemcpyHostToDevice, st));
  for (int s0 = 0; s0 < num_samps; s0 += KJX_SAMPLE_MAXS) {
    const int S = std::min<int>(KJX_SAMPLE_MAXS, num_samps - s0);
    const size_t smem = sample_noise_smem(d, S);
    KJX_CUDA_CHECK(cudaFuncSetAttribute(sample_noise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_noise_kernel<<<(unsigned)std::min<int64_t>(N + 1, 148 * 8), KJX_SAMPLE_THREADS, smem, st>>>(da, N, dt, pinf_dev, S, s0, seed, W);
    const int nw = KJX_SAMPLE_THREADS / 32;
    sample_path_kernel<<<(unsigned)((S + nw - 1) / nw), KJX_SAMPLE_THREADS, (size_t)nw * 2 * d * sizeof(double), st>>>(
        da, N, dt, h_dev, W, S, s0, num_samps, f_sample);
  }
  KJX_CUDA_CHECK(cudaGetLastError());
  return KJX_OK;
}

}  // extern "C"
