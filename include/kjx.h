/*
 * kjx.h — C ABI of libkjx.so: the B200 (sm_100a) Kalman filter / RTS smoother / site-update path.
 *
 * This is the drop-in boundary for kalman-jax's shared inference core.  The reference has no
 * FFI/plugin layer of its own (it is pure Python on JAX); the seam these entry points replace is the
 * set of Python methods below (paths relative to /root/reference/kalmanjax):
 *
 *   kjx_filter_*        <- SDEGP.kalman_filter                      sde_gp.py:230-309
 *   kjx_smoother_*      <- SDEGP.rauch_tung_striebel_smoother       sde_gp.py:311-384
 *   kjx_run_*           <- SDEGP.run (filter -> smoother, no grad)  sde_gp.py:163-188
 *   kjx_site_update_*   <- ApproxInf.update / Likelihood.moment_match,
 *                          statistical_linear_regression, variational_expectation,
 *                          analytical_linearisation                 approximate_inference.py:36-323,
 *                                                                   likelihoods.py:122-321
 *                          (and the vmapped NLPD of sde_gp.py:139-142)
 *   kjx_discretise_*    <- Prior.state_transition closed forms      priors.py:163-175, 238-255,
 *                          and Q = Pinf - A Pinf A^T                1044-1084, 1398-1414, 1163-1181;
 *                                                                   sde_gp.py:409
 *   kjx_*_summary/_apply<- (new) the time-sharded parallel-in-time scan; no reference counterpart
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / JAX types.
 *   - Every array pointer is a DEVICE pointer unless its name ends in `_host` (the two small model
 *     matrices kjx_model_t.Pinf / .H are HOST pointers too); arrays are dense,
 *     row-major, in the reference's layouts (y[N,f], dt[N], site_mean[N,f,1], site_cov[N,f,f],
 *     filt_mean[N,d,1], filt_cov[N,d,d]).  Outputs and workspace are caller-allocated.
 *   - `?` in a comment marks a nullable pointer.
 *   - Calls enqueue work on `stream` and return immediately; the library keeps no state between
 *     calls.  The workspace passed to a call must stay alive until the stream has drained.
 *   - Return value: 0 on success, negative kjx_status otherwise.  Numerical failure is NOT an
 *     error: like the reference (utils.py:25-38 — a failed Cholesky yields NaN, no exception) NaNs
 *     propagate into the outputs.
 *   - Hyperparameters are the already-transformed positive values (after the reference's
 *     softplus, utils.py:41-69); func_dim f = 1 (scalar sites), or f = 2 for the heteroscedastic-noise model.
 *   - `_f64` entry points compute in IEEE double, `_f32` in float (same signatures with float data).
 */
#ifndef KJX_H_
#define KJX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* kjx_stream_t; /* == cudaStream_t */

enum kjx_status {
  KJX_OK = 0,
  KJX_ERR_INVALID_ARG = -1,   /* null pointer, N < 0, bad enum, misaligned pointer ...          */
  KJX_ERR_UNSUPPORTED = -2,   /* valid request this build has no kernel for (e.g. f > 2)        */
  KJX_ERR_WORKSPACE = -3,     /* ws_bytes smaller than kjx_workspace_bytes()                    */
  KJX_ERR_CUDA = -4           /* a CUDA runtime call / kernel launch failed                     */
};

/* one diagonal block of the (always block-diagonal) transition matrix A(dt) = expm(F dt) */
enum kjx_block_kind {
  KJX_BLOCK_MATERN32 = 1,      /* 2x2, priors.py:163-175   : lam = sqrt(3)/ell                    */
  KJX_BLOCK_MATERN52 = 2,      /* 3x3, priors.py:238-255   : lam = sqrt(5)/ell                    */
  KJX_BLOCK_QP_MATERN32 = 3,   /* 4x4 harmonic block of QuasiPeriodicMatern32, priors.py:1076-1084:
                                  exp(-dt lam) [[(1+dt lam)R, dt R],[-dt lam^2 R,(1-dt lam)R]],
                                  R = rotation(dt*omega), utils.py:253-265                        */
  KJX_BLOCK_MATERN12 = 4,      /* 1x1, priors.py:94-104 (Exponential / Matern12): exp(-dt lam), lam = 1/ell */
  KJX_BLOCK_ROTATION = 5,      /* 2x2, exp(-dt lam) R(dt omega): Cosine (priors.py:397-408, lam = 0), a harmonic
                                  of Periodic (:791-820, lam = 0) or of QuasiPeriodicMatern12 (:910-939, lam = 1/ell) */
  KJX_BLOCK_MATERN72 = 6,      /* 4x4, priors.py:326-350   : lam = sqrt(7)/ell                    */
  KJX_BLOCK_SUBBAND_MATERN52 = 7 /* 6x6, SubbandMatern52 (priors.py:605-721): (3x3 Matern-5/2 polynomial block) (x)
                                  R(dt omega), lam = sqrt(5)/ell; states ordered (f, f', f'') x (cos, sin) pairs */
};

typedef struct kjx_block {
  int32_t kind;    /* kjx_block_kind                                                              */
  int32_t count;   /* this block repeated `count` times (spatio-temporal: M inducing points)      */
  double lam;      /* decay rate                                                                  */
  double omega;    /* angular frequency of the harmonic (KJX_BLOCK_QP_MATERN32, KJX_BLOCK_ROTATION) */
} kjx_block_t;

enum kjx_likelihood {
  KJX_LIK_GAUSSIAN = 1,          /* likelihoods.py:331-404; hyp = noise variance                  */
  KJX_LIK_BERNOULLI_PROBIT = 2,  /* likelihoods.py:407-518, link 'probit' (with its 1e-10 jitter) */
  KJX_LIK_BERNOULLI_LOGIT = 3,   /* likelihoods.py:407-518, link 'logit'                          */
  KJX_LIK_POISSON_EXP = 4,       /* likelihoods.py:551-644, link 'exp'                            */
  KJX_LIK_POISSON_LOGISTIC = 5,  /* likelihoods.py:551-644, link 'logistic' (softplus)            */
  KJX_LIK_HETERO_SOFTPLUS = 6,   /* HeteroscedasticNoise, likelihoods.py:647-850, link 'softplus'; func_dim = 2 */
  KJX_LIK_HETERO_EXP = 7         /* HeteroscedasticNoise, link 'exp'; func_dim = 2                 */
};

enum kjx_method {
  KJX_INF_EP = 1,    /* ExpectationPropagation.update      approximate_inference.py:50-73    */
  KJX_INF_EEP = 2,   /* ExtendedEP.update (EKS: power 0)   approximate_inference.py:99-133   */
  KJX_INF_SLEP = 3,  /* StatisticallyLinearisedEP.update   approximate_inference.py:180-210  */
  KJX_INF_VI = 4,    /* VariationalInference.update        approximate_inference.py:302-323  */
  KJX_INF_MOMENT_MATCH = 5 /* no update rule: Likelihood.moment_match(y, m, v, hyp, power) itself, likelihoods.py:122-185
                        (kjx_site_update only; `power` raises the likelihood and scales the site variance) */
};

#define KJX_MAX_BLOCKS 16
#define KJX_MAX_CUBATURE 32

typedef struct kjx_model {
  /* ---- prior (priors.py): block-diagonal A(dt), stationary covariance, measurement model ---- */
  int32_t state_dim;               /* d = sum over blocks of count * block size                   */
  int32_t func_dim;                /* f, rows of H: 1, or 2 for an Independent prior (priors.py:1417-1487) with a
                                      HeteroscedasticNoise likelihood and any of the four update rules (state_dim <= 8;
                                      sites [N,2,1], [N,2,2]) */
  int32_t n_blocks;
  int32_t reserved0;
  kjx_block_t blocks[KJX_MAX_BLOCKS];
  const double* Pinf;              /* HOST [d,d] row-major stationary covariance (read at call time) */
  const double* H;                 /* HOST [f,d] constant measurement model, or NULL with H_steps    */
  const double* H_steps;           /* DEVICE [N,f,d] per-step measurement model (spatio-temporal prior
                                      with scattered data, priors.py:1148-1161), else NULL           */
  /* ---- likelihood + approximate-inference method ---- */
  int32_t likelihood;              /* kjx_likelihood                                              */
  int32_t method;                  /* kjx_method                                                  */
  double lik_hyp;                  /* Gaussian noise variance (unused otherwise)                  */
  double power;                    /* EP power / fraction                                         */
  double damping;                  /* site damping                                                */
  int32_t n_cub;                   /* number of cubature points, 1..KJX_MAX_CUBATURE              */
  int32_t cub_rule;                /* func_dim 2 with KJX_INF_SLEP only: which TWO-dimensional rule the table stands for
                                      (ApproxInf.cubature_func(2), approximate_inference.py:25-34): 0 = Gauss-Hermite, the
                                      product rule of the 1-D table (utils.py:418-463); 3 / 5 = the symmetric 3rd- /
                                      5th-order rule with the reference's literals (utils.py:484-488, 512-516).  Else 0. */
  double cub_x[KJX_MAX_CUBATURE];  /* unit sigma points  (utils.py:455-463 / 466-511)             */
  double cub_w[KJX_MAX_CUBATURE];  /* weights                                                     */
} kjx_model_t;

/* flags */
#define KJX_FLAG_SEQUENTIAL 1u   /* force the one-thread-chain sequential kernels (reference order) */
#define KJX_FLAG_RETURN_FULL 2u  /* smoother: write the full state posterior [N,d],[N,d,d]          */
#define KJX_FLAG_NO_SITE_UPDATE 4u /* kjx_run: smoother pass does not update the sites              */
#define KJX_FLAG_RESIDENT_SITES 8u /* kjx_run_host: the site parameters are iteration state that lives in
                                    * dev_scratch: read the ones the previous call left there (site_*_host
                                    * may be NULL; if given they are uploaded first) and keep the new ones
                                    * there (new_site_*_host may be NULL = no copy back)                   */
#define KJX_FLAG_RESIDENT_INIT 16u /* kjx_run_host: first call of such a sequence: sites come from the host
                                    * pointers (NULL = no sites yet, sde_gp.py:287), new sites stay resident */

int kjx_version(void);
const char* kjx_status_string(int status);

/* Bytes of device workspace any call below needs for N steps of model `m` (dtype_bytes 8 or 4). */
int kjx_workspace_bytes(const kjx_model_t* m, int64_t N, int dtype_bytes, size_t* bytes);

/* Forward filter.  site_mean/site_cov null => sites are initialised from the predictive
 * (sde_gp.py:287) and, when out_site_* are given, stored.  With a non-Gaussian likelihood that
 * recursion is inherently sequential and runs as one dependent chain; otherwise the
 * parallel-in-time scan is used.  filt_mean/filt_cov null => energy only (store=False). */
int kjx_filter_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y,
                   const uint8_t* mask /*?*/, const double* site_mean /*?*/, const double* site_cov /*?*/,
                   double* filt_mean /*?*/, double* filt_cov /*?*/,
                   double* out_site_mean /*?*/, double* out_site_cov /*?*/,
                   double* nlml /* [1] */, void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);

/* Backward smoother.  When site_mean/site_cov (the previous sites) and y are given the sites are
 * updated per step (sde_gp.py:371-379) into new_site_*.  smooth_* (nullable) receive H m, H P H^T
 * ([N,f],[N,f,f]) or, with KJX_FLAG_RETURN_FULL, the full state ([N,d],[N,d,d]). */
int kjx_smoother_f64(const kjx_model_t* m, int64_t N, const double* dt,
                     const double* filt_mean, const double* filt_cov,
                     const double* y /*?*/, const double* site_mean /*?*/, const double* site_cov /*?*/,
                     double* smooth_mean /*?*/, double* smooth_cov /*?*/,
                     double* new_site_mean /*?*/, double* new_site_cov /*?*/,
                     void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);

/* One SDEGP.run() iteration (sde_gp.py:179-187 without the gradient): filter with the given (or
 * initialised) sites, then smoother + site update, as one fused parallel-in-time scan.  site_* in;
 * new_site_* out; post_mean/post_var = smoothed marginals H m, H P H^T (nullable).  filt_mean/filt_cov
 * are optional: run() only uses the filtered moments internally (sde_gp.py:183-187), so by default they
 * stay in the workspace in a packed private layout; pass pointers to also get them in the reference's
 * dense layout. */
int kjx_run_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y,
                const double* site_mean /*?*/, const double* site_cov /*?*/,
                double* filt_mean /*?*/, double* filt_cov /*?*/,
                double* new_site_mean, double* new_site_cov,
                double* post_mean /*?*/, double* post_var /*?*/,
                double* nlml, void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);

/* Posterior marginals on a merged train/test grid — SDEGP.predict_everywhere (sde_gp.py:85-103) in one fused scan:
 * filter with the given sites where masked steps (mask[n] != 0, the test inputs) are predict-only
 * (sde_gp.py:286, 297-300), then the smoother without site update; post_mean/post_var = H m, H P H^T per step.
 * nlml (nullable) receives the filter energy of the unmasked steps. */
int kjx_posterior_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y, const uint8_t* mask /*?*/,
                      const double* site_mean, const double* site_cov, double* post_mean, double* post_var,
                      double* nlml /*?*/, void* ws, size_t ws_bytes, kjx_stream_t stream);

/* Per-step site update, embarrassingly parallel over N (f = 1):
 * (log_lik, site_mean, site_var) = method.update(likelihood, y, post_mean, post_var, prev site?).
 * With prev null this is the "initialise" branch; log_lik alone (site outputs null) with method EP
 * is the vmapped NLPD term of sde_gp.py:139-142. */
int kjx_site_update_f64(const kjx_model_t* m, int64_t N, const double* y,
                        const double* post_mean, const double* post_var,
                        const double* site_mean_prev /*?*/, const double* site_var_prev /*?*/,
                        double* log_lik /*?*/, double* site_mean /*?*/, double* site_var /*?*/,
                        kjx_stream_t stream);

/* Batched discretisation of the prior SDE: A[n] = expm(F dt[n]) (closed forms) and, when Q is
 * non-null, Q[n] = Pinf - A[n] Pinf A[n]^T. */
int kjx_discretise_f64(const kjx_model_t* m, int64_t N, const double* dt,
                       double* A /* [N,d,d] */, double* Q /*? [N,d,d] */, kjx_stream_t stream);

/* ---- time-sharded form (one contiguous shard of the time axis per GPU) ---------------------------
 * The fused iteration split at its one exchange point.  Each shard reduces its steps to ONE filter scan
 * element (A, b, C, eta, J); after an all-gather of those fixed-size summaries every rank folds the ones
 * of the shards before it into the filtered state entering its shard, and the ones after it into the
 * information (eta, J) about its last state — which is all the smoother needs (two-filter identity), so
 * there is a single collective per iteration.  The three calls of a shard share `ws` (it carries the
 * per-chunk prefixes and the filtered moments between them). */
/* packed size of a summary: A[d,d], b[d], upper(C), eta[d], upper(J) -> 2 d^2 + 3 d (27 at d = 3) */
size_t kjx_filter_summary_doubles(int32_t state_dim);
/* phases 1+2: fold + scan this shard; summary_out (device) receives the shard's total element */
int kjx_filter_summary_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y,
                           const double* site_mean /*?*/, const double* site_cov /*?*/,
                           double* summary_out, void* ws, size_t ws_bytes, kjx_stream_t stream);
/* the fold (one tiny kernel): summaries = gathered [world] rows of `stride` doubles (>= packed size, so
 * extra per-rank scalars can ride in the same all-gather); state_out = (mean[d], cov[d,d]) entering
 * shard `rank`; info_out = (eta[d], J[d,d]) of all later shards. */
int kjx_fold_shard_summaries_f64(const kjx_model_t* m, const double* summaries, int64_t stride, int32_t world,
                                 int32_t rank, double* state_out, double* info_out, kjx_stream_t stream);
int kjx_fold_shard_summaries_host(const kjx_model_t* m, const double* summaries_host, int64_t stride, int32_t world,
                                  int32_t rank, double* state_out_host, double* info_out_host);
/* The same exchange + fold fused into ONE kernel over NVLink / NVSwitch peer memory (no NCCL call): every
 * rank's kernel stores its element straight into all peers' receive buffers, publishes a flag with system
 * scope, waits for the world's flags in its own buffer and folds.  peer_bufs: DEVICE array [world] of pointers
 * to each rank's receive buffer (symmetric memory mapped in every process, zero-initialised before first use):
 *   doubles [2][world][stride] (two parities) followed by 64-bit flags [2][world].
 * epoch: DEVICE counter owned by this rank (starts at 0); incremented by the kernel, so the call can be
 * captured in a CUDA graph and replayed.  All `world` ranks must make the call the same number of times. */
int kjx_exchange_fold_f64(const kjx_model_t* m, const double* my_summary, double* const* peer_bufs, int64_t stride,
                          int32_t world, int32_t rank, uint64_t* epoch, double* state_out, double* info_out,
                          kjx_stream_t stream);
/* ONE call per iteration and shard with the exchange INSIDE the scan / filter / smoother kernels (no collective call, no
 * separate exchange launch; 5 kernels): the block that completes the shard's scan stores the shard's element into every
 * rank's receive buffer over NVLink / NVSwitch and releases an epoch flag; the filter kernel's blocks wait for the
 * elements of the earlier shards and fold them into the state entering the shard, the smoother kernel's blocks fold the
 * later shards' elements into the information about the shard's last state; one extra block of the smoother kernel
 * publishes the shard's partial nlml and sums all shards' partials in rank order into nlml[0] (every rank gets the same
 * bits) while the other blocks work; the block that finishes last closes the epoch.  peer_bufs: DEVICE array [world] of receive-buffer pointers (symmetric memory;
 * kjx_shard_link_doubles(world, stride) doubles each, ZERO before first use): summaries [2][world][stride] | flags
 * [2][world] | nlml [2][world] | nlml flags [2][world]; my_buf: this rank's own buffer, by value.  ctrl: DEVICE uint64[4]
 * owned by this rank, zero before first use (epoch and tickets).  world <= 32.  Graph-capturable; all ranks must make the call the same number of times.  world == 1
 * works with an ordinary device buffer.  Same outputs as kjx_filter_summary / exchange / filter_apply / smoother_apply. */
size_t kjx_shard_link_doubles(int32_t world, int64_t stride);
int kjx_sharded_iteration_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y,
                              const double* site_mean /*?*/, const double* site_cov /*?*/,
                              double* const* peer_bufs, double* my_buf /* == peer_bufs[rank] */, int64_t stride,
                              int32_t world, int32_t rank, uint64_t* ctrl,
                              double* post_mean /*?*/, double* post_var /*?*/,
                              double* new_site_mean /*?*/, double* new_site_cov /*?*/, double* nlml /* [1] */,
                              void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);
/* phase 3: the reference filter recursion over the shard from state_in; nlml_partial[1] = this shard's
 * share of the negative log marginal likelihood (sum the shards' values) */
int kjx_filter_apply_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y,
                         const double* site_mean /*?*/, const double* site_cov /*?*/,
                         const double* state_in, double* nlml_partial,
                         void* ws, size_t ws_bytes, kjx_stream_t stream);
/* phase 4: RTS recursion backwards + site update; info_in from the fold (zeros for the last shard) */
int kjx_smoother_apply_f64(const kjx_model_t* m, int64_t N, const double* dt, const double* y,
                           const double* site_mean /*?*/, const double* site_cov /*?*/, const double* info_in,
                           double* post_mean /*?*/, double* post_var /*?*/,
                           double* new_site_mean /*?*/, double* new_site_cov /*?*/,
                           void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);

/* ---- prior samples (SDEGP.prior_sample, sde_gp.py:386-417): num_samps independent draws of f = H x along the
 * series, x_0 = chol(Pinf) eps, x_k = A(dt_k) x_{k-1} + chol(Q(dt_k) + 1e-6 I) eps.  f_sample: DEVICE [N, num_samps]
 * (the reference's [N, func_dim = 1, num_samps]).  Counter-based Philox keyed by `seed`: the same seed gives the
 * same draws; parity with the reference's PRNGKey(i * k + k) stream is in distribution only.  Block-diagonal priors
 * with a constant measurement row and state_dim <= 64 (the reference has no spatio-temporal sampling either). -- */
int kjx_prior_sample_workspace_bytes(const kjx_model_t* m, int64_t N, int32_t num_samps, size_t* bytes);
int kjx_prior_sample_f64(const kjx_model_t* m, int64_t N, const double* dt, int32_t num_samps, uint64_t seed,
                         double* f_sample, void* ws, size_t ws_bytes, kjx_stream_t stream);

/* ---- host-buffer entry point (end-to-end path): one run() iteration where every data array is a
 * HOST pointer (pinned memory for full PCIe speed).  Inputs are copied to the device, the fused
 * filter/smoother/site-update runs, new sites (+ optional posterior marginals) and nlml are copied back;
 * the call returns after the stream has drained.  dev_scratch: caller-owned device memory of at least
 * kjx_run_host_scratch_bytes() bytes (holds the staged arrays, the filter output and the workspace).
 * A training loop (sde_gp.py:179 called once per optimiser step) only needs nlml back: with
 * KJX_FLAG_RESIDENT_INIT on the first call and KJX_FLAG_RESIDENT_SITES afterwards the site parameters — the
 * state the iteration updates, self.sites.site_params in the reference — stay in dev_scratch between calls and
 * only dt, y go up and nlml comes down; pass new_site_*_host on the call whose sites you want to read. -- */
int kjx_run_host_scratch_bytes(const kjx_model_t* m, int64_t N, int dtype_bytes, size_t* bytes);
int kjx_run_host_f64(const kjx_model_t* m, int64_t N, const double* dt_host, const double* y_host,
                     const double* site_mean_host /*?*/, const double* site_cov_host /*?*/,
                     double* new_site_mean_host, double* new_site_cov_host,
                     double* post_mean_host /*?*/, double* post_var_host /*?*/, double* nlml_host,
                     void* dev_scratch, size_t dev_scratch_bytes, uint32_t flags, kjx_stream_t stream);

/* The same iteration WITHOUT the final synchronisation, for loops that overlap the next step's upload with this step's
 * kernels (the PCIe copy of dt, y is 4.8 ms of a 5.9 ms step at N = 2^24): the inputs go host -> device on `copy_stream`
 * into one of two input slots of dev_scratch, the iteration (sites resident in dev_scratch: the sequence must have been
 * started by kjx_run_host_f64 with KJX_FLAG_RESIDENT_INIT) runs on `stream` once the upload has landed, nlml_host is
 * written by a copy enqueued on `stream`.  Returns as soon as everything is enqueued.  The caller alternates slot 0 / 1
 * and must not submit into a slot before the iteration that last used it has completed (e.g. an event recorded on
 * `stream` after the previous submit into that slot), and reads nlml_host after `stream` has passed this call. */
int kjx_run_host_submit_f64(const kjx_model_t* m, int64_t N, const double* dt_host, const double* y_host,
                            double* nlml_host, void* dev_scratch, size_t dev_scratch_bytes, int32_t slot,
                            kjx_stream_t copy_stream, kjx_stream_t stream);

/* float variants: identical semantics, float data (Pinf/H in the model stay double). */
int kjx_filter_f32(const kjx_model_t* m, int64_t N, const float* dt, const float* y,
                   const uint8_t* mask, const float* site_mean, const float* site_cov,
                   float* filt_mean, float* filt_cov, float* out_site_mean, float* out_site_cov,
                   float* nlml, void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);
int kjx_smoother_f32(const kjx_model_t* m, int64_t N, const float* dt,
                     const float* filt_mean, const float* filt_cov,
                     const float* y, const float* site_mean, const float* site_cov,
                     float* smooth_mean, float* smooth_cov, float* new_site_mean, float* new_site_cov,
                     void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);
int kjx_run_f32(const kjx_model_t* m, int64_t N, const float* dt, const float* y,
                const float* site_mean, const float* site_cov, float* filt_mean, float* filt_cov,
                float* new_site_mean, float* new_site_cov, float* post_mean, float* post_var,
                float* nlml, void* ws, size_t ws_bytes, uint32_t flags, kjx_stream_t stream);
int kjx_site_update_f32(const kjx_model_t* m, int64_t N, const float* y, const float* post_mean,
                        const float* post_var, const float* site_mean_prev, const float* site_var_prev,
                        float* log_lik, float* site_mean, float* site_var, kjx_stream_t stream);
int kjx_discretise_f32(const kjx_model_t* m, int64_t N, const float* dt, float* A, float* Q,
                       kjx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KJX_H_ */
