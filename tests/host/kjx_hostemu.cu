// kjx_hostemu.cu — TEST INFRASTRUCTURE.  Compiles the product's __host__ __device__ math headers for
// the CPU and replays the kernels' three-phase scan structure with plain loops, so the element / fold /
// combine / apply algebra and the per-step site updates can be checked against the NumPy oracle in the
// GPU-less build container (tests/test_host_emulation.py).  Nothing here is shipped or timed.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../kalman_jax_b200/csrc/kjx_small_math.cuh"

using namespace kjx;

template <int KIND>
static void emu_run(double lam, const double* Pinf_, int64_t N, int L, const double* dt, const double* ytilde,
                    const double* R, const uint8_t* mask, double* fm, double* fP, double* sm, double* sP, int two_filter) {
  constexpr int D = BlockDim<KIND>::D;
  double Pinf[D][D];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) Pinf[i][j] = Pinf_[i * D + j];
  const int64_t nch = (N + L - 1) / L;
  // F1: per-chunk summaries
  std::vector<FSum<double, D>> sums(nch), prefix(nch);
  for (int64_t c = 0; c < nch; ++c) {
    fsum_identity<double, D>(sums[c]);
    for (int64_t k = c * L; k < std::min<int64_t>(N, (c + 1) * L); ++k) {
      double F[D][D]; Transition<double, KIND>::eval(dt[k], lam, F);
      fsum_fold_step<double, D>(sums[c], F, Pinf, ytilde[k], R[k], mask && mask[k]);
    }
  }
  // F2: exclusive prefix by a (tree-ordered) scan: emulate pairwise combination order of a Kogge-Stone scan
  {
    std::vector<FSum<double, D>> inc = sums;
    for (int64_t off = 1; off < nch; off <<= 1) {
      std::vector<FSum<double, D>> nxt = inc;
      for (int64_t c = off; c < nch; ++c) fsum_combine<double, D>(inc[c - off], inc[c], nxt[c]);
      inc.swap(nxt);
    }
    for (int64_t c = 0; c < nch; ++c) { if (c == 0) fsum_identity<double, D>(prefix[c]); else prefix[c] = inc[c - 1]; }
  }
  // F3: apply + reference recursion, folding smoother elements on the way
  std::vector<SSum<double, D>> ssums(nch);
  for (int64_t c = 0; c < nch; ++c) {
    FiltState<double, D> x;
    for (int i = 0; i < D; ++i) { x.m[i] = 0; for (int j = 0; j < D; ++j) x.P[i][j] = Pinf[i][j]; }
    fsum_apply<double, D>(prefix[c], x);
    ssum_identity<double, D>(ssums[c]);
    const int64_t s0 = c * L, s1 = std::min<int64_t>(N, (c + 1) * L);
    double F[D][D]; Transition<double, KIND>::eval(dt[s0], lam, F);
    filter_predict<double, D>(F, Pinf, x);
    for (int64_t k = s0; k < s1; ++k) {
      if (!(mask && mask[k])) filter_update<double, D>(x, ytilde[k], R[k]);
      for (int i = 0; i < D; ++i) { fm[k * D + i] = x.m[i]; for (int j = 0; j < D; ++j) fP[k * D * D + i * D + j] = x.P[i][j]; }
      SSum<double, D> e, r;
      if (k + 1 < N) {
        Transition<double, KIND>::eval(dt[k + 1], lam, F);
        SmoothGain<double, D> g; smoother_gain<double, D>(F, Pinf, x.m, x.P, g);
        ssum_element<double, D>(g, x.m, x.P, e);
        for (int i = 0; i < D; ++i) { x.m[i] = g.mp[i]; for (int j = 0; j < D; ++j) x.P[i][j] = g.Pp[i][j]; }
      } else {
        ssum_terminal<double, D>(x.m, x.P, e);
      }
      ssum_combine<double, D>(ssums[c], e, r); ssums[c] = r;
    }
  }
  // S2: exclusive suffix
  std::vector<SSum<double, D>> suffix(nch);
  {
    std::vector<SSum<double, D>> inc = ssums;
    for (int64_t off = 1; off < nch; off <<= 1) {
      std::vector<SSum<double, D>> nxt = inc;
      for (int64_t c = 0; c + off < nch; ++c) ssum_combine<double, D>(inc[c], inc[c + off], nxt[c]);
      inc.swap(nxt);
    }
    for (int64_t c = 0; c < nch; ++c) { if (c == nch - 1) ssum_identity<double, D>(suffix[c]); else suffix[c] = inc[c + 1]; }
  }
  // two-filter variant: boundary states from a suffix scan over the FILTER elements
  std::vector<Info<double, D>> finfo(nch);
  {
    std::vector<FSum<double, D>> inc = sums;
    for (int64_t off = 1; off < nch; off <<= 1) {
      std::vector<FSum<double, D>> nxt = inc;
      for (int64_t c = 0; c + off < nch; ++c) fsum_combine<double, D>(inc[c], inc[c + off], nxt[c]);
      inc.swap(nxt);
    }
    for (int64_t c = 0; c < nch; ++c) { if (c == nch - 1) info_zero<double, D>(finfo[c]); else info_of<double, D>(inc[c + 1], finfo[c]); }
  }
  // S3: apply + reference RTS recursion
  for (int64_t c = nch - 1; c >= 0; --c) {
    const int64_t s0 = c * L, s1 = std::min<int64_t>(N, (c + 1) * L);
    FiltState<double, D> x;
    if (two_filter) {
      // smoothed state of the chunk's LAST step directly, then RTS for the remaining steps
      const int64_t k = s1 - 1;
      for (int i = 0; i < D; ++i) { x.m[i] = fm[k * D + i]; for (int j = 0; j < D; ++j) x.P[i][j] = fP[k * D * D + i * D + j]; }
      info_smooth<double, D>(finfo[c], x);
      for (int i = 0; i < D; ++i) { sm[k * D + i] = x.m[i]; for (int j = 0; j < D; ++j) sP[k * D * D + i * D + j] = x.P[i][j]; }
      for (int64_t kk = s1 - 2; kk >= s0; --kk) {
        double F[D][D]; Transition<double, KIND>::eval(dt[kk + 1], lam, F);
        double mf[D], Pf[D][D];
        for (int i = 0; i < D; ++i) { mf[i] = fm[kk * D + i]; for (int j = 0; j < D; ++j) Pf[i][j] = fP[kk * D * D + i * D + j]; }
        SmoothGain<double, D> g; smoother_gain<double, D>(F, Pinf, mf, Pf, g);
        smoother_step<double, D>(g, mf, Pf, x);
        for (int i = 0; i < D; ++i) { sm[kk * D + i] = x.m[i]; for (int j = 0; j < D; ++j) sP[kk * D * D + i * D + j] = x.P[i][j]; }
      }
      continue;
    }
    if (s1 == N) {
      for (int i = 0; i < D; ++i) { x.m[i] = fm[(N - 1) * D + i]; for (int j = 0; j < D; ++j) x.P[i][j] = fP[(N - 1) * D * D + i * D + j]; }
    } else {
      for (int i = 0; i < D; ++i) { x.m[i] = 0; for (int j = 0; j < D; ++j) x.P[i][j] = 0; }
      ssum_apply<double, D>(suffix[c], x);
    }
    for (int64_t k = s1 - 1; k >= s0; --k) {
      double F[D][D]; Transition<double, KIND>::eval((k + 1 < N) ? dt[k + 1] : 0.0, lam, F);
      double mf[D], Pf[D][D];
      for (int i = 0; i < D; ++i) { mf[i] = fm[k * D + i]; for (int j = 0; j < D; ++j) Pf[i][j] = fP[k * D * D + i * D + j]; }
      SmoothGain<double, D> g; smoother_gain<double, D>(F, Pinf, mf, Pf, g);
      smoother_step<double, D>(g, mf, Pf, x);
      for (int i = 0; i < D; ++i) { sm[k * D + i] = x.m[i]; for (int j = 0; j < D; ++j) sP[k * D * D + i * D + j] = x.P[i][j]; }
    }
  }
}

template <int KIND>
static void emu_shard_summary_t(double lam, const double* Pinf_, int64_t N, const double* dt, const double* yt, const double* R, double* out) {
  constexpr int D = BlockDim<KIND>::D;
  double Pinf[D][D];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) Pinf[i][j] = Pinf_[i * D + j];
  FSum<double, D> s; fsum_identity<double, D>(s);
  for (int64_t k = 0; k < N; ++k) {
    double F[D][D]; Transition<double, KIND>::eval(dt[k], lam, F);
    fsum_fold_step<double, D>(s, F, Pinf, yt[k], R[k], false);
  }
  fsum_pack<double, D>(s, out);
}

template <int KIND>
static void emu_shard_apply_t(double lam, const double* Pinf_, int64_t N, const double* dt, const double* yt, const double* R,
                              const double* state_in, const double* info_in, double* fm, double* fP, double* sm, double* sP) {
  constexpr int D = BlockDim<KIND>::D;
  double Pinf[D][D];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) Pinf[i][j] = Pinf_[i * D + j];
  FiltState<double, D> x;
  for (int i = 0; i < D; ++i) { x.m[i] = state_in[i]; for (int j = 0; j < D; ++j) x.P[i][j] = state_in[D + i * D + j]; }
  for (int64_t k = 0; k < N; ++k) {
    double F[D][D]; Transition<double, KIND>::eval(dt[k], lam, F);
    filter_predict<double, D>(F, Pinf, x);
    filter_update<double, D>(x, yt[k], R[k]);
    for (int i = 0; i < D; ++i) { fm[k * D + i] = x.m[i]; for (int j = 0; j < D; ++j) fP[k * D * D + i * D + j] = x.P[i][j]; }
  }
  Info<double, D> f;
  for (int i = 0; i < D; ++i) { f.eta[i] = info_in[i]; for (int j = 0; j < D; ++j) f.J[i][j] = info_in[D + i * D + j]; }
  info_smooth<double, D>(f, x);                    // smoothed state of the shard's last step (two-filter identity)
  for (int64_t k = N - 1; k >= 0; --k) {
    if (k < N - 1) {
      double F[D][D]; Transition<double, KIND>::eval(dt[k + 1], lam, F);
      double mf[D], Pf[D][D];
      for (int i = 0; i < D; ++i) { mf[i] = fm[k * D + i]; for (int j = 0; j < D; ++j) Pf[i][j] = fP[k * D * D + i * D + j]; }
      SmoothGain<double, D> g; smoother_gain<double, D>(F, Pinf, mf, Pf, g);
      smoother_step<double, D>(g, mf, Pf, x);
    }
    for (int i = 0; i < D; ++i) { sm[k * D + i] = x.m[i]; for (int j = 0; j < D; ++j) sP[k * D * D + i * D + j] = x.P[i][j]; }
  }
}

extern "C" {

// one shard of the time-sharded iteration, emulated on the host with the product's math
int emu_shard_summary(int kind, double lam, const double* Pinf, int64_t N, const double* dt, const double* yt, const double* R, double* out) {
  if (kind == KJX_BLOCK_MATERN32) emu_shard_summary_t<KJX_BLOCK_MATERN32>(lam, Pinf, N, dt, yt, R, out);
  else if (kind == KJX_BLOCK_MATERN52) emu_shard_summary_t<KJX_BLOCK_MATERN52>(lam, Pinf, N, dt, yt, R, out);
  else return -1;
  return 0;
}
int emu_shard_apply(int kind, double lam, const double* Pinf, int64_t N, const double* dt, const double* yt, const double* R,
                    const double* state_in, const double* info_in, double* fm, double* fP, double* sm, double* sP) {
  if (kind == KJX_BLOCK_MATERN32) emu_shard_apply_t<KJX_BLOCK_MATERN32>(lam, Pinf, N, dt, yt, R, state_in, info_in, fm, fP, sm, sP);
  else if (kind == KJX_BLOCK_MATERN52) emu_shard_apply_t<KJX_BLOCK_MATERN52>(lam, Pinf, N, dt, yt, R, state_in, info_in, fm, fP, sm, sP);
  else return -1;
  return 0;
}

// three-phase scan emulation of filter + smoother for given pseudo-observations (ytilde, R)
int emu_filter_smoother(int kind, double lam, const double* Pinf, int64_t N, int L, const double* dt,
                        const double* ytilde, const double* R, const uint8_t* mask, double* fm, double* fP,
                        double* sm, double* sP, int two_filter) {
  if (kind == KJX_BLOCK_MATERN32) emu_run<KJX_BLOCK_MATERN32>(lam, Pinf, N, L, dt, ytilde, R, mask, fm, fP, sm, sP, two_filter);
  else if (kind == KJX_BLOCK_MATERN52) emu_run<KJX_BLOCK_MATERN52>(lam, Pinf, N, L, dt, ytilde, R, mask, fm, fP, sm, sP, two_filter);
  else return -1;
  return 0;
}

// per-step site updates with the product's device math, run on the host
int emu_site_update(const kjx_model_t* m, int64_t N, const double* y, const double* pm, const double* pv,
                    const double* sprev_mean, const double* sprev_var, double* lml, double* smean, double* svar) {
  SiteModel s; memset(&s, 0, sizeof(s));
  s.lik = m->likelihood; s.method = m->method; s.n_cub = m->n_cub; s.lik_hyp = m->lik_hyp; s.power = m->power; s.damping = m->damping;
  for (int i = 0; i < KJX_MAX_CUBATURE; ++i) { s.cub_x[i] = m->cub_x[i]; s.cub_w[i] = m->cub_w[i]; }
  for (int64_t k = 0; k < N; ++k) {
    SiteOut<double> o = site_update<double>(s, y[k], pm[k], pv[k], sprev_mean != nullptr,
                                            sprev_mean ? sprev_mean[k] : 0.0, sprev_var ? sprev_var[k] : 0.0);
    lml[k] = o.lml; smean[k] = o.mean; svar[k] = o.var;
  }
  return 0;
}

}
