// kjx_sites.cuh — per-step site-update moments for scalar sites (f = 1), device side.
//
// Device restatement of the reference's approximate-inference rules and likelihood primitives
// (paths relative to /root/reference/kalmanjax):
//   approximate_inference.py:8-15    compute_cavity
//   approximate_inference.py:50-73   ExpectationPropagation.update   (EP / PEP)
//   approximate_inference.py:99-133  ExtendedEP.update               (EEP / EKS / EKF)
//   approximate_inference.py:180-210 StatisticallyLinearisedEP.update(SLEP / GHEP / UEP / PL / *KS)
//   approximate_inference.py:302-323 VariationalInference.update     (VI / CVI)
//   likelihoods.py:122-185, 203-242, 275-321   cubature moment match / SLR / variational expectation
//   likelihoods.py:331-404 Gaussian, 407-518 Bernoulli(probit|logit), 551-644 Poisson(exp|logistic)
//   utils.py:15-38, 200-244          ensure_positive_precision, inv, logphi, gaussian_moment_match
// The numerical-safety constants (1e-10 jitters with their signs, 1e-8 clamp, 0.01 precision
// floor) and the NaN behaviour of a failed 1x1 Cholesky are reproduced on purpose.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/kjx.h"

namespace kjx {

// Likelihood + inference-method description, passed to kernels by value (constant bank).
struct SiteModel {
  int lik;       // kjx_likelihood
  int method;    // kjx_method
  int n_cub;
  int cub_rule;  // multi-dimensional cubature (func_dim 2, SLR family): 0 = Gauss-Hermite product of the 1-D table,
                 // 3 / 5 = the reference's 3rd- / 5th-order symmetric rule (utils.py:466-516)
  double lik_hyp, power, damping;
  double cub_x[KJX_MAX_CUBATURE];
  double cub_w[KJX_MAX_CUBATURE];
};

template <typename T> struct SiteOut { T lml, mean, var; };

#define KJX_HD __host__ __device__ __forceinline__
#define KJX_HDN __host__ __device__

template <typename T> KJX_HD T qnan() {
#ifdef __CUDA_ARCH__
  return T(__longlong_as_double(0x7ff8000000000000LL));
#else
  return T(NAN);
#endif
}

// np.maximum(x, lo): NaN in x propagates (fmax would drop it)
template <typename T> KJX_HD T max_nan(T x, T lo) { return (x < lo) ? lo : x; }

// utils.py:33-38 on a 1x1 matrix: Cholesky inverse, NaN when the entry is negative
template <typename T> KJX_HD T inv1(T v) { return (v < T(0)) ? qnan<T>() : T(1) / v; }
// cho_factor of a 1x1 matrix
template <typename T> KJX_HD T chol1(T v) { return sqrt(v); }
// utils.py:15-22 on a 1x1 matrix
template <typename T> KJX_HD T ensure_pos1(T k) { return (k < T(0)) ? T(1e-2) : k; }

#define KJX_PI 3.141592653589793

// How the cubature sums are evaluated: by one thread (SerialRed), or spread over the 32 lanes of a warp whose lanes
// all hold the same arguments (WarpRed: lane i evaluates sigma points i, i+32, ...; butterfly reduction, so every
// lane ends up with the same total).  The one-chain kernels use WarpRed: the chain is sequential in time, but the
// 20 Gauss-Hermite evaluations of a step are not.
struct SerialRed {
  static KJX_HD int first() { return 0; }
  static KJX_HD int stride() { return 1; }
  template <typename T> static KJX_HD T sum(T v) { return v; }
};
struct WarpRed {
  static KJX_HD int first() {
#ifdef __CUDA_ARCH__
    return threadIdx.x & 31;
#else
    return 0;
#endif
  }
  static KJX_HD int stride() {
#ifdef __CUDA_ARCH__
    return 32;
#else
    return 1;
#endif
  }
  template <typename T> static KJX_HD T sum(T v) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
    return v;
  }
};

// ------------------------------------------------------------------------------------------------
// link functions (likelihoods.py:421-433, 571-576; utils.py:67-73)
template <typename T> KJX_HD T link_fn(int lik, T f) {
  switch (lik) {
    case KJX_LIK_BERNOULLI_LOGIT: return T(1) / (T(1) + exp(-f));
    case KJX_LIK_BERNOULLI_PROBIT:
      return T(0.5) * (T(1) + erf(f / T(1.4142135623730951))) * T(1.0 - 2e-10) + T(1e-10);
    case KJX_LIK_POISSON_EXP: return exp(f);
    case KJX_LIK_POISSON_LOGISTIC: return log(T(1) + exp(-fabs(f))) + ((f < T(0)) ? T(0) : f);
    default: return f;  // Gaussian: identity
  }
}
template <typename T> KJX_HD T dlink_fn(int lik, T f) {
  switch (lik) {
    case KJX_LIK_BERNOULLI_LOGIT: { T e = exp(f); return e / ((T(1) + e) * (T(1) + e)); }
    case KJX_LIK_BERNOULLI_PROBIT:
      return T(1.0 - 2e-10) * exp(T(-0.5) * f * f) / T(2.5066282746310002);  // grad of :428-429
    case KJX_LIK_POISSON_EXP: return exp(f);
    case KJX_LIK_POISSON_LOGISTIC: { T e = exp(f); return e / (e + T(1)); }
    default: return T(1);
  }
}

// Per-step constants hoisted out of the cubature loop.
template <typename T> struct LikStep {
  T y;          // observation as the likelihood sees it
  T c0;         // Gaussian: (2 pi hyp)^-1/2 ; Poisson: exp(gammaln(y+1))
  T c1;         // Gaussian: -0.5 log(2 pi hyp); Poisson: gammaln(y+1)
  T hyp;
};
template <typename T> KJX_HD LikStep<T> lik_step(const SiteModel& sm, T y) {
  LikStep<T> s; s.y = y; s.hyp = T(sm.lik_hyp); s.c0 = T(0); s.c1 = T(0);
  if (sm.lik == KJX_LIK_GAUSSIAN) {
    s.c0 = pow(T(2 * KJX_PI) * s.hyp, T(-0.5));            // likelihoods.py:357
    s.c1 = T(-0.5) * log(T(2 * KJX_PI) * s.hyp);           // likelihoods.py:371
  } else if (sm.lik == KJX_LIK_POISSON_EXP || sm.lik == KJX_LIK_POISSON_LOGISTIC) {
    s.c1 = lgamma(y + T(1));
    s.c0 = exp(s.c1);
  }
  return s;
}

// p(y | f)   (likelihoods.py:357, 445, 595-596)
template <typename T> KJX_HD T eval_lik(const SiteModel& sm, const LikStep<T>& s, T f) {
  switch (sm.lik) {
    case KJX_LIK_GAUSSIAN: { T r = s.y - f; return s.c0 * exp(T(-0.5) * r * r / s.hyp); }
    case KJX_LIK_BERNOULLI_PROBIT:
    case KJX_LIK_BERNOULLI_LOGIT: { T p = link_fn<T>(sm.lik, f); return (s.y == T(1)) ? p : T(1) - p; }
    default: { T mu = link_fn<T>(sm.lik, f); return pow(mu, s.y) * exp(-mu) / s.c0; }
  }
}
// log p(y | f)   (likelihoods.py:371, 456, 612-613)
template <typename T> KJX_HD T eval_loglik(const SiteModel& sm, const LikStep<T>& s, T f) {
  switch (sm.lik) {
    case KJX_LIK_GAUSSIAN: { T r = s.y - f; return s.c1 - T(0.5) * r * r / s.hyp; }
    case KJX_LIK_BERNOULLI_PROBIT:
    case KJX_LIK_BERNOULLI_LOGIT: return log(eval_lik<T>(sm, s, f));
    default: { T mu = link_fn<T>(sm.lik, f); return s.y * log(mu) - mu - s.c1; }
  }
}
// E[y|f], Var[y|f]   (likelihoods.py:380-381, 465, 631)
template <typename T> KJX_HD void cond_moments(const SiteModel& sm, T f, T& E, T& V) {
  switch (sm.lik) {
    case KJX_LIK_GAUSSIAN: E = f; V = T(sm.lik_hyp); break;
    case KJX_LIK_BERNOULLI_PROBIT:
    case KJX_LIK_BERNOULLI_LOGIT: { T p = link_fn<T>(sm.lik, f); E = p; V = p - p * p; break; }
    default: { T mu = link_fn<T>(sm.lik, f); E = mu; V = mu; }
  }
}
// Jacobians of h(f, s) = E[y|f] + sqrt(Var[y|f]) s at (m, 0)  (likelihoods.py:507-518, 636-644;
// Gaussian: exact derivative of the base-class observation model, likelihoods.py:250-273)
template <typename T> KJX_HD void lin_jacobians(const SiteModel& sm, T m, T& Jf, T& Jsigma) {
  switch (sm.lik) {
    case KJX_LIK_GAUSSIAN: Jf = T(1); Jsigma = sqrt(T(sm.lik_hyp)); break;
    case KJX_LIK_BERNOULLI_PROBIT:
    case KJX_LIK_BERNOULLI_LOGIT: {
      T p = link_fn<T>(sm.lik, m), dp = dlink_fn<T>(sm.lik, m);
      T q = p * (T(1) - p);
      Jf = dp + (T(0.5) * pow(q, T(-0.5)) * dp * (T(1) - T(2) * p)) * T(0);   // sigma = 0 term kept: inf*0 = NaN
      Jsigma = sqrt(q);
      break;
    }
    default: {
      T mu = link_fn<T>(sm.lik, m), dmu = dlink_fn<T>(sm.lik, m);
      Jf = mu + T(0.5) * pow(mu, T(-0.5)) * dmu * T(0);   // link_fn, not dlink_fn: as written at :642
      Jsigma = sqrt(mu);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// utils.py:219-244
template <typename T> KJX_HD SiteOut<T> gaussian_moment_match(T y, T m, T v, T hyp) {
  SiteOut<T> o;
  T r = y - m, s = hyp + v;
  o.lml = -(r * r) / s / T(2) - log(max_nan(T(2 * KJX_PI) * s, T(1e-10))) / T(2);
  o.mean = y; o.var = hyp;
  return o;
}

// utils.py:200-216
template <typename T> KJX_HD void logphi(T z, T& lp, T& dlp) {
  lp = log(erfc(-z / T(1.4142135623730951)) / T(2));
  dlp = exp(-z * z / T(2) - lp) / T(2.5066282746310002);
}

// likelihoods.py:122-185 (the live moment_match_cubature)
template <typename T, class Red = SerialRed>
KJX_HDN SiteOut<T> moment_match_cubature(const SiteModel& sm, T y, T cm, T cv, T power) {
  const LikStep<T> ls = lik_step<T>(sm, y);
  const T cho = chol1(cv);
  const T invC = inv1(cv);
  T Z = T(0), dZ = T(0), d2Z = T(0);
  for (int i = Red::first(); i < sm.n_cub; i += Red::stride()) {
    const T f = cho * T(sm.cub_x[i]) + cm;                       // :147
    T l = eval_lik<T>(sm, ls, f);
    if (power != T(1)) l = pow(l, power);
    const T wl = T(sm.cub_w[i]) * l;                             // :149
    const T d = f - cm;
    Z += wl;                                                     // :154
    dZ += invC * d * wl;                                         // :163-166
    d2Z += (invC * d * d * invC - invC) * wl;                    // :173-176
  }
  Z = Red::sum(Z); dZ = Red::sum(dZ); d2Z = Red::sum(d2Z);
  SiteOut<T> o;
  const T Zc = max_nan(Z, T(1e-8));
  o.lml = log(Zc);                                               // :157
  const T Zinv = T(1) / Zc;                                      // :158
  const T dlZ = Zinv * dZ;
  const T d2lZ = -dlZ * dlZ + Zinv * d2Z;                        // :181
  const T id2lZ = inv1(ensure_pos1(-d2lZ) - T(1e-10));           // :182 (minus jitter as written)
  o.mean = cm + id2lZ * dlZ;                                     // :183
  o.var = power * (-cv + id2lZ);                                 // :184
  return o;
}

// Likelihood.moment_match dispatch: Gaussian closed form (likelihoods.py:385-404), probit closed form
// when power == 1 (likelihoods.py:487-502), cubature otherwise.
template <typename T, class Red = SerialRed>
KJX_HDN SiteOut<T> moment_match(const SiteModel& sm, T y, T m, T v, T power) {
  if (sm.lik == KJX_LIK_GAUSSIAN) return gaussian_moment_match<T>(y, m, v, T(sm.lik_hyp));
  if (sm.lik == KJX_LIK_BERNOULLI_PROBIT || sm.lik == KJX_LIK_BERNOULLI_LOGIT) {
    y = (y > T(0)) ? T(1) : ((y < T(0)) ? T(-1) : y);            // np.sign, :488 (NaN stays NaN)
    if (power == T(1) && sm.lik == KJX_LIK_BERNOULLI_PROBIT) {
      T ys = y - T(0.01);
      ys = (ys > T(0)) ? T(1) : ((ys < T(0)) ? T(-1) : ys);      // :490
      const T sq = sqrt(T(1) + v);
      const T z = (m / sq) * ys;
      T lp, dlp; logphi<T>(z, lp, dlp);
      const T dlZ = ys * dlp / sq;
      const T d2lZ = -dlp * (z + dlp) / (T(1) + v);
      SiteOut<T> o; o.lml = lp; o.mean = m - dlZ / d2lZ; o.var = -(v + T(1) / d2lZ);
      return o;
    }
  }
  return moment_match_cubature<T, Red>(sm, y, m, v, power);
}

// approximate_inference.py:8-15
template <typename T>
KJX_HD void compute_cavity(T pm, T pv, T smean, T svar, T power, T& cm, T& cv) {
  const T post_prec = inv1(pv), site_prec = inv1(svar);
  cv = inv1(post_prec - power * site_prec);
  cm = cv * (post_prec * pm - power * site_prec * smean);
}

// likelihoods.py:203-242
template <typename T, class Red = SerialRed>
KJX_HDN void slr_cubature(const SiteModel& sm, T cm, T cv, T& mu, T& S, T& C, T& omega) {
  const T cho = chol1(cv), invC = inv1(cv);
  mu = T(0);
  for (int i = Red::first(); i < sm.n_cub; i += Red::stride()) {
    T E, V; cond_moments<T>(sm, cho * T(sm.cub_x[i]) + cm, E, V);
    mu += T(sm.cub_w[i]) * E;                                    // :219-221
  }
  mu = Red::sum(mu);
  S = T(0); C = T(0); omega = T(0);
  for (int i = Red::first(); i < sm.n_cub; i += Red::stride()) {
    const T f = cho * T(sm.cub_x[i]) + cm, w = T(sm.cub_w[i]);
    T E, V; cond_moments<T>(sm, f, E, V);
    S += w * ((E - mu) * (E - mu) + V);                          // :226-228
    C += w * (f - cm) * (E - mu);                                // :232-234
    omega += w * E * (invC * (f - cm));                          // :238-240
  }
  S = Red::sum(S); C = Red::sum(C); omega = Red::sum(omega);
}

// likelihoods.py:275-321
template <typename T, class Red = SerialRed>
KJX_HDN void var_exp_cubature(const SiteModel& sm, T y, T pm, T pv, T& dE_dm, T& dE_dv) {
  const LikStep<T> ls = lik_step<T>(sm, y);
  const T cho = chol1(pv), invv = T(1) / pv;                     // :307 plain reciprocal
  dE_dm = T(0); dE_dv = T(0);
  for (int i = Red::first(); i < sm.n_cub; i += Red::stride()) {
    const T f = cho * T(sm.cub_x[i]) + pm;
    const T wll = T(sm.cub_w[i]) * eval_loglik<T>(sm, ls, f);    // :297
    const T d = f - pm;
    dE_dm += invv * d * wll;                                     // :308-311
    dE_dv += (T(0.5) * (invv * invv * d * d) - T(0.5) * invv) * wll;   // :315-318
  }
  dE_dm = Red::sum(dE_dm); dE_dv = Red::sum(dE_dv);
}

// natural-parameter damping shared by the rules (each passes its own jitter)
template <typename T>
KJX_HD void damp(T nat2, T smean, T nat2_prev, T smean_prev, T damping, T jitter,
                                     T& out_mean, T& out_var) {
  const T nat1 = nat2 * smean, nat1_prev = nat2_prev * smean_prev;
  out_var = inv1((T(1) - damping) * nat2_prev + damping * nat2 + jitter);
  out_mean = out_var * ((T(1) - damping) * nat1_prev + damping * nat1);
}

// ------------------------------------------------------------------------------------------------
// ApproxInf.update: has_prev == false is the `site_params is None` branch used by the filter.
template <typename T, class Red = SerialRed>
KJX_HDN SiteOut<T> site_update(const SiteModel& sm, T y, T pm, T pv, bool has_prev, T sprev_mean, T sprev_var) {
  const T damping = T(sm.damping);
  SiteOut<T> o;
  switch (sm.method) {
    case KJX_INF_EP: {                                           // approximate_inference.py:50-73
      if (!has_prev) return moment_match<T, Red>(sm, y, pm, pv, T(1));
      const T power = T(sm.power);
      T cm, cv; compute_cavity<T>(pm, pv, sprev_mean, sprev_var, power, cm, cv);
      o = moment_match<T, Red>(sm, y, cm, cv, power);
      if (damping != T(1))
        damp<T>(inv1(o.var), o.mean, inv1(sprev_var), sprev_mean, damping, T(0), o.mean, o.var);   // :68-72
      return o;
    }
    case KJX_INF_EEP: {                                          // approximate_inference.py:99-133
      const T power = has_prev ? T(sm.power) : T(1);
      T cm = pm, cv = pv;
      if (has_prev && power != T(0)) compute_cavity<T>(pm, pv, sprev_mean, sprev_var, power, cm, cv);
      T Jf, Js; lin_jacobians<T>(sm, cm, Jf, Js);                // :112
      T E, V; cond_moments<T>(sm, cm, E, V);                     // :114
      const T residual = y - E;
      const T sigma = Js * Js + power * Jf * cv * Jf;            // :116
      const T nat2 = Jf * inv1(Js * Js) * Jf;                    // :117
      o.var = inv1(nat2 + T(1e-10));                             // :118
      o.mean = cm + (o.var + power * cv) * Jf * inv1(sigma) * residual;   // :119
      const T sml = Js * Js + Jf * cv * Jf;                      // :121
      const T ch = sqrt(sml);
      o.lml = -(T(0.5) * log(T(2 * KJX_PI)) + log(ch) + T(0.5) * (residual * ((residual / ch) / ch)));   // :123-126
      if (has_prev && damping != T(1))
        damp<T>(nat2, o.mean, inv1(sprev_var + T(1e-10)), sprev_mean, damping, T(1e-10), o.mean, o.var);  // :127-132
      return o;
    }
    case KJX_INF_SLEP: {                                         // approximate_inference.py:180-210
      const T power = has_prev ? T(sm.power) : T(1);
      o.lml = moment_match<T, Red>(sm, y, pm, pv, T(1)).lml;     // :186
      T cm = pm, cv = pv;
      if (has_prev && power != T(0)) compute_cavity<T>(pm, pv, sprev_mean, sprev_var, power, cm, cv);
      T mu, S, C, omega; slr_cubature<T, Red>(sm, cm, cv, mu, S, C, omega);     // :194
      const T residual = y - mu;
      const T sigma = S + (power - T(1)) * C * inv1(cv + T(1e-10)) * C;    // :197
      const T isig = inv1(sigma);
      const T osigo = inv1(omega * isig * omega + T(1e-10));     // :198-200
      o.mean = cm + osigo * omega * isig * residual;             // :201
      o.var = -power * cv + osigo;                               // :202
      if (has_prev && damping != T(1))
        damp<T>(inv1(o.var + T(1e-10)), o.mean, inv1(sprev_var + T(1e-10)), sprev_mean, damping, T(1e-10),
                o.mean, o.var);                                  // :203-209
      return o;
    }
    case KJX_INF_MOMENT_MATCH:                                   // likelihoods.py:122-185 as called by sde_gp.py:139-142
      return moment_match<T, Red>(sm, y, pm, pv, T(sm.power));
    default: {                                                   // VI: approximate_inference.py:302-323
      T dE_dm, dE_dv; var_exp_cubature<T, Red>(sm, y, pm, pv, dE_dm, dE_dv);
      if (!has_prev) {
        o.var = T(0.5) * inv1(ensure_pos1(-dE_dv) + T(1e-10));   // :309
        o.mean = pm + o.var * dE_dm;                             // :310
      } else {
        dE_dv = -ensure_pos1(-dE_dv);                            // :315
        T lam2 = inv1(sprev_var + T(1e-10));                     // :316
        T lam1 = lam2 * sprev_mean;
        lam1 = (T(1) - damping) * lam1 + damping * (dE_dm - T(2) * dE_dv * pm);   // :318
        lam2 = (T(1) - damping) * lam2 + damping * (T(-2) * dE_dv);               // :319
        o.var = inv1(lam2 + T(1e-10));                           // :320
        o.mean = o.var * lam1;
      }
      o.lml = moment_match<T, Red>(sm, y, pm, pv, T(1)).lml;     // :322
      return o;
    }
  }
}

// log-likelihood term of the filter (sde_gp.py:287 with site_params=None): only `lml` of update()
template <typename T, class Red = SerialRed>
KJX_HD T filter_loglik(const SiteModel& sm, T y, T pm, T pv) {
  if (sm.method == KJX_INF_EEP) return site_update<T, Red>(sm, y, pm, pv, false, T(0), T(0)).lml;
  return moment_match<T, Red>(sm, y, pm, pv, T(1)).lml;   // EP :57, SLEP :186, VI :322
}

}  // namespace kjx
