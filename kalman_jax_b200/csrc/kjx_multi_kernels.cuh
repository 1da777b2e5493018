// kjx_multi_kernels.cuh — multi-latent sites (func_dim f = 2): `Independent` priors (priors.py:1417-1487: block-diagonal
// A and Pinf, block-diagonal H with one function value per component) with the `HeteroscedasticNoise` likelihood
// (likelihoods.py:647-728) and (power-)EP (approximate_inference.py:8-15, 50-73 in matrix form: 2x2 Cholesky inverses).
//
// The series these models are used on are short (the motorcycle data: 133 points), so this path is ONE warp walking the
// series in reference order (sde_gp.py:271-306, :349-379) with the 20 cubature points of a step spread over its lanes;
// the state (d <= 8) lives in registers / local memory of every lane (all lanes hold the same values), lane 0 stores.
// A parallel-in-time form needs scan elements with rank-2 updates and is not built.
#pragma once
#include <stdint.h>
#include "kjx_elementwise_kernels.cuh"

namespace kjx {

constexpr int KJX_MULTI_DMAX = 8;

template <typename T> struct MultiArgs {
  DiscArgs disc; SiteModel site; int d; int64_t N;
  double Pinf[KJX_MULTI_DMAX * KJX_MULTI_DMAX]; double H[2 * KJX_MULTI_DMAX];
  const T* dt; const T* y; const uint8_t* mask; const T* site_mean; const T* site_cov;     // sites: [N,2], [N,2,2]
  T* filt_mean; T* filt_cov; T* out_site_mean; T* out_site_cov; T* nlml;
  const T* filt_mean_in; const T* filt_cov_in;
  T* smooth_mean; T* smooth_cov; T* new_site_mean; T* new_site_cov;
  int return_full; int update_sites; int init_sites;
};

// ---- 2 x 2 helpers in the reference's arithmetic (Cholesky-based inverse: NaN when not positive definite) -------------
template <typename T> struct M2 { T a, b, c, d; };          // [[a, b], [c, d]]
template <typename T> __device__ __forceinline__ M2<T> inv2(const M2<T>& P) {      // utils.py:33-38 (upper factor of cho_factor)
  const T u00 = sqrt(P.a), u01 = P.b / u00, u11 = sqrt(P.d - u01 * u01);
  // P^-1 = U^-1 U^-T
  const T i00 = T(1) / u00, i11 = T(1) / u11, i01 = -u01 * i00 * i11;
  M2<T> R; R.a = i00 * i00 + i01 * i01; R.b = i01 * i11; R.c = R.b; R.d = i11 * i11;
  return R;
}
template <typename T> __device__ __forceinline__ M2<T> ensure_pos2(const M2<T>& K) {   // utils.py:15-22
  if (K.a < T(0) || K.d < T(0)) { M2<T> R; R.a = (K.a < T(0)) ? T(1e-2) : K.a; R.d = (K.d < T(0)) ? T(1e-2) : K.d; R.b = T(0); R.c = T(0); return R; }
  return K;
}
template <typename T> struct Site2 { T lml; T m0, m1; M2<T> C; };

// HeteroscedasticNoise.moment_match, likelihoods.py:690-728: one-dimensional cubature over f2
template <typename T, class Red>
__device__ __forceinline__ Site2<T> hetero_moment_match(const SiteModel& sm, T y, T cm0, T cm1, const M2<T>& cc, T power) {
  const T s1 = sqrt(cc.d);
  T Z = T(0), dZ1 = T(0), dZ2 = T(0), d2Z1 = T(0), d2Z2 = T(0);
  for (int i = Red::first(); i < sm.n_cub; i += Red::stride()) {
    const T sp = s1 * T(sm.cub_x[i]) + cm1, w = T(sm.cub_w[i]);                            // :698
    const T u = sp - T(0.5);
    const T link = (sm.lik == KJX_LIK_HETERO_EXP) ? exp(u) : (log(T(1) + exp(-fabs(u))) + ((u < T(0)) ? T(0) : u)) + T(1e-10);   // :658-661
    const T f2 = link * link / power, obs_var = f2 + cc.a;                                 // :700-701
    const T cst = pow(power, T(-0.5)) * pow(T(2 * KJX_PI) * link * link, T(0.5) - T(0.5) * power);   // :702
    const T r = y - cm0;
    const T npdf = cst * pow(T(2 * KJX_PI) * obs_var, T(-0.5)) * exp(T(-0.5) * r * r / obs_var);
    const T e2 = (sp - cm1) / cc.d;
    Z += w * npdf;
    dZ1 += w * (r / obs_var * npdf);                                                       // :708
    dZ2 += w * (e2 * npdf);                                                                // :711
    d2Z1 += w * ((-(T(1) / obs_var) + (r / obs_var) * (r / obs_var)) * npdf);              // :714
    d2Z2 += w * ((-(T(1) / cc.d) + e2 * e2) * npdf);                                       // :717
  }
  Z = Red::sum(Z); dZ1 = Red::sum(dZ1); dZ2 = Red::sum(dZ2); d2Z1 = Red::sum(d2Z1); d2Z2 = Red::sum(d2Z2);
  const T Zc = max_nan(Z, T(1e-8)), Zinv = T(1) / Zc;
  Site2<T> o;
  o.lml = log(Zc);
  const T g1 = Zinv * dZ1, g2 = Zinv * dZ2;
  M2<T> nH; nH.a = -(-g1 * g1 + Zinv * d2Z1); nH.d = -(-g2 * g2 + Zinv * d2Z2); nH.b = T(0); nH.c = T(0);   // -d2lZ (diagonal)
  M2<T> K = ensure_pos2(nH); K.a -= T(1e-10); K.d -= T(1e-10);                             // :724 (minus jitter)
  const M2<T> iK = inv2(K);
  o.m0 = cm0 + iK.a * g1 + iK.b * g2; o.m1 = cm1 + iK.c * g1 + iK.d * g2;                  // :725
  o.C.a = power * (-cc.a + iK.a); o.C.b = power * (-cc.b + iK.b); o.C.c = power * (-cc.c + iK.c); o.C.d = power * (-cc.d + iK.d);
  return o;
}

// ExpectationPropagation.update (approximate_inference.py:50-73); has_prev = false: initialise from the given marginal
template <typename T, class Red>
__device__ __forceinline__ Site2<T> ep_update2(const SiteModel& sm, T y, T pm0, T pm1, const M2<T>& pc, bool has_prev,
                                               T sm0, T sm1, const M2<T>& sc, bool pure_moment_match) {
  if (pure_moment_match) return hetero_moment_match<T, Red>(sm, y, pm0, pm1, pc, T(sm.power));
  if (!has_prev) return hetero_moment_match<T, Red>(sm, y, pm0, pm1, pc, T(1));            // :57
  const T power = T(sm.power), damping = T(sm.damping);
  const M2<T> pp = inv2(pc), sp = inv2(sc);                                                // :12
  M2<T> cp; cp.a = pp.a - power * sp.a; cp.b = pp.b - power * sp.b; cp.c = pp.c - power * sp.c; cp.d = pp.d - power * sp.d;
  const M2<T> cc = inv2(cp);                                                               // :13
  const T v0 = pp.a * pm0 + pp.b * pm1 - power * (sp.a * sm0 + sp.b * sm1), v1 = pp.c * pm0 + pp.d * pm1 - power * (sp.c * sm0 + sp.d * sm1);
  const T cm0 = cc.a * v0 + cc.b * v1, cm1 = cc.c * v0 + cc.d * v1;                        // :14
  Site2<T> o = hetero_moment_match<T, Red>(sm, y, cm0, cm1, cc, power);                    // :66
  if (damping != T(1)) {                                                                   // :68-72
    const M2<T> n2 = inv2(o.C), n2p = inv2(sc);
    const T n10 = n2.a * o.m0 + n2.b * o.m1, n11 = n2.c * o.m0 + n2.d * o.m1;
    const T p10 = n2p.a * sm0 + n2p.b * sm1, p11 = n2p.c * sm0 + n2p.d * sm1;
    M2<T> q; q.a = (T(1) - damping) * n2p.a + damping * n2.a; q.b = (T(1) - damping) * n2p.b + damping * n2.b;
    q.c = (T(1) - damping) * n2p.c + damping * n2.c; q.d = (T(1) - damping) * n2p.d + damping * n2.d;
    o.C = inv2(q);
    const T w0 = (T(1) - damping) * p10 + damping * n10, w1 = (T(1) - damping) * p11 + damping * n11;
    o.m0 = o.C.a * w0 + o.C.b * w1; o.m1 = o.C.c * w0 + o.C.d * w1;
  }
  return o;
}

// link(f2) of HeteroscedasticNoise (likelihoods.py:657-661): exp(f - 0.5) or softplus(f - 0.5) + 1e-10
template <typename T> __device__ __forceinline__ T hetero_link(const SiteModel& sm, T f) {
  const T u = f - T(0.5);
  return (sm.lik == KJX_LIK_HETERO_EXP) ? exp(u) : (log(T(1) + exp(-fabs(u))) + ((u < T(0)) ? T(0) : u)) + T(1e-10);
}
template <typename T> __device__ __forceinline__ M2<T> add_jitter2(const M2<T>& P, T j) { M2<T> R = P; R.a += j; R.d += j; return R; }
template <typename T> __device__ __forceinline__ M2<T> mix2(T wa, const M2<T>& A, T wb, const M2<T>& B) {
  M2<T> R; R.a = wa * A.a + wb * B.a; R.b = wa * A.b + wb * B.b; R.c = wa * A.c + wb * B.c; R.d = wa * A.d + wb * B.d; return R;
}
// cavity, approximate_inference.py:8-15 in matrix form
template <typename T>
__device__ __forceinline__ void cavity2(T pm0, T pm1, const M2<T>& pc, T sm0, T sm1, const M2<T>& sc, T power, T& cm0, T& cm1, M2<T>& cc) {
  const M2<T> pp = inv2(pc), sp = inv2(sc);
  cc = inv2(mix2(T(1), pp, -power, sp));
  const T v0 = pp.a * pm0 + pp.b * pm1 - power * (sp.a * sm0 + sp.b * sm1), v1 = pp.c * pm0 + pp.d * pm1 - power * (sp.c * sm0 + sp.d * sm1);
  cm0 = cc.a * v0 + cc.b * v1; cm1 = cc.c * v0 + cc.d * v1;
}
// the damping step of EEP / SLEP (approximate_inference.py:125-132, :203-209): natural parameters of the new and the
// previous site mixed; n2 = precision of the new site as the rule defines it
template <typename T>
__device__ __forceinline__ void damp2(Site2<T>& o, const M2<T>& n2, T sm0, T sm1, const M2<T>& sc, T damping) {
  const M2<T> n2p = inv2(add_jitter2(sc, T(1e-10)));
  const T n10 = n2.a * o.m0 + n2.b * o.m1, n11 = n2.c * o.m0 + n2.d * o.m1;
  const T p10 = n2p.a * sm0 + n2p.b * sm1, p11 = n2p.c * sm0 + n2p.d * sm1;
  o.C = inv2(add_jitter2(mix2(T(1) - damping, n2p, damping, n2), T(1e-10)));
  const T w0 = (T(1) - damping) * p10 + damping * n10, w1 = (T(1) - damping) * p11 + damping * n11;
  o.m0 = o.C.a * w0 + o.C.b * w1; o.m1 = o.C.c * w0 + o.C.d * w1;
}

// ExtendedEP.update (EKS: power 0), approximate_inference.py:99-133, with HeteroscedasticNoise.analytical_linearisation
// (likelihoods.py:845-850) evaluated at sigma = 0 as the rule calls it: Jf = [1, 0], Jsigma = link(m2)
template <typename T>
__device__ __forceinline__ Site2<T> eep_update2(const SiteModel& sm, T y, T pm0, T pm1, const M2<T>& pc, bool has_prev,
                                                T sm0, T sm1, const M2<T>& sc) {
  const T power = has_prev ? T(sm.power) : T(1), damping = T(sm.damping);
  T cm0 = pm0, cm1 = pm1; M2<T> cc = pc;
  if (has_prev && power != T(0)) cavity2<T>(pm0, pm1, pc, sm0, sm1, sc, power, cm0, cm1, cc);   // :103-108
  const T link = hetero_link<T>(sm, cm1), jj = link * link;                               // Jsigma obs_cov Jsigma^T
  const T r = y - cm0;                                                                     // :113 (E[y|f] = f1)
  const T sigma = jj + power * cc.a;                                                       // :114
  M2<T> n2; n2.a = inv1(jj); n2.b = n2.c = n2.d = T(0);                                    // :115  Jf^T inv(.) Jf
  Site2<T> o;
  o.C = inv2(add_jitter2(n2, T(1e-10)));                                                   // :116
  const T g = inv1(sigma) * r;
  o.m0 = cm0 + (o.C.a + power * cc.a) * g; o.m1 = cm1 + (o.C.c + power * cc.c) * g;        // :117
  const T sml = jj + cc.a;                                                                 // :119
  o.lml = -(T(0.5) * T(2) * log(T(2 * KJX_PI)) + log(sqrt(sml)) + T(0.5) * (r * (r / sml)));   // :120-123 (site_cov.shape[0] = 2)
  if (has_prev && damping != T(1)) damp2<T>(o, n2, sm0, sm1, sc, damping);                 // :124-132
  return o;
}

// the reference's two-dimensional symmetric rules, 4-digit literals as written (utils.py:484-488, 512-516)
__device__ __forceinline__ void sym_rule2(int rule, int q, double& x0, double& x1, double& w) {
  if (rule == 3) {
    const double u = 1.4142;
    const double X0[5] = {0., u, 0., -u, 0.}, X1[5] = {0., 0., u, 0., -u};
    x0 = X0[q]; x1 = X1[q]; w = q ? 0.25 : 0.;
  } else {
    const double u = 1.7321;
    const double X0[9] = {0., u, -u, 0., 0., u, -u, u, -u}, X1[9] = {0., 0., 0., u, -u, u, -u, -u, u};
    x0 = X0[q]; x1 = X1[q]; w = (q == 0) ? 0.4444 : (q < 5 ? 0.1111 : 0.0278);
  }
}
// StatisticallyLinearisedEP.update, approximate_inference.py:180-210, with HeteroscedasticNoise.statistical_linear_regression
// (likelihoods.py:808-842): a TWO-dimensional rule over (f1, f2), sigma points through the Cholesky factor of the cavity
template <typename T, class Red>
__device__ __forceinline__ Site2<T> slep_update2(const SiteModel& sm, T y, T pm0, T pm1, const M2<T>& pc, bool has_prev,
                                                 T sm0, T sm1, const M2<T>& sc) {
  const T power = has_prev ? T(sm.power) : T(1), damping = T(sm.damping);
  Site2<T> o;
  o.lml = hetero_moment_match<T, Red>(sm, y, pm0, pm1, pc, T(1)).lml;                      // :187
  T cm0 = pm0, cm1 = pm1; M2<T> cc = pc;
  if (has_prev && power != T(0)) cavity2<T>(pm0, pm1, pc, sm0, sm1, sc, power, cm0, cm1, cc);   // :188-193
  const T l00 = sqrt(cc.a), l10 = cc.c / l00, l11 = sqrt(cc.d - l10 * l10);                // cholesky(cav_cov), lower
  const int n1 = sm.n_cub, nq = (sm.cub_rule == 3) ? 5 : (sm.cub_rule == 5) ? 9 : n1 * n1;
  T Sv = T(0), C0 = T(0), C1 = T(0);
  for (int q = Red::first(); q < nq; q += Red::stride()) {
    double x0, x1, w;
    if (sm.cub_rule == 0) { const int a = q / n1, b = q - a * n1; x0 = sm.cub_x[a]; x1 = sm.cub_x[b]; w = sm.cub_w[a] * sm.cub_w[b]; }   // utils.py:448-450, 461-462
    else sym_rule2(sm.cub_rule, q, x0, x1, w);
    const T d0 = l00 * T(x0), d1 = l10 * T(x0) + l11 * T(x1);                              // sigma point minus cavity mean (:820)
    const T link = hetero_link<T>(sm, d1 + cm1);
    Sv += T(w) * (link * link);                                                            // :829-831
    C0 += T(w) * (d0 * d0); C1 += T(w) * (d1 * d0);                                        // :836-838
  }
  Sv = Red::sum(Sv); C0 = Red::sum(C0); C1 = Red::sum(C1);
  const T S = cc.a + Sv, r = y - cm0;                                                      // mu = m0 (:825)
  const M2<T> ic = inv2(add_jitter2(cc, T(1e-10)));
  const T sigma = S + (power - T(1)) * (C0 * (ic.a * C0 + ic.b * C1) + C1 * (ic.c * C0 + ic.d * C1));   // :197
  const T isg = inv1(sigma);
  M2<T> oso; oso.a = isg; oso.b = oso.c = oso.d = T(0);                                    // omega = [1, 0]; diagonal kept (:198-199)
  const M2<T> osigo = inv2(add_jitter2(oso, T(1e-10)));                                    // :200
  o.m0 = cm0 + osigo.a * (isg * r); o.m1 = cm1 + osigo.c * (isg * r);                      // :201
  o.C = mix2(-power, cc, T(1), osigo);                                                     // :202
  if (has_prev && damping != T(1)) damp2<T>(o, inv2(add_jitter2(o.C, T(1e-10))), sm0, sm1, sc, damping);   // :203-209
  return o;
}

// VariationalInference.update, approximate_inference.py:302-323, with HeteroscedasticNoise.variational_expectation
// (likelihoods.py:763-805): one-dimensional cubature over f2
template <typename T, class Red>
__device__ __forceinline__ Site2<T> vi_update2(const SiteModel& sm, T y, T pm0, T pm1, const M2<T>& pc, bool has_prev,
                                               T sm0, T sm1, const M2<T>& sc) {
  const T damping = T(sm.damping);
  const T s1 = sqrt(pc.d), v0 = pc.a, v1 = pc.d;
  T dm1 = T(0), dm2 = T(0), dv1 = T(0), dv2 = T(0);
  for (int i = Red::first(); i < sm.n_cub; i += Red::stride()) {
    const T sp = s1 * T(sm.cub_x[i]) + pm1, w = T(sm.cub_w[i]);                            // :772
    const T link = hetero_link<T>(sm, sp), var = link * link, iv = T(1) / var;
    const T wll = w * (log(var) + iv * ((y - pm0) * (y - pm0) + v0));                      // :775-776
    dm1 += (iv * (y - pm0 + v0)) * w;                                                      // :784-786 (as written)
    dm2 += wll * (T(1) / v1) * (sp - pm1);                                                 // :789-791
    dv1 += iv * w;                                                                         // :793-795
    dv2 += ((sp - pm1) * (sp - pm1) / (v1 * v1) - T(1) / v1) * wll;                        // :798-801
  }
  dm1 = Red::sum(dm1); dm2 = T(-0.5) * Red::sum(dm2); dv1 = T(-0.5) * Red::sum(dv1); dv2 = T(-0.25) * Red::sum(dv2);
  M2<T> nv; nv.a = -dv1; nv.d = -dv2; nv.b = nv.c = T(0);
  const M2<T> K = ensure_pos2(nv);                                                         // ensure_positive_precision(-dE_dv)
  Site2<T> o;
  if (!has_prev) {
    const M2<T> h = inv2(add_jitter2(K, T(1e-10)));                                        // :311
    o.C.a = T(0.5) * h.a; o.C.b = T(0.5) * h.b; o.C.c = T(0.5) * h.c; o.C.d = T(0.5) * h.d;
    o.m0 = pm0 + o.C.a * dm1 + o.C.b * dm2; o.m1 = pm1 + o.C.c * dm1 + o.C.d * dm2;        // :312
  } else {
    const M2<T> l2p = inv2(add_jitter2(sc, T(1e-10)));                                     // :317
    const T l10 = l2p.a * sm0 + l2p.b * sm1, l11 = l2p.c * sm0 + l2p.d * sm1;              // :318
    // dE_dv = -K: lambda_1 <- (1-d) lambda_1 + d (dE_dm + 2 K m),  lambda_2 <- (1-d) lambda_2 + d 2 K     (:319-320)
    const T t0 = dm1 + T(2) * (K.a * pm0 + K.b * pm1), t1 = dm2 + T(2) * (K.c * pm0 + K.d * pm1);
    const T w0 = (T(1) - damping) * l10 + damping * t0, w1 = (T(1) - damping) * l11 + damping * t1;
    const M2<T> l2 = mix2(T(1) - damping, l2p, T(2) * damping, K);
    o.C = inv2(add_jitter2(l2, T(1e-10)));                                                 // :321
    o.m0 = o.C.a * w0 + o.C.b * w1; o.m1 = o.C.c * w0 + o.C.d * w1;                        // :322
  }
  o.lml = hetero_moment_match<T, Red>(sm, y, pm0, pm1, pc, T(1)).lml;                      // :323
  return o;
}

// the update rule the model names (kjx_method); pure_moment_match: Likelihood.moment_match itself (kjx_site_update only)
template <typename T, class Red>
__device__ __forceinline__ Site2<T> site_update2(const SiteModel& sm, T y, T pm0, T pm1, const M2<T>& pc, bool has_prev,
                                                 T sm0, T sm1, const M2<T>& sc, bool pure_moment_match) {
  if (!pure_moment_match) {
    if (sm.method == KJX_INF_VI) return vi_update2<T, Red>(sm, y, pm0, pm1, pc, has_prev, sm0, sm1, sc);
    if (sm.method == KJX_INF_SLEP) return slep_update2<T, Red>(sm, y, pm0, pm1, pc, has_prev, sm0, sm1, sc);
    if (sm.method == KJX_INF_EEP) return eep_update2<T>(sm, y, pm0, pm1, pc, has_prev, sm0, sm1, sc);
  }
  return ep_update2<T, Red>(sm, y, pm0, pm1, pc, has_prev, sm0, sm1, sc, pure_moment_match);
}

// ---- dense d x d helpers (d <= 8; every lane holds the same values) --------------------------------------------------
template <typename T> struct MultiState { T m[KJX_MULTI_DMAX]; T P[KJX_MULTI_DMAX][KJX_MULTI_DMAX]; };

template <typename T>
__device__ __forceinline__ void multi_transition(const DiscArgs& da, int d, T dt, T (&A)[KJX_MULTI_DMAX][KJX_MULTI_DMAX]) {
  for (int i = 0; i < d; ++i) {
    for (int j = 0; j < d; ++j) A[i][j] = T(0);
    int c0, nnz; T v[KJX_ROW_W] = {};
    transition_row<T>(da, i, dt, c0, nnz, v);
    for (int q = 0; q < nnz; ++q) A[i][c0 + q] = v[q];
  }
}
// Pp = A (P - Pinf) A^T + Pinf  (sde_gp.py:278, :354);  mp = A m
template <typename T>
__device__ __forceinline__ void multi_predict(int d, const T (&A)[KJX_MULTI_DMAX][KJX_MULTI_DMAX], const double* Pinf, const T* m,
                                              const T (&P)[KJX_MULTI_DMAX][KJX_MULTI_DMAX], T* mp, T (&Pp)[KJX_MULTI_DMAX][KJX_MULTI_DMAX]) {
  T Y[KJX_MULTI_DMAX][KJX_MULTI_DMAX];
  for (int i = 0; i < d; ++i) {
    T s = T(0);
    for (int k = 0; k < d; ++k) s += A[i][k] * m[k];
    mp[i] = s;
    for (int j = 0; j < d; ++j) { T t = T(0); for (int k = 0; k < d; ++k) t += A[i][k] * (P[k][j] - (T)Pinf[k * d + j]); Y[i][j] = t; }
  }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) { T t = T(0); for (int k = 0; k < d; ++k) t += Y[i][k] * A[j][k]; Pp[i][j] = t + (T)Pinf[i * d + j]; }
}
// X = solve(Pp, B) by Cholesky (utils.py:25-30)
template <typename T>
__device__ __forceinline__ void multi_chol_solve(int d, const T (&Pp)[KJX_MULTI_DMAX][KJX_MULTI_DMAX], const T (&B)[KJX_MULTI_DMAX][KJX_MULTI_DMAX],
                                                 T (&X)[KJX_MULTI_DMAX][KJX_MULTI_DMAX]) {
  T L[KJX_MULTI_DMAX][KJX_MULTI_DMAX];
  for (int j = 0; j < d; ++j) {
    T s = Pp[j][j];
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    L[j][j] = sqrt(s);
    for (int i = j + 1; i < d; ++i) { T v = Pp[i][j]; for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k]; L[i][j] = v / L[j][j]; }
  }
  for (int c = 0; c < d; ++c) {
    T z[KJX_MULTI_DMAX];
    for (int i = 0; i < d; ++i) { T v = B[i][c]; for (int k = 0; k < i; ++k) v -= L[i][k] * z[k]; z[i] = v / L[i][i]; }
    for (int i = d - 1; i >= 0; --i) { T v = z[i]; for (int k = i + 1; k < d; ++k) v -= L[k][i] * X[k][c]; X[i][c] = v / L[i][i]; }
  }
}
// H m (2) and H P H^T (2x2), H P (2 x d)
template <typename T>
__device__ __forceinline__ void multi_project(int d, const double* H, const T* m, const T (&P)[KJX_MULTI_DMAX][KJX_MULTI_DMAX], T& f0, T& f1,
                                              M2<T>& V, T (*HP)[KJX_MULTI_DMAX]) {
  T hp[2][KJX_MULTI_DMAX];
  f0 = T(0); f1 = T(0);
  for (int k = 0; k < d; ++k) { f0 += (T)H[k] * m[k]; f1 += (T)H[d + k] * m[k]; }
  for (int a = 0; a < 2; ++a)
    for (int j = 0; j < d; ++j) { T s = T(0); for (int k = 0; k < d; ++k) s += (T)H[a * d + k] * P[k][j]; hp[a][j] = s; if (HP) HP[a][j] = s; }
  V.a = V.b = V.c = V.d = T(0);
  for (int k = 0; k < d; ++k) { V.a += hp[0][k] * (T)H[k]; V.b += hp[0][k] * (T)H[d + k]; V.c += hp[1][k] * (T)H[k]; V.d += hp[1][k] * (T)H[d + k]; }
}

// ---- forward filter, sde_gp.py:271-306 ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32) multi_filter_kernel(const MultiArgs<T> g) {
  const int d = g.d; const bool lead = threadIdx.x == 0;
  MultiState<T> x;
  for (int i = 0; i < d; ++i) { x.m[i] = T(0); for (int j = 0; j < d; ++j) x.P[i][j] = (T)g.Pinf[i * d + j]; }   // :265
  T nl = T(0);
  for (int64_t n = 0; n < g.N; ++n) {
    T A[KJX_MULTI_DMAX][KJX_MULTI_DMAX], mp[KJX_MULTI_DMAX], Pp[KJX_MULTI_DMAX][KJX_MULTI_DMAX], HP[2][KJX_MULTI_DMAX];
    multi_transition<T>(g.disc, d, g.dt[n], A);                                            // :276
    multi_predict<T>(d, A, g.Pinf, x.m, x.P, mp, Pp);                                      // :277-278
    T pm0, pm1; M2<T> pv;
    multi_project<T>(d, g.H, mp, Pp, pm0, pm1, pv, HP);                                    // :283-284
    const bool masked = g.mask ? (g.mask[n] != 0) : false;
    const T yn = masked ? pm0 : g.y[n];                                                    // :286 (y has one column)
    M2<T> none; none.a = none.b = none.c = none.d = T(0);
    Site2<T> s = site_update2<T, WarpRed>(g.site, yn, pm0, pm1, pv, false, T(0), T(0), none, false);   // :287
    if (!g.init_sites) {                                                                   // :290
      s.m0 = g.site_mean[2 * n]; s.m1 = g.site_mean[2 * n + 1];
      s.C.a = g.site_cov[4 * n]; s.C.b = g.site_cov[4 * n + 1]; s.C.c = g.site_cov[4 * n + 2]; s.C.d = g.site_cov[4 * n + 3];
    }
    if (!masked) {
      M2<T> S; S.a = pv.a + s.C.a; S.b = pv.b + s.C.b; S.c = pv.c + s.C.c; S.d = pv.d + s.C.d;   // :292
      const M2<T> Si = inv2(S);                                                            // :294  K = solve(S, HP)^T
      const T r0 = s.m0 - pm0, r1 = s.m1 - pm1;
      T K[KJX_MULTI_DMAX][2];
      for (int i = 0; i < d; ++i) { K[i][0] = HP[0][i] * Si.a + HP[1][i] * Si.c; K[i][1] = HP[0][i] * Si.b + HP[1][i] * Si.d; }
      for (int i = 0; i < d; ++i) {
        x.m[i] = mp[i] + K[i][0] * r0 + K[i][1] * r1;                                      // :295
        for (int j = 0; j < d; ++j) x.P[i][j] = Pp[i][j] - (K[i][0] * HP[0][j] + K[i][1] * HP[1][j]);   // :296
      }
      nl += s.lml;                                                                         // :301
    } else {
      for (int i = 0; i < d; ++i) { x.m[i] = mp[i]; for (int j = 0; j < d; ++j) x.P[i][j] = Pp[i][j]; }   // :297-300
    }
    if (lead) {
      if (g.filt_mean) { for (int i = 0; i < d; ++i) { g.filt_mean[n * d + i] = x.m[i]; for (int j = 0; j < d; ++j) g.filt_cov[(n * d + i) * d + j] = x.P[i][j]; } }
      if (g.out_site_mean) {
        g.out_site_mean[2 * n] = s.m0; g.out_site_mean[2 * n + 1] = s.m1;
        g.out_site_cov[4 * n] = s.C.a; g.out_site_cov[4 * n + 1] = s.C.b; g.out_site_cov[4 * n + 2] = s.C.c; g.out_site_cov[4 * n + 3] = s.C.d;
      }
    }
  }
  if (lead && g.nlml) g.nlml[0] = -nl;
}

// ---- backward smoother, sde_gp.py:337-379 ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32) multi_smoother_kernel(const MultiArgs<T> g) {
  const int d = g.d; const bool lead = threadIdx.x == 0;
  MultiState<T> x;
  const int64_t last = g.N - 1;
  for (int i = 0; i < d; ++i) { x.m[i] = g.filt_mean_in[last * d + i]; for (int j = 0; j < d; ++j) x.P[i][j] = g.filt_cov_in[(last * d + i) * d + j]; }   // :339
  for (int64_t n = last; n >= 0; --n) {
    T A[KJX_MULTI_DMAX][KJX_MULTI_DMAX], mf[KJX_MULTI_DMAX], Pf[KJX_MULTI_DMAX][KJX_MULTI_DMAX], mp[KJX_MULTI_DMAX], Pp[KJX_MULTI_DMAX][KJX_MULTI_DMAX];
    T AP[KJX_MULTI_DMAX][KJX_MULTI_DMAX], Gt[KJX_MULTI_DMAX][KJX_MULTI_DMAX];
    multi_transition<T>(g.disc, d, (n + 1 < g.N) ? g.dt[n + 1] : T(0), A);                 // :337, :351
    for (int i = 0; i < d; ++i) { mf[i] = g.filt_mean_in[n * d + i]; for (int j = 0; j < d; ++j) Pf[i][j] = g.filt_cov_in[(n * d + i) * d + j]; }
    multi_predict<T>(d, A, g.Pinf, mf, Pf, mp, Pp);                                        // :352, :354
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { T s = T(0); for (int k = 0; k < d; ++k) s += A[i][k] * Pf[k][j]; AP[i][j] = s; }   // :353
    multi_chol_solve<T>(d, Pp, AP, Gt);                                                    // :359
    T mn[KJX_MULTI_DMAX], W[KJX_MULTI_DMAX][KJX_MULTI_DMAX], Pn[KJX_MULTI_DMAX][KJX_MULTI_DMAX];
    for (int i = 0; i < d; ++i) {
      T s = mf[i];
      for (int k = 0; k < d; ++k) s += Gt[k][i] * (x.m[k] - mp[k]);
      mn[i] = s;                                                                           // :360
      for (int j = 0; j < d; ++j) { T t = T(0); for (int k = 0; k < d; ++k) t += Gt[k][i] * (x.P[k][j] - Pp[k][j]); W[i][j] = t; }
    }
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { T t = Pf[i][j]; for (int k = 0; k < d; ++k) t += W[i][k] * Gt[k][j]; Pn[i][j] = t; }   // :361
    for (int i = 0; i < d; ++i) { x.m[i] = mn[i]; for (int j = 0; j < d; ++j) x.P[i][j] = Pn[i][j]; }
    T f0, f1; M2<T> V;
    multi_project<T>(d, g.H, x.m, x.P, f0, f1, V, (T(*)[KJX_MULTI_DMAX]) nullptr);          // :368-369, :373
    if (lead && g.smooth_mean) {
      if (g.return_full) { for (int i = 0; i < d; ++i) { g.smooth_mean[n * d + i] = x.m[i]; for (int j = 0; j < d; ++j) g.smooth_cov[(n * d + i) * d + j] = x.P[i][j]; } }
      else { g.smooth_mean[2 * n] = f0; g.smooth_mean[2 * n + 1] = f1; g.smooth_cov[4 * n] = V.a; g.smooth_cov[4 * n + 1] = V.b; g.smooth_cov[4 * n + 2] = V.c; g.smooth_cov[4 * n + 3] = V.d; }
    }
    if (g.update_sites) {                                                                  // :375-379
      M2<T> sc; sc.a = g.site_cov[4 * n]; sc.b = g.site_cov[4 * n + 1]; sc.c = g.site_cov[4 * n + 2]; sc.d = g.site_cov[4 * n + 3];
      const T p0 = g.site_mean[2 * n], p1 = g.site_mean[2 * n + 1];
      __syncwarp();                      // every lane has read the previous site before lane 0 may overwrite it in place
      const Site2<T> s = site_update2<T, WarpRed>(g.site, g.y[n], f0, f1, V, true, p0, p1, sc, false);
      if (lead) {
        g.new_site_mean[2 * n] = s.m0; g.new_site_mean[2 * n + 1] = s.m1;
        g.new_site_cov[4 * n] = s.C.a; g.new_site_cov[4 * n + 1] = s.C.b; g.new_site_cov[4 * n + 2] = s.C.c; g.new_site_cov[4 * n + 3] = s.C.d;
      }
    }
  }
}

// ---- batched site update / moment match over N entries (kjx_site_update with f = 2) ---------------------------------
template <typename T>
__global__ void __launch_bounds__(128) multi_site_update_kernel(const SiteModel sm, int64_t N, const T* __restrict__ y, const T* __restrict__ pm,
                                                                const T* __restrict__ pv, const T* sprev_mean, const T* sprev_cov,
                                                                T* log_lik, T* site_mean, T* site_cov) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  M2<T> V; V.a = pv[4 * k]; V.b = pv[4 * k + 1]; V.c = pv[4 * k + 2]; V.d = pv[4 * k + 3];
  M2<T> sc; sc.a = sc.b = sc.c = sc.d = T(0);
  T p0 = T(0), p1 = T(0);
  const bool has_prev = sprev_mean != nullptr;
  if (has_prev) { p0 = sprev_mean[2 * k]; p1 = sprev_mean[2 * k + 1]; sc.a = sprev_cov[4 * k]; sc.b = sprev_cov[4 * k + 1]; sc.c = sprev_cov[4 * k + 2]; sc.d = sprev_cov[4 * k + 3]; }
  const Site2<T> s = site_update2<T, SerialRed>(sm, y[k], pm[2 * k], pm[2 * k + 1], V, has_prev, p0, p1, sc, sm.method == KJX_INF_MOMENT_MATCH);
  if (log_lik) log_lik[k] = s.lml;
  if (site_mean) {
    site_mean[2 * k] = s.m0; site_mean[2 * k + 1] = s.m1;
    site_cov[4 * k] = s.C.a; site_cov[4 * k + 1] = s.C.b; site_cov[4 * k + 2] = s.C.c; site_cov[4 * k + 3] = s.C.d;
  }
}

}  // namespace kjx
