/* ORACLE (test infrastructure / CPU baseline, not product): plain-C restatement of one
 * SDEGP.run() iteration of the reference for a Matern-3/2 or Matern-5/2 prior with a Gaussian
 * likelihood and (power-)EP sites — BASELINE.json's headline configuration.
 *
 * Follows /root/reference/kalmanjax:
 *   sde_gp.py:271-306   kalman_filter loop            (predict, project, log Z, update, store)
 *   sde_gp.py:337-379   rauch_tung_striebel_smoother  (gain by Cholesky solve, smoothed moments, site update)
 *   priors.py:172-174, 246-255  closed-form A(dt)
 *   utils.py:25-38      Cholesky solve / inverse,  utils.py:219-244 Gaussian moment match
 *   approximate_inference.py:8-15, 60-73  cavity + EP update with damping
 * One strictly sequential chain, dense d x d loops in the reference's order of operations — this is the
 * stand-in for the XLA:CPU while-loop the reference compiles to (JAX is not installable here).
 * Checked against the NumPy oracle by tests/test_oracle_c.py.  Built by __graft_entry__.build().
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DMAX 3
static const double PI = 3.141592653589793;

static void transition(int d, double dt, double lam, double A[DMAX][DMAX]) {
  if (d == 2) { /* priors.py:172-174 */
    double e = exp(-dt * lam);
    A[0][0] = e * (dt * lam + 1.0); A[0][1] = e * (dt * 1.0);
    A[1][0] = e * (dt * (-(lam * lam))); A[1][1] = e * (dt * (-lam) + 1.0);
  } else {      /* priors.py:246-255 */
    double dtlam = dt * lam, e = exp(-dtlam), lam2 = lam * lam;
    A[0][0] = e * (dt * (lam * (0.5 * dtlam + 1.0)) + 1.0);
    A[0][1] = e * (dt * (dtlam + 1.0));
    A[0][2] = e * (dt * (0.5 * dt));
    A[1][0] = e * (dt * (-0.5 * dtlam * lam2));
    A[1][1] = e * (dt * (lam * (1.0 - dtlam)) + 1.0);
    A[1][2] = e * (dt * (1.0 - 0.5 * dtlam));
    A[2][0] = e * (dt * (lam2 * lam * (0.5 * dtlam - 1.0)));
    A[2][1] = e * (dt * (lam2 * (dtlam - 3.0)));
    A[2][2] = e * (dt * (lam * (0.5 * dtlam - 2.0)) + 1.0);
  }
}

/* X = A (P - Pinf) A^T + Pinf   (sde_gp.py:278, :354) */
static void predict_cov(int d, double A[DMAX][DMAX], const double* P, double Pinf[DMAX][DMAX], double X[DMAX][DMAX]) {
  double D[DMAX][DMAX], Y[DMAX][DMAX];
  for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) D[i][j] = P[i * d + j] - Pinf[i][j];
  for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { double s = 0; for (int k = 0; k < d; ++k) s += A[i][k] * D[k][j]; Y[i][j] = s; }
  for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { double s = 0; for (int k = 0; k < d; ++k) s += Y[i][k] * A[j][k]; X[i][j] = s + Pinf[i][j]; }
}

/* utils.py:25-30: solve P X = B by Cholesky (upper factor, like cho_factor's default) */
static void chol_solve(int d, double P[DMAX][DMAX], double B[DMAX][DMAX], double X[DMAX][DMAX]) {
  double L[DMAX][DMAX];
  for (int j = 0; j < d; ++j) {
    double s = P[j][j];
    for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
    L[j][j] = sqrt(s);
    for (int i = j + 1; i < d; ++i) { double v = P[i][j]; for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k]; L[i][j] = v / L[j][j]; }
  }
  for (int c = 0; c < d; ++c) {
    double z[DMAX];
    for (int i = 0; i < d; ++i) { double v = B[i][c]; for (int k = 0; k < i; ++k) v -= L[i][k] * z[k]; z[i] = v / L[i][i]; }
    for (int i = d - 1; i >= 0; --i) { double v = z[i]; for (int k = i + 1; k < d; ++k) v -= L[k][i] * X[k][c]; X[i][c] = v / L[i][i]; }
  }
}

static double inv1(double v) { double L = sqrt(v); return (1.0 / L) / L; } /* utils.py:33-38, 1x1 */

/* One run() iteration.  kind: 1 = Matern-3/2 (d=2), 2 = Matern-5/2 (d=3).  site_mean/site_var may be
 * NULL (first iteration: sites initialised to (y, lik_var), utils.py:241-244).  Returns nlml. */
double kjx_oracle_run(int kind, double variance, double lengthscale, double lik_var, double ep_power, double damping,
                      long N, const double* dt, const double* y, const double* site_mean, const double* site_var,
                      double* filt_mean, double* filt_cov, double* new_site_mean, double* new_site_var,
                      double* post_mean, double* post_var) {
  const int d = kind == 1 ? 2 : 3;
  double lam, Pinf[DMAX][DMAX];
  memset(Pinf, 0, sizeof(Pinf));
  if (kind == 1) { /* priors.py:148-155 */
    lam = sqrt(3.0) / lengthscale;
    Pinf[0][0] = variance; Pinf[1][1] = 3.0 * variance / (lengthscale * lengthscale);
  } else {         /* priors.py:219-230 */
    lam = sqrt(5.0) / lengthscale;
    double kappa = 5.0 / 3.0 * variance / (lengthscale * lengthscale);
    Pinf[0][0] = variance; Pinf[0][2] = -kappa; Pinf[1][1] = kappa; Pinf[2][0] = -kappa;
    Pinf[2][2] = 25.0 * variance / (lengthscale * lengthscale * lengthscale * lengthscale);
  }
  double m[DMAX] = {0, 0, 0}, P[DMAX * DMAX], A[DMAX][DMAX], Pp[DMAX][DMAX];
  for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) P[i * d + j] = Pinf[i][j];   /* sde_gp.py:265 */
  double nlml = 0.0;
  for (long n = 0; n < N; ++n) {                                                       /* sde_gp.py:271 */
    transition(d, dt[n], lam, A);
    double m_[DMAX];
    for (int i = 0; i < d; ++i) { double s = 0; for (int k = 0; k < d; ++k) s += A[i][k] * m[k]; m_[i] = s; }
    predict_cov(d, A, P, Pinf, Pp);
    const double pm = m_[0], pv = Pp[0][0];                                            /* H = [1,0,..] */
    const double lZ = -(y[n] - pm) * (y[n] - pm) / (lik_var + pv) / 2.0
                      - log(fmax(2.0 * PI * (lik_var + pv), 1e-10)) / 2.0;             /* utils.py:237-240 */
    const double smean = site_mean ? site_mean[n] : y[n], svar = site_var ? site_var[n] : lik_var;
    const double S = pv + svar, Sinv = inv1(S);
    double K[DMAX];
    for (int i = 0; i < d; ++i) K[i] = Pp[0][i] * Sinv;
    for (int i = 0; i < d; ++i) m[i] = m_[i] + K[i] * (smean - pm);
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) P[i * d + j] = Pp[i][j] - K[i] * Pp[0][j];
    nlml -= lZ;
    memcpy(filt_mean + n * d, m, sizeof(double) * d);
    memcpy(filt_cov + n * d * d, P, sizeof(double) * d * d);
  }
  double ms[DMAX], Ps[DMAX * DMAX];
  memcpy(ms, filt_mean + (N - 1) * d, sizeof(double) * d);                             /* sde_gp.py:339 */
  memcpy(Ps, filt_cov + (N - 1) * d * d, sizeof(double) * d * d);
  for (long n = N - 1; n >= 0; --n) {                                                  /* sde_gp.py:349 */
    const double* mf = filt_mean + n * d; const double* Pf = filt_cov + n * d * d;
    transition(d, n + 1 < N ? dt[n + 1] : 0.0, lam, A);                                /* sde_gp.py:337 */
    double mp[DMAX], T[DMAX][DMAX], G[DMAX][DMAX];
    for (int i = 0; i < d; ++i) { double s = 0; for (int k = 0; k < d; ++k) s += A[i][k] * mf[k]; mp[i] = s; }
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { double s = 0; for (int k = 0; k < d; ++k) s += A[i][k] * Pf[k * d + j]; T[i][j] = s; }
    predict_cov(d, A, Pf, Pinf, Pp);
    chol_solve(d, Pp, T, G);                                                           /* G = G_transpose */
    double dm[DMAX], dP[DMAX][DMAX], W[DMAX][DMAX], mn[DMAX], Pn[DMAX * DMAX];
    for (int i = 0; i < d; ++i) { dm[i] = ms[i] - mp[i]; for (int j = 0; j < d; ++j) dP[i][j] = Ps[i * d + j] - Pp[i][j]; }
    for (int i = 0; i < d; ++i) { double s = mf[i]; for (int k = 0; k < d; ++k) s += G[k][i] * dm[k]; mn[i] = s; }
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { double s = 0; for (int k = 0; k < d; ++k) s += G[k][i] * dP[k][j]; W[i][j] = s; }
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) { double s = Pf[i * d + j]; for (int k = 0; k < d; ++k) s += W[i][k] * G[k][j]; Pn[i * d + j] = s; }
    memcpy(ms, mn, sizeof(double) * d); memcpy(Ps, Pn, sizeof(double) * d * d);
    if (post_mean) { post_mean[n] = ms[0]; post_var[n] = Ps[0]; }
    /* EP site update (approximate_inference.py:60-73): cavity, Gaussian moment match, damping */
    const double smean = site_mean ? site_mean[n] : y[n], svar = site_var ? site_var[n] : lik_var;
    const double post_prec = inv1(Ps[0]), site_prec = inv1(svar);
    const double cav_cov = inv1(post_prec - ep_power * site_prec);
    const double cav_mean = cav_cov * (post_prec * ms[0] - ep_power * site_prec * smean);
    (void)cav_mean;                                     /* the Gaussian moment match returns (y, sigma^2) */
    double nm = y[n], nv = lik_var;
    if (damping != 1.0) {
      double nat2 = inv1(nv), nat2p = inv1(svar), nat1 = nat2 * nm, nat1p = nat2p * smean;
      nv = inv1((1.0 - damping) * nat2p + damping * nat2);
      nm = nv * ((1.0 - damping) * nat1p + damping * nat1);
    }
    new_site_mean[n] = nm; new_site_var[n] = nv;
  }
  return nlml;
}
