"""ORACLE (test infrastructure, not product): the sequential Kalman filter and RTS smoother.

NumPy fp64 restatement of /root/reference/kalmanjax/sde_gp.py:85-103 (predict_everywhere),
:110-143 (NLPD), :163-188 (run, without the autodiff gradient), :230-309 (kalman_filter) and
:311-384 (rauch_tung_striebel_smoother), plus utils.py:87-197 (input admin).  One python loop
iteration per time step, exactly like the reference's `for n in s.range(N)`.
Scalar sites only (func_dim f = 1), which covers every configuration in BASELINE.json.

Pinned by tests/test_oracle_golden.py against fixtures recorded from the unmodified reference
(tests/golden/make_golden.py) — including the notebook known-answers K1, K2, K3, K5.
"""
import numpy as np
import scipy.linalg as sl

from . import approx_inf as ai
from . import cubature


def solve(P, Q):
    """utils.py:25-30 — Cholesky solve; a failed factorisation yields NaN, not an exception."""
    try:
        return sl.cho_solve(sl.cho_factor(P, check_finite=False), Q, check_finite=False)
    except (sl.LinAlgError, ValueError):
        return np.full(np.shape(Q), np.nan)


def input_admin(t, y, r=None):
    """utils.py:87-114."""
    t, y = np.asarray(t, dtype=float), np.asarray(y, dtype=float)
    assert t.shape[0] == y.shape[0]
    if t.ndim < 2:
        t = t[:, None]
    if y.ndim < 2:
        y = y[:, None]
    if r is None:
        r = np.nan * t
    r = np.asarray(r, dtype=float)
    if r.ndim < 2:
        r = r[:, None]
    ind = np.argsort(t[:, 0], axis=0)
    t, y, r = t[ind], y[ind], r[ind]
    dt = np.concatenate([np.array([0.0]), np.diff(t[:, 0])])
    return t, y, r, dt


def test_input_admin(t, y, r, t_test, y_test, r_test):
    """utils.py:117-197 (train/test merge, sort, mask).  Inputs t,y,r are already sorted train data."""
    t_train, y_train, r_train = t, y, r
    if t_test is None:
        t_test = np.empty((1,) + t_train.shape[1:]) * np.nan
        r_test = np.empty((1,) + t_train.shape[1:]) * np.nan
    else:
        t_test = np.asarray(t_test, dtype=float)
        if t_test.ndim < 2:
            t_test = t_test[:, None]
        test_sort_ind = np.argsort(t_test[:, 0], axis=0)
        t_test = t_test[test_sort_ind]
        if y_test is not None:
            y_test = np.asarray(y_test, dtype=float)[test_sort_ind].reshape((-1,) + y.shape[1:])
        if r_test is not None:
            r_test = np.asarray(r_test, dtype=float)
            if r_test.ndim < 2:
                r_test = r_test[:, None]
            r_test = r_test[test_sort_ind]
        else:
            r_test = np.nan * t_test
    t_train_test = np.concatenate([t_train, t_test])
    keep_ind = ~np.isnan(t_train_test[:, 0])
    t_train_test = t_train_test[keep_ind]
    r_test_nan = np.nan * np.zeros([r_test.shape[0], r_train.shape[1]]) if r_test.shape[1] != r_train.shape[1] else r_test
    r_train_test = np.concatenate([r_train, r_test_nan])[keep_ind]
    t_ind = np.argsort(t_train_test[:, 0])           # NB: default (unstable) sort, like the reference
    t_all, r_all = t_train_test[t_ind], r_train_test[t_ind]
    reverse_ind = np.argsort(t_ind)
    n_train = t_train.shape[0]
    train_id, test_id = reverse_ind[:n_train], reverse_ind[n_train:]
    y_all = np.nan * np.zeros([t_all.shape[0], y_train.shape[1]])
    y_all[train_id] = y_train
    if y_test is not None:
        y_all[test_id] = y_test
    mask = np.ones_like(y_all, dtype=bool)
    mask[train_id] = False
    dt_all = np.concatenate([np.array([0.0]), np.diff(t_all[:, 0])])
    return t_all, y_all, r_all, r_test, dt_all, train_id, test_id, mask


class SDEGP(object):
    """Same constructor and method names as sde_gp.py:12-50; hyperparameters live on the prior /
    likelihood objects as already-transformed values."""
    def __init__(self, prior, likelihood, t, y, r=None, approx_inf=None):
        self.t, self.y, self.r, self.dt = input_admin(t, y, r)
        self.prior, self.likelihood = prior, likelihood
        _, self.Pinf = prior.kernel_to_state_space()
        self.state_dim = self.Pinf.shape[0]
        self.func_dim = prior.measurement_model(self.r[0]).shape[0]
        assert self.func_dim == 1, "oracle restates scalar sites only"
        self.sites = ai.EP() if approx_inf is None else approx_inf

    # ------------------------------------------------------------------ sde_gp.py:230-309
    def kalman_filter(self, y, dt, store=False, mask=None, site_params=None, r=None):
        _, Pinf = self.prior.kernel_to_state_space()
        N, d = dt.shape[0], self.state_dim
        nlml = 0.0
        m, P = np.zeros([d, 1]), Pinf                                                     # :265
        if store:
            fm, fP = np.zeros([N, d, 1]), np.zeros([N, d, d])
            smean, scov = np.zeros([N, 1, 1]), np.zeros([N, 1, 1])
        for n in range(N):
            y_n = y[n][..., None]
            A = self.prior.state_transition(dt[n])                                        # :276
            m_ = A @ m                                                                    # :277
            P_ = A @ (P - Pinf) @ A.T + Pinf                                              # :278
            H = self.prior.measurement_model(r[n])                                        # :282
            predict_mean = H @ m_
            predict_cov = H @ P_ @ H.T
            if mask is not None:
                y_n = np.where(mask[n][..., None], predict_mean[:y_n.shape[0]], y_n)      # :286
            log_lik_n, site_mean, site_cov = self.sites.update(self.likelihood, y_n, predict_mean,
                                                               predict_cov, None)         # :287
            if site_params is not None:
                site_mean, site_cov = site_params[0][n], site_params[1][n]                # :290
            site_mean, site_cov = np.reshape(site_mean, (1, 1)), np.reshape(site_cov, (1, 1))
            S = predict_cov + site_cov                                                    # :292
            HP = H @ P_
            K = solve(S, HP).T
            m_new = m_ + K @ (site_mean - predict_mean)                                   # :295
            P_new = P_ - K @ HP                                                           # :296
            if mask is not None and np.any(mask[n]):                                      # :297-300
                m_new, P_new = m_, P_
                log_lik_n = np.where(mask[n][..., 0], 0.0, np.reshape(log_lik_n, -1))
            m, P = m_new, P_new
            nlml -= np.sum(log_lik_n)                                                     # :301
            if store:
                fm[n], fP[n], smean[n], scov[n] = m, P, site_mean, site_cov
        if store:
            return nlml, (fm, fP, (smean, scov))
        return nlml

    # ------------------------------------------------------------------ sde_gp.py:311-384
    def rauch_tung_striebel_smoother(self, m_filtered, P_filtered, dt, store=False, return_full=False,
                                     y=None, site_params=None, r=None):
        _, Pinf = self.prior.kernel_to_state_space()
        N, d = dt.shape[0], self.state_dim
        dt = np.concatenate([dt[1:], np.array([0.0])], axis=0)                            # :337
        m, P = m_filtered[-1], P_filtered[-1]
        k = d if return_full else 1
        sm, sc = np.zeros([N, k, 1]), np.zeros([N, k, k])
        if site_params is not None:
            new_mean, new_var = np.zeros([N, 1, 1]), np.zeros([N, 1, 1])
        for n in range(N - 1, -1, -1):
            A = self.prior.state_transition(dt[n])                                        # :351
            m_predicted = A @ m_filtered[n]
            tmp_gain_cov = A @ P_filtered[n]
            P_predicted = A @ (P_filtered[n] - Pinf) @ A.T + Pinf                         # :354
            G_transpose = solve(P_predicted, tmp_gain_cov)                                # :359
            m = m_filtered[n] + G_transpose.T @ (m - m_predicted)
            P = P_filtered[n] + G_transpose.T @ (P - P_predicted) @ G_transpose           # :361
            H = self.prior.measurement_model(r[n])
            if store:
                if return_full:
                    sm[n], sc[n] = m, P
                else:
                    sm[n], sc[n] = H @ m, H @ P @ H.T
            if site_params is not None:
                post_mean, post_cov = H @ m, H @ P @ H.T                                  # :373
                _, site_mu, site_cov = self.sites.update(self.likelihood, y[n][..., None], post_mean, post_cov,
                                                         (site_params[0][n], site_params[1][n]))   # :375
                new_mean[n], new_var[n] = np.reshape(site_mu, (1, 1)), np.reshape(site_cov, (1, 1))
        if site_params is not None:
            site_params = (new_mean, new_var)
        if store:
            return site_params, sm, sc
        return site_params

    # ------------------------------------------------------------------ sde_gp.py:163-188 (no gradient)
    def run(self):
        nlml, (fm, fP, sites) = self.kalman_filter(self.y, self.dt, True, None, self.sites.site_params, self.r)
        self.sites.site_params = sites
        self.sites.site_params = self.rauch_tung_striebel_smoother(fm, fP, self.dt, False, False, self.y,
                                                                   self.sites.site_params, self.r)
        return nlml

    # ------------------------------------------------------------------ sde_gp.py:85-103
    def predict_everywhere(self, y, r, dt, train_id, mask, site_params=None, sampling=False, return_full=False):
        site_params = self.sites.site_params if site_params is None else site_params
        if site_params is not None and not sampling:
            site_mean = np.zeros([dt.shape[0], 1, 1])
            site_cov = 1e5 * np.tile(np.eye(1), (dt.shape[0], 1, 1))                      # :95
            site_mean[train_id] += site_params[0]
            site_cov[train_id] = site_params[1]
            site_params = (site_mean, site_cov)
        _, (fm, fP, site_params) = self.kalman_filter(y, dt, True, mask, site_params, r)
        _, pm, pc = self.rauch_tung_striebel_smoother(fm, fP, dt, True, return_full, None, None, r)
        return pm, pc, site_params

    def predict(self, t=None, r=None):
        """sde_gp.py:52-83 (same-dimensional spatial test points only)."""
        (_, y, r_all, _, dt, train_id, test_id, mask) = test_input_admin(self.t, self.y, self.r, t, None, r)
        pm, pc, _ = self.predict_everywhere(y, r_all, dt, train_id, mask)
        return np.squeeze(pm[test_id]), np.squeeze(pc[test_id])

    def negative_log_predictive_density(self, t=None, y=None, r=None):
        """sde_gp.py:110-143."""
        (_, y_all, r_all, _, dt, train_id, test_id, mask) = test_input_admin(self.t, self.y, self.r, t, y, r)
        pm, pc, _ = self.predict_everywhere(y_all, r_all, dt, train_id, mask)
        # :139-142 passes cubature_func=None, so the likelihood falls back to 20-point Gauss-Hermite
        # (likelihoods.py:141-142) whatever rule the inference method was built with
        lpd, _, _ = self.likelihood.moment_match(y_all[test_id][:, 0], pm[test_id][:, 0, 0], pc[test_id][:, 0, 0],
                                                 1, cubature.gauss_hermite(20))
        return -np.mean(lpd)
