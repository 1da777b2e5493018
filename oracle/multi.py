"""ORACLE (test infrastructure, not product): the multi-latent case, func_dim f > 1 — `Independent` priors with the
`HeteroscedasticNoise` likelihood and every update rule (EP, EEP / EKS, the SLR family, VI) with f x f sites.

NumPy fp64 restatement of /root/reference/kalmanjax:
  priors.py:1417-1487            Independent (block-diagonal A, Pinf; block-diagonal H: one function value per component)
  likelihoods.py:647-728         HeteroscedasticNoise: p(y | f1, f2) = N(y | f1, link(f2)^2), moment match with a
                                 one-dimensional cubature over f2 (f1 is integrated analytically), diagonal Hessian
  likelihoods.py:763-850         its variational expectation (1-D cubature), statistical linear regression (2-D cubature,
                                 utils.py:418-516 through oracle/cubature.py) and analytical linearisation
  approximate_inference.py:8-15, 50-73   cavity and EP update in matrix form (Cholesky inverses, utils.py:33-38)
  approximate_inference.py:99-133, 180-210, 302-323   EEP / EKS, SLEP family and VI updates in matrix form
  sde_gp.py:230-384, 85-143      filter / smoother / predict / NLPD with matrix-valued sites
Pinned by the reference-generated fixtures `het_*` (tests/test_oracle_golden.py).
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve

from . import cubature
from .sde_gp import input_admin, test_input_admin

pi = 3.141592653589793


def solve(P, Q):      # utils.py:25-30; a failed factorisation yields NaN (as under JAX), not an exception
    try:
        return cho_solve(cho_factor(P, check_finite=False), Q, check_finite=False)
    except (np.linalg.LinAlgError, ValueError):
        return np.full(np.shape(Q), np.nan)


def inv(P):           # utils.py:33-38
    return solve(P, np.eye(P.shape[0]))


def chol_lower(P):    # jnp.linalg.cholesky: NaN instead of an exception
    try:
        return np.linalg.cholesky(P)
    except np.linalg.LinAlgError:
        return np.full(np.shape(P), np.nan)


def ensure_positive_precision(K):   # utils.py:15-22
    K_diag = np.diag(np.diag(K))
    return np.where(np.any(np.diag(K) < 0), np.where(K_diag < 0, 1e-2, K_diag), K)


def softplus(x):
    return np.log(1 + np.exp(-np.abs(x))) + np.maximum(x, 0)


class Independent(object):
    """priors.py:1417-1487."""
    def __init__(self, components):
        self.components = components
        self.name = "Independent"

    @staticmethod
    def _blockdiag(mats):
        n = sum(m.shape[0] for m in mats); k = sum(m.shape[1] for m in mats)
        out = np.zeros([n, k]); i = j = 0
        for m in mats:
            out[i:i + m.shape[0], j:j + m.shape[1]] = m; i += m.shape[0]; j += m.shape[1]
        return out

    def kernel_to_state_space(self):
        parts = [c.kernel_to_state_space() for c in self.components]
        return self._blockdiag([p[0] for p in parts]), self._blockdiag([p[1] for p in parts])   # H, Pinf

    def measurement_model(self, r=None):
        return self._blockdiag([c.measurement_model(r) for c in self.components])

    def state_transition(self, dt):
        return self._blockdiag([c.state_transition(dt) for c in self.components])

    @property
    def state_dim(self):
        return self.kernel_to_state_space()[1].shape[0]


class HeteroscedasticNoise(object):
    """likelihoods.py:647-728 (softplus link with its -0.5 offset and 1e-10 floor, :660-661)."""
    name = "Heteroscedastic Noise"

    def __init__(self, link="softplus"):
        if link == "exp":
            self.link_fn = lambda mu: np.exp(mu - 0.5)
        elif link == "softplus":
            self.link_fn = lambda mu: softplus(mu - 0.5) + 1e-10
        else:
            raise NotImplementedError("link function not implemented")
        self.hyp = np.array([])

    def moment_match(self, y, cav_mean, cav_cov, power, cub):
        """:690-728.  cub: the ONE-dimensional rule (the reference calls cubature_func(1))."""
        x, w = cub
        y = float(np.reshape(y, -1)[0])
        with np.errstate(all="ignore"):
            sigma_points = np.sqrt(cav_cov[1, 1]) * x + cav_mean[1, 0]                      # :698
            link = self.link_fn(sigma_points)
            f2 = link ** 2. / power                                                         # :700
            obs_var = f2 + cav_cov[0, 0]
            const = power ** -0.5 * (2 * pi * link ** 2.) ** (0.5 - 0.5 * power)            # :702
            normpdf = const * (2 * pi * obs_var) ** -0.5 * np.exp(-0.5 * (y - cav_mean[0, 0]) ** 2 / obs_var)
            Z = np.sum(w * normpdf)
            Zinv = 1. / np.maximum(Z, 1e-8)
            lZ = np.log(np.maximum(Z, 1e-8))
            dlZ1 = Zinv * np.sum(w * ((y - cav_mean[0, 0]) / obs_var * normpdf))            # :708-709
            dlZ2 = Zinv * np.sum(w * ((sigma_points - cav_mean[1, 0]) / cav_cov[1, 1] * normpdf))
            d2lZ1 = -dlZ1 ** 2 + Zinv * np.sum(w * ((-(f2 + cav_cov[0, 0]) ** -1 + ((y - cav_mean[0, 0]) / obs_var) ** 2) * normpdf))
            d2lZ2 = -dlZ2 ** 2 + Zinv * np.sum(w * ((-cav_cov[1, 1] ** -1 + ((sigma_points - cav_mean[1, 0]) / cav_cov[1, 1]) ** 2) * normpdf))
            dlZ = np.array([[dlZ1], [dlZ2]])
            d2lZ = np.array([[d2lZ1, 0.], [0., d2lZ2]])
            id2lZ = inv(ensure_positive_precision(-d2lZ) - 1e-10 * np.eye(2))               # :724
            site_mean = cav_mean + id2lZ @ dlZ
            site_cov = power * (-cav_cov + id2lZ)
        return lZ, site_mean, site_cov


    def variational_expectation(self, y, m, v, cub):
        """:763-805: one-dimensional cubature over f2; returns E[log p], dE/dm [2,1], dE/dv [2,2] (diagonal)."""
        x, w = cub
        y = float(np.reshape(y, -1)[0])
        m0, m1, v0, v1 = m[0, 0], m[1, 0], v[0, 0], v[1, 1]
        with np.errstate(all="ignore"):
            sigma_points = np.sqrt(v1) * x + m1
            var = self.link_fn(sigma_points) ** 2
            log_lik = np.log(var) + var ** -1 * ((y - m0) ** 2 + v0)
            wll = w * log_lik
            exp_log_lik = -0.5 * np.log(2 * pi) - 0.5 * np.sum(wll)
            dE_dm1 = np.sum((var ** -1 * (y - m0 + v0)) * w)                                # :784 (as written, "+ v0" included)
            dE_dm2 = -0.5 * np.sum(wll * v1 ** -1 * (sigma_points - m1))
            dE_dv1 = -0.5 * np.sum(var ** -1 * w)
            dE_dv2 = -0.25 * np.sum((v1 ** -2 * (sigma_points - m1) ** 2 - v1 ** -1) * wll)
        return exp_log_lik, np.array([[dE_dm1], [dE_dm2]]), np.array([[dE_dv1, 0.], [0., dE_dv2]])

    def statistical_linear_regression(self, cav_mean, cav_cov, cub2):
        """:808-842.  cub2: the TWO-dimensional rule (x [2,Q], w [Q])."""
        x, w = cub2
        m0, v0 = cav_mean[0, 0], cav_cov[0, 0]
        with np.errstate(all="ignore"):
            sigma_points = chol_lower(cav_cov) @ x + cav_mean
            var = self.link_fn(sigma_points[1]) ** 2
            S = (v0 + np.sum(w * var)).reshape(1, 1)
            C = np.sum(w * (sigma_points - cav_mean) * (sigma_points[0] - m0), axis=-1).reshape(2, 1)
        return m0.reshape(1, 1), S, C, np.array([[1., 0.]])

    def analytical_linearisation(self, m):
        """:845-850 with sigma = 0 (the only way the update rules call it): Jf = [1, 0], Jsigma = link(m[1])."""
        return np.array([[1.0, 0.0]]), np.reshape(self.link_fn(np.array([m[1, 0]])), (1, 1))


def compute_cavity(post_mean, post_cov, site_mean, site_cov, power):
    """approximate_inference.py:8-15."""
    post_precision, site_precision = inv(post_cov), inv(site_cov)
    cav_cov = inv(post_precision - power * site_precision)
    cav_mean = cav_cov @ (post_precision @ post_mean - power * site_precision @ site_mean)
    return cav_mean, cav_cov


def ep_update(rule, likelihood, y, post_mean, post_cov, site_params=None):
    """ExpectationPropagation.update, approximate_inference.py:50-73, matrix form."""
    cub = cubature.make(rule.intmethod, rule.num_cub_pts)
    with np.errstate(all="ignore"):
        if site_params is None:
            return likelihood.moment_match(y, post_mean, post_cov, 1.0, cub)                # :57
        site_mean_prev, site_cov_prev = site_params
        cav_mean, cav_cov = compute_cavity(post_mean, post_cov, site_mean_prev, site_cov_prev, rule.power)
        lml, site_mean, site_cov = likelihood.moment_match(y, cav_mean, cav_cov, rule.power, cub)
        if rule.damping != 1.:
            site_nat2, site_nat2_prev = inv(site_cov), inv(site_cov_prev)                   # :69-72
            site_nat1, site_nat1_prev = site_nat2 @ site_mean, site_nat2_prev @ site_mean_prev
            site_cov = inv((1. - rule.damping) * site_nat2_prev + rule.damping * site_nat2)
            site_mean = site_cov @ ((1. - rule.damping) * site_nat1_prev + rule.damping * site_nat1)
    return lml, site_mean, site_cov


def _damp(rule, site_mean, site_cov, site_nat2, site_params, jitter_new, jitter_prev, jitter_sum):
    """The damping step shared by the rules (each with its own jitter placement): approximate_inference.py:69-72,
    125-132, 203-209."""
    site_mean_prev, site_cov_prev = site_params
    eye = np.eye(site_cov.shape[0])
    nat2 = inv(site_cov + jitter_new * eye) if site_nat2 is None else site_nat2
    nat2_prev = inv(site_cov_prev + jitter_prev * eye)
    nat1, nat1_prev = nat2 @ site_mean, nat2_prev @ site_mean_prev
    site_cov = inv((1. - rule.damping) * nat2_prev + rule.damping * nat2 + jitter_sum * eye)
    site_mean = site_cov @ ((1. - rule.damping) * nat1_prev + rule.damping * nat1)
    return site_mean, site_cov


def eep_update(rule, likelihood, y, post_mean, post_cov, site_params=None):
    """ExtendedEP.update, approximate_inference.py:99-133, matrix form (EKS: power 0)."""
    y = np.reshape(np.asarray(y, dtype=float), (1, 1))
    power = 1. if site_params is None else rule.power
    with np.errstate(all="ignore"):
        if site_params is None or power == 0:
            cav_mean, cav_cov = post_mean, post_cov
        else:
            cav_mean, cav_cov = compute_cavity(post_mean, post_cov, site_params[0], site_params[1], power)
        Jf, Jsigma = likelihood.analytical_linearisation(cav_mean)
        residual = y - cav_mean[0:1, 0:1]                                                   # conditional_moments: E[y|f] = f1
        sigma = Jsigma @ Jsigma.T + power * Jf @ cav_cov @ Jf.T
        site_nat2 = Jf.T @ inv(Jsigma @ Jsigma.T) @ Jf
        site_cov = inv(site_nat2 + 1e-10 * np.eye(2))
        site_mean = cav_mean + (site_cov + power * cav_cov) @ Jf.T @ inv(sigma) @ residual
        sigma_marg = Jsigma @ Jsigma.T + Jf @ cav_cov @ Jf.T
        chol = chol_lower(sigma_marg)
        lml = -1 * (.5 * site_cov.shape[0] * np.log(2 * pi) + np.sum(np.log(np.diag(chol)))
                    + .5 * (residual.T @ solve(sigma_marg, residual)).item())                  # :119-123
        if site_params is not None and rule.damping != 1.:
            site_mean, site_cov = _damp(rule, site_mean, site_cov, site_nat2, site_params, 0., 1e-10, 1e-10)
    return lml, site_mean, site_cov


def slep_update(rule, likelihood, y, post_mean, post_cov, site_params=None):
    """StatisticallyLinearisedEP.update, approximate_inference.py:180-210, matrix form."""
    y = np.reshape(np.asarray(y, dtype=float), (1, 1))
    cub, cub2 = cubature.make(rule.intmethod, rule.num_cub_pts), cubature.make_2d(rule.intmethod, rule.num_cub_pts)
    power = 1. if site_params is None else rule.power
    with np.errstate(all="ignore"):
        lml, _, _ = likelihood.moment_match(y, post_mean, post_cov, 1.0, cub)
        if site_params is None or power == 0:
            cav_mean, cav_cov = post_mean, post_cov
        else:
            cav_mean, cav_cov = compute_cavity(post_mean, post_cov, site_params[0], site_params[1], power)
        mu, S, C, omega = likelihood.statistical_linear_regression(cav_mean, cav_cov, cub2)
        residual = y - mu
        sigma = S + (power - 1) * C.T @ inv(cav_cov + 1e-10 * np.eye(2)) @ C
        om_sig_om = omega.T @ inv(sigma) @ omega
        om_sig_om = np.diag(np.diag(om_sig_om))
        osigo = inv(om_sig_om + 1e-10 * np.eye(2))
        site_mean = cav_mean + osigo @ omega.T @ inv(sigma) @ residual
        site_cov = -power * cav_cov + osigo
        if site_params is not None and rule.damping != 1.:
            site_mean, site_cov = _damp(rule, site_mean, site_cov, None, site_params, 1e-10, 1e-10, 1e-10)
    return lml, site_mean, site_cov


def vi_update(rule, likelihood, y, post_mean, post_cov, site_params=None):
    """VariationalInference.update, approximate_inference.py:302-323, matrix form."""
    cub = cubature.make(rule.intmethod, rule.num_cub_pts)
    eye = np.eye(2)
    with np.errstate(all="ignore"):
        _, dE_dm, dE_dv = likelihood.variational_expectation(y, post_mean, post_cov, cub)
        if site_params is None:
            site_cov = 0.5 * inv(ensure_positive_precision(-dE_dv) + 1e-10 * eye)
            site_mean = post_mean + site_cov @ dE_dm
        else:
            site_mean, site_cov = site_params
            dE_dv = -ensure_positive_precision(-dE_dv)
            lambda_t_2 = inv(site_cov + 1e-10 * eye)
            lambda_t_1 = lambda_t_2 @ site_mean
            lambda_t_1 = (1 - rule.damping) * lambda_t_1 + rule.damping * (dE_dm - 2 * dE_dv @ post_mean)
            lambda_t_2 = (1 - rule.damping) * lambda_t_2 + rule.damping * (-2 * dE_dv)
            site_cov = inv(lambda_t_2 + 1e-10 * eye)
            site_mean = site_cov @ lambda_t_1
        lml, _, _ = likelihood.moment_match(y, post_mean, post_cov, 1.0, cub)
    return lml, site_mean, site_cov


def site_update(rule, likelihood, y, post_mean, post_cov, site_params=None):
    """The rule's `update` in matrix form, chosen by the class of `rule` (oracle.approx_inf's class tree)."""
    from . import approx_inf as ai
    if isinstance(rule, ai.VariationalInference):
        return vi_update(rule, likelihood, y, post_mean, post_cov, site_params)
    if isinstance(rule, ai.StatisticallyLinearisedEP):
        return slep_update(rule, likelihood, y, post_mean, post_cov, site_params)
    if isinstance(rule, ai.ExtendedEP):
        return eep_update(rule, likelihood, y, post_mean, post_cov, site_params)
    return ep_update(rule, likelihood, y, post_mean, post_cov, site_params)


class SDEGPMulti(object):
    """sde_gp.py with matrix-valued sites (f = func_dim rows of H)."""
    def __init__(self, prior, likelihood, t, y, r=None, approx_inf=None):
        self.t, self.y, self.r, self.dt = input_admin(t, y, r)
        self.prior, self.likelihood, self.sites = prior, likelihood, approx_inf
        _, self.Pinf = prior.kernel_to_state_space()
        self.state_dim = self.Pinf.shape[0]
        self.func_dim = prior.measurement_model(self.r[0]).shape[0]

    def kalman_filter(self, y, dt, store=False, mask=None, site_params=None, r=None):
        """sde_gp.py:230-309."""
        Pinf, N, d, f = self.Pinf, dt.shape[0], self.state_dim, self.func_dim
        nlml = 0.0
        m, P = np.zeros([d, 1]), Pinf
        fm, fP = np.zeros([N, d, 1]), np.zeros([N, d, d])
        smean, scov = np.zeros([N, f, 1]), np.zeros([N, f, f])
        for n in range(N):
            y_n = y[n][..., None]
            A = self.prior.state_transition(dt[n])
            m_ = A @ m
            P_ = A @ (P - Pinf) @ A.T + Pinf
            H = self.prior.measurement_model(r[n])
            predict_mean, predict_cov = H @ m_, H @ P_ @ H.T
            if mask is not None:
                y_n = np.where(mask[n][..., None], predict_mean[:y_n.shape[0]], y_n)      # :286
            log_lik_n, site_mean, site_cov = site_update(self.sites, self.likelihood, y_n, predict_mean, predict_cov, None)
            if site_params is not None:
                site_mean, site_cov = site_params[0][n], site_params[1][n]                # :290
            S = predict_cov + site_cov
            HP = H @ P_
            K = solve(S, HP).T
            m_new, P_new = m_ + K @ (site_mean - predict_mean), P_ - K @ HP
            if mask is not None and np.any(mask[n]):
                m_new, P_new, log_lik_n = m_, P_, 0.0
            m, P = m_new, P_new
            nlml -= float(np.sum(log_lik_n))
            fm[n], fP[n], smean[n], scov[n] = m, P, site_mean, site_cov
        if store:
            return nlml, (fm, fP, (smean, scov))
        return nlml

    def rauch_tung_striebel_smoother(self, m_filtered, P_filtered, dt, store=False, return_full=False, y=None,
                                     site_params=None, r=None):
        """sde_gp.py:311-384."""
        Pinf, N, d, f = self.Pinf, dt.shape[0], self.state_dim, self.func_dim
        dt = np.concatenate([dt[1:], np.array([0.0])], axis=0)
        m, P = m_filtered[-1], P_filtered[-1]
        k = d if return_full else f
        sm, sc = np.zeros([N, k, 1]), np.zeros([N, k, k])
        if site_params is not None:
            new_mean, new_cov = np.zeros([N, f, 1]), np.zeros([N, f, f])
        for n in range(N - 1, -1, -1):
            A = self.prior.state_transition(dt[n])
            m_predicted = A @ m_filtered[n]
            tmp_gain_cov = A @ P_filtered[n]
            P_predicted = A @ (P_filtered[n] - Pinf) @ A.T + Pinf
            G_transpose = solve(P_predicted, tmp_gain_cov)
            m = m_filtered[n] + G_transpose.T @ (m - m_predicted)
            P = P_filtered[n] + G_transpose.T @ (P - P_predicted) @ G_transpose
            H = self.prior.measurement_model(r[n])
            if store:
                sm[n], sc[n] = (m, P) if return_full else (H @ m, H @ P @ H.T)
            if site_params is not None:
                _, mu, cov = site_update(self.sites, self.likelihood, y[n][..., None], H @ m, H @ P @ H.T,
                                       (site_params[0][n], site_params[1][n]))
                new_mean[n], new_cov[n] = mu, cov
        if site_params is not None:
            site_params = (new_mean, new_cov)
        return (site_params, sm, sc) if store else site_params

    def run(self):
        nlml, (fm, fP, sites) = self.kalman_filter(self.y, self.dt, True, None, self.sites.site_params, self.r)
        self.sites.site_params = self.rauch_tung_striebel_smoother(fm, fP, self.dt, False, False, self.y, sites, self.r)
        return nlml

    def predict_everywhere(self, y, r, dt, train_id, mask, site_params=None):
        f = self.func_dim
        site_params = self.sites.site_params if site_params is None else site_params
        if site_params is not None:
            site_mean = np.zeros([dt.shape[0], f, 1])
            site_cov = 1e5 * np.tile(np.eye(f), (dt.shape[0], 1, 1))                      # sde_gp.py:95
            site_mean[train_id] += site_params[0]
            site_cov[train_id] = site_params[1]
            site_params = (site_mean, site_cov)
        _, (fm, fP, site_params) = self.kalman_filter(y, dt, True, mask, site_params, r)
        _, pm, pc = self.rauch_tung_striebel_smoother(fm, fP, dt, True, False, None, None, r)
        return pm, pc, site_params

    def predict(self, t=None, r=None):
        (_, y, r_all, _, dt, train_id, test_id, mask) = test_input_admin(self.t, self.y, self.r, t, None, r)
        pm, pc, _ = self.predict_everywhere(y, r_all, dt, train_id, mask)
        return np.squeeze(pm[test_id]), np.squeeze(pc[test_id])

    def negative_log_predictive_density(self, t=None, y=None, r=None):
        (_, y_all, r_all, _, dt, train_id, test_id, mask) = test_input_admin(self.t, self.y, self.r, t, y, r)
        pm, pc, _ = self.predict_everywhere(y_all, r_all, dt, train_id, mask)
        cub = cubature.gauss_hermite(20)                                                   # sde_gp.py:139-142: cubature_func=None
        lpd = [self.likelihood.moment_match(y_all[i], pm[i], pc[i], 1, cub)[0] for i in test_id]
        return -np.mean(lpd)
