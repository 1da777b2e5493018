"""ORACLE (test infrastructure, not product): per-step site-update rules for scalar sites (f = 1).

NumPy fp64 restatement of /root/reference/kalmanjax/approximate_inference.py.  Array-generic like
oracle/likelihoods.py: one call can update the sites of many time steps at once (the reference
calls `update` once per step from sde_gp.py:287 and sde_gp.py:375).
`site_params=None` selects the "initialise from the predictive" branches used by the filter.
"""
import numpy as np

from . import cubature
from .likelihoods import inv1, ensure_positive_precision1

pi = 3.141592653589793


def compute_cavity(post_mean, post_cov, site_mean, site_cov, power):
    """approximate_inference.py:8-15."""
    with np.errstate(all="ignore"):
        post_precision, site_precision = inv1(post_cov), inv1(site_cov)
        cav_cov = inv1(post_precision - power * site_precision)
        cav_mean = cav_cov * (post_precision * post_mean - power * site_precision * site_mean)
    return cav_mean, cav_cov


class ApproxInf(object):
    """approximate_inference.py:18-37."""
    def __init__(self, site_params=None, intmethod="GH", num_cub_pts=20):
        self.site_params = site_params
        self.intmethod, self.num_cub_pts = intmethod, num_cub_pts
        self.cub = cubature.make(intmethod, num_cub_pts)

    def update(self, likelihood, y, post_mean, post_cov, site_params=None):
        raise NotImplementedError


class ExpectationPropagation(ApproxInf):
    """approximate_inference.py:40-73."""
    def __init__(self, site_params=None, damping=1.0, power=1.0, intmethod="GH", num_cub_pts=20):
        self.damping, self.power = damping, power
        super().__init__(site_params, intmethod, num_cub_pts)
        self.name = "Expectation Propagation (EP)"

    def update(self, likelihood, y, post_mean, post_cov, site_params=None):
        if site_params is None:
            return likelihood.moment_match(y, post_mean, post_cov, 1.0, self.cub)                # :57
        site_mean_prev, site_cov_prev = site_params
        cav_mean, cav_cov = compute_cavity(post_mean, post_cov, site_mean_prev, site_cov_prev, self.power)
        lml, site_mean, site_cov = likelihood.moment_match(y, cav_mean, cav_cov, self.power, self.cub)   # :66
        if self.damping != 1.:
            with np.errstate(all="ignore"):
                site_nat2, site_nat2_prev = inv1(site_cov), inv1(site_cov_prev)                  # :69
                site_nat1, site_nat1_prev = site_nat2 * site_mean, site_nat2_prev * site_mean_prev
                site_cov = inv1((1. - self.damping) * site_nat2_prev + self.damping * site_nat2)
                site_mean = site_cov * ((1. - self.damping) * site_nat1_prev + self.damping * site_nat1)
        return lml, site_mean, site_cov


EP = PEP = PowerExpectationPropagation = ExpectationPropagation


class ExtendedEP(ApproxInf):
    """approximate_inference.py:88-133."""
    def __init__(self, site_params=None, power=1.0, damping=1.0):
        self.power, self.damping = power, damping
        super().__init__(site_params)
        self.name = "Extended Expectation Propagation (EEP)"

    def update(self, likelihood, y, post_mean, post_cov, site_params=None):
        y = np.asarray(y, dtype=float)
        power = 1. if site_params is None else self.power                                        # :104
        if (site_params is None) or (power == 0):
            cav_mean, cav_cov = np.asarray(post_mean, dtype=float), np.asarray(post_cov, dtype=float)
        else:
            cav_mean, cav_cov = compute_cavity(post_mean, post_cov, site_params[0], site_params[1], power)
        with np.errstate(all="ignore"):
            Jf, Jsigma = likelihood.analytical_linearisation(cav_mean)                           # :112
            likelihood_expectation, _ = likelihood.conditional_moments(cav_mean)                 # :114
            residual = y - likelihood_expectation
            sigma = Jsigma * Jsigma + power * Jf * cav_cov * Jf                                  # :116
            site_nat2 = Jf * inv1(Jsigma * Jsigma) * Jf                                          # :117
            site_cov = inv1(site_nat2 + 1e-10)                                                   # :118
            site_mean = cav_mean + (site_cov + power * cav_cov) * Jf * inv1(sigma) * residual    # :119
            sigma_marg_lik = Jsigma * Jsigma + Jf * cav_cov * Jf                                 # :121
            chol_sigma = np.sqrt(sigma_marg_lik)                                                 # :122
            log_marg_lik = -1 * (.5 * np.log(2 * pi) + np.log(chol_sigma)
                                 + .5 * (residual * ((residual / chol_sigma) / chol_sigma)))     # :123-126
            if (site_params is not None) and (self.damping != 1.):
                site_mean_prev, site_cov_prev = site_params
                site_nat2_prev = inv1(site_cov_prev + 1e-10)                                     # :129
                site_nat1, site_nat1_prev = site_nat2 * site_mean, site_nat2_prev * site_mean_prev
                site_cov = inv1((1. - self.damping) * site_nat2_prev + self.damping * site_nat2 + 1e-10)
                site_mean = site_cov * ((1. - self.damping) * site_nat1_prev + self.damping * site_nat1)
        return log_marg_lik, site_mean, site_cov


EEP = ExtendedEP


class ExtendedKalmanSmoother(ExtendedEP):
    """approximate_inference.py:140-146."""
    def __init__(self, site_params=None, damping=1.0):
        super().__init__(site_params=site_params, power=0, damping=damping)
        self.name = "Extended Kalman Smoother (EKS)"


EKS = EKF = ExtendedKalmanFilter = ExtendedKalmanSmoother


class StatisticallyLinearisedEP(ApproxInf):
    """approximate_inference.py:166-210."""
    def __init__(self, site_params=None, power=1.0, damping=1.0, intmethod="GH", num_cub_pts=20):
        self.power, self.damping = power, damping
        super().__init__(site_params, intmethod, num_cub_pts)
        self.name = "Statistically Linearised Expectation Propagation (SLEP)"

    def update(self, likelihood, y, post_mean, post_cov, site_params=None):
        y = np.asarray(y, dtype=float)
        power = 1. if site_params is None else self.power                                        # :185
        log_marg_lik, _, _ = likelihood.moment_match(y, post_mean, post_cov, 1.0, self.cub)      # :186
        if (site_params is None) or (power == 0):
            cav_mean, cav_cov = np.asarray(post_mean, dtype=float), np.asarray(post_cov, dtype=float)
        else:
            cav_mean, cav_cov = compute_cavity(post_mean, post_cov, site_params[0], site_params[1], power)
        mu, S, C, omega = likelihood.statistical_linear_regression(cav_mean, cav_cov, self.cub)  # :194
        with np.errstate(all="ignore"):
            residual = y - mu
            sigma = S + (power - 1) * C * inv1(cav_cov + 1e-10) * C                              # :197
            om_sig_om = omega * inv1(sigma) * omega                                              # :198-199
            osigo = inv1(om_sig_om + 1e-10)                                                      # :200
            site_mean = cav_mean + osigo * omega * inv1(sigma) * residual                        # :201
            site_cov = -power * cav_cov + osigo                                                  # :202
            if (site_params is not None) and (self.damping != 1.):
                site_mean_prev, site_cov_prev = site_params
                jitter = 1e-10
                site_nat2, site_nat2_prev = inv1(site_cov + jitter), inv1(site_cov_prev + jitter)
                site_nat1, site_nat1_prev = site_nat2 * site_mean, site_nat2_prev * site_mean_prev
                site_cov = inv1((1. - self.damping) * site_nat2_prev + self.damping * site_nat2 + jitter)
                site_mean = site_cov * ((1. - self.damping) * site_nat1_prev + self.damping * site_nat1)
        return log_marg_lik, site_mean, site_cov


SLEP = PL = PosteriorLinearisation = StatisticallyLinearisedEP


class GaussHermiteEP(StatisticallyLinearisedEP):
    def __init__(self, site_params=None, power=1, damping=1, num_cub_pts=20):
        super().__init__(site_params, power, damping, "GH", num_cub_pts)
        self.name = "Gauss-Hermite Expectation Propagation (GHEP)"


GHEP = GaussHermiteEP


class GaussHermiteKalmanSmoother(GaussHermiteEP):
    def __init__(self, site_params=None, damping=1, num_cub_pts=20):
        super().__init__(site_params, 0, damping, num_cub_pts)
        self.name = "Gauss-Hermite Kalman Smoother (GHKS)"


GHKS = GHKF = GaussHermiteKalmanFilter = GaussHermiteKalmanSmoother


class UnscentedEP(StatisticallyLinearisedEP):
    def __init__(self, site_params=None, power=1, damping=1):
        super().__init__(site_params, power, damping, "UT")
        self.name = "Unscented Expectation Propagation (UEP)"


UEP = UnscentedEP


class UnscentedKalmanSmoother(UnscentedEP):
    def __init__(self, site_params=None, damping=1):
        super().__init__(site_params, 0, damping)
        self.name = "Unscented Kalman Smoother (UKS)"


UKS = UKF = UnscentedKalmanFilter = UnscentedKalmanSmoother


class VariationalInference(ApproxInf):
    """approximate_inference.py:289-323."""
    def __init__(self, site_params=None, damping=1., intmethod="GH", num_cub_pts=20):
        self.damping = damping
        super().__init__(site_params, intmethod, num_cub_pts)
        self.name = "Variational Inference (VI)"

    def update(self, likelihood, y, post_mean, post_cov, site_params=None):
        post_mean = np.asarray(post_mean, dtype=float)
        _, dE_dm, dE_dv = likelihood.variational_expectation(y, post_mean, post_cov, self.cub)   # :307/312
        with np.errstate(all="ignore"):
            if site_params is None:
                site_cov = 0.5 * inv1(ensure_positive_precision1(-dE_dv) + 1e-10)                # :309
                site_mean = post_mean + site_cov * dE_dm                                         # :310
            else:
                site_mean, site_cov = site_params
                dE_dv = -ensure_positive_precision1(-dE_dv)                                      # :315
                lambda_t_2 = inv1(site_cov + 1e-10)                                              # :316
                lambda_t_1 = lambda_t_2 * site_mean
                lambda_t_1 = (1 - self.damping) * lambda_t_1 + self.damping * (dE_dm - 2 * dE_dv * post_mean)
                lambda_t_2 = (1 - self.damping) * lambda_t_2 + self.damping * (-2 * dE_dv)
                site_cov = inv1(lambda_t_2 + 1e-10)                                              # :320
                site_mean = site_cov * lambda_t_1
        log_marg_lik, _, _ = likelihood.moment_match(y, post_mean, post_cov, 1.0, self.cub)      # :322
        return log_marg_lik, site_mean, site_cov


VI = CVI = VariationalInference
