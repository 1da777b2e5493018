"""ORACLE (test infrastructure, not product): likelihood moment primitives for scalar sites (f = 1).

NumPy fp64 restatement of /root/reference/kalmanjax/likelihoods.py and the helpers it uses in
utils.py.  Every function is array-generic: `y`, `m`, `v` may be scalars or arrays of any common
shape (one entry per time step); cubature adds a trailing axis of Q points.  1x1 "matrix" operations
of the reference become scalar ones, keeping the reference's order of operations and its NaN
behaviour (a Cholesky of a negative 1x1 matrix is NaN: utils.py:25-38).
"""
import numpy as np
from scipy.special import erf, erfc, gammaln

pi = 3.141592653589793  # likelihoods.py:7


def chol1(v):
    """cho_factor of a 1x1 matrix (utils.py:29,37): sqrt, NaN when negative."""
    with np.errstate(invalid="ignore"):
        return np.sqrt(np.asarray(v, dtype=float))


def inv1(v):
    """utils.py:33-38 for a 1x1 matrix: cho_factor then cho_solve against 1 -> (1/L)/L."""
    L = chol1(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (1.0 / L) / L


def ensure_positive_precision1(k):
    """utils.py:15-22 for a 1x1 matrix: a negative entry becomes 0.01 (NaN stays NaN)."""
    k = np.asarray(k, dtype=float)
    return np.where(k < 0, 1e-2, k)


def logphi(z):
    """utils.py:200-216."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        lp = np.log(erfc(-z / np.sqrt(2.0)) / 2.0)
        dlp = np.exp(-z * z / 2.0 - lp) / np.sqrt(2.0 * pi)
    return lp, dlp


def softplus(x):
    # utils.py:67-69
    return np.log(1 + np.exp(-np.abs(x))) + np.maximum(x, 0)


def sigmoid(x):
    # utils.py:72-73
    return np.exp(x) / (np.exp(x) + 1.)


class Likelihood(object):
    """Base class: the generic cubature versions (likelihoods.py:122-185, 203-242, 275-321)."""
    name = "base"
    hyp = None  # already-transformed likelihood hyperparameter (positive), or None

    # -- to be provided by subclasses (likelihoods.py:41-47)
    def evaluate_likelihood(self, y, f):
        raise NotImplementedError

    def evaluate_log_likelihood(self, y, f):
        raise NotImplementedError

    def conditional_moments(self, f):
        raise NotImplementedError

    def moment_match_cubature(self, y, cav_mean, cav_cov, power, cub):
        """likelihoods.py:122-185 (the live definition; the one at 49-120 is shadowed)."""
        x, w = cub
        y, cav_mean, cav_cov = (np.asarray(a, dtype=float)[..., None] for a in (y, cav_mean, cav_cov))
        with np.errstate(all="ignore"):
            cav_cho = chol1(cav_cov)                                  # :145
            sigma_points = cav_cho * x + cav_mean                     # :147
            wle = w * self.evaluate_likelihood(y, sigma_points) ** power   # :149
            Z = np.sum(wle, axis=-1)                                  # :154-156
            lZ = np.log(np.maximum(Z, 1e-8))                          # :157
            Zinv = 1.0 / np.maximum(Z, 1e-8)                          # :158
            invC = inv1(cav_cov)                                      # :11-12 (per sigma point in the ref)
            dZ = np.sum(invC * (sigma_points - cav_mean) * wle, axis=-1)       # :163-166
            dlZ = Zinv * dZ                                           # :168
            d = sigma_points - cav_mean
            d2Z = np.sum((invC * d * d * invC - invC) * wle, axis=-1)  # :15-17, :173-176
            d2lZ = -dlZ * dlZ + Zinv * d2Z                            # :181
            id2lZ = inv1(ensure_positive_precision1(-d2lZ) - 1e-10)   # :182  (minus jitter, as written)
            site_mean = cav_mean[..., 0] + id2lZ * dlZ                # :183
            site_cov = power * (-cav_cov[..., 0] + id2lZ)             # :184
        return lZ, site_mean, site_cov

    def moment_match(self, y, m, v, power, cub):
        return self.moment_match_cubature(y, m, v, power, cub)       # :187-192

    def statistical_linear_regression(self, cav_mean, cav_cov, cub):
        """likelihoods.py:203-242 with f = 1 -> (mu, S, C, omega)."""
        x, w = cub
        cav_mean, cav_cov = (np.asarray(a, dtype=float)[..., None] for a in (cav_mean, cav_cov))
        with np.errstate(all="ignore"):
            sigma_points = chol1(cav_cov) * x + cav_mean              # :214
            E, V = self.conditional_moments(sigma_points)             # :215
            mu = np.sum(w * E, axis=-1, keepdims=True)                # :219-221
            S = np.sum(w * ((E - mu) * (E - mu) + V), axis=-1)        # :226-228
            C = np.sum(w * (sigma_points - cav_mean) * (E - mu), axis=-1)      # :232-234
            omega = np.sum(w * E * (inv1(cav_cov) * (sigma_points - cav_mean)), axis=-1)  # :238-240
        return mu[..., 0], S, C, omega

    def variational_expectation(self, y, post_mean, post_cov, cub):
        """likelihoods.py:275-321 with f = 1 -> (E[log p], dE/dm, dE/dv)."""
        x, w = cub
        y, post_mean, post_cov = (np.asarray(a, dtype=float)[..., None] for a in (y, post_mean, post_cov))
        with np.errstate(all="ignore"):
            sigma_points = chol1(post_cov) * x + post_mean            # :295
            wll = w * self.evaluate_log_likelihood(y, sigma_points)   # :297
            exp_log_lik = np.sum(wll, axis=-1)                        # :301-303
            invv = post_cov ** -1                                     # :307 (plain reciprocal of the diagonal)
            dE_dm = np.sum(invv * (sigma_points - post_mean) * wll, axis=-1)   # :308-311
            dE_dv = np.sum((0.5 * (invv ** 2 * (sigma_points - post_mean) ** 2) - 0.5 * invv) * wll, axis=-1)
        return exp_log_lik, dE_dm, dE_dv

    def analytical_linearisation(self, m):
        """Jacobians (Jf, Jsigma) of h(f, s) = E[y|f] + sqrt(Var[y|f]) s at (m, 0)."""
        raise NotImplementedError


class Gaussian(Likelihood):
    """likelihoods.py:331-404."""
    name = "Gaussian"

    def __init__(self, variance=0.1):
        self.hyp = float(variance)

    def evaluate_likelihood(self, y, f):
        return (2 * pi * self.hyp) ** -0.5 * np.exp(-0.5 * (y - f) ** 2 / self.hyp)       # :357

    def evaluate_log_likelihood(self, y, f):
        return -0.5 * np.log(2 * pi * self.hyp) - 0.5 * (y - f) ** 2 / self.hyp           # :371

    def conditional_moments(self, f):
        return f, self.hyp                                                               # :380-381

    def moment_match(self, y, m, v, power, cub):
        # likelihoods.py:385-404 -> utils.py:219-244 (power is ignored, site is (y, sigma^2))
        y, m, v = (np.asarray(a, dtype=float) for a in (y, m, v))
        with np.errstate(all="ignore"):
            lZ = (- (y - m) ** 2 / (self.hyp + v) / 2
                  - np.log(np.maximum(2 * pi * (self.hyp + v), 1e-10)) / 2)
        # site = (y, sigma^2) whatever the cavity is, even a NaN one (utils.py:241-244)
        return lZ, np.broadcast_to(y, lZ.shape).copy(), np.full(lZ.shape, self.hyp)

    def analytical_linearisation(self, m):
        # base class jacrev of f + chol(hyp) s (likelihoods.py:250-273): exact derivatives of a linear map
        m = np.asarray(m, dtype=float)
        return np.ones(m.shape), np.full(m.shape, np.sqrt(self.hyp))


class Bernoulli(Likelihood):
    """likelihoods.py:407-518."""
    name = "Bernoulli"

    def __init__(self, link="probit"):
        if link not in ("logit", "probit"):
            raise NotImplementedError("link function not implemented")
        self.link = link

    def link_fn(self, f):
        with np.errstate(all="ignore"):
            if self.link == "logit":
                return 1 / (1 + np.exp(-f))                                               # :423
            jitter = 1e-10
            return 0.5 * (1.0 + erf(f / np.sqrt(2.0))) * (1 - 2 * jitter) + jitter        # :428-429

    def dlink_fn(self, f):
        with np.errstate(all="ignore"):
            if self.link == "logit":
                return np.exp(f) / (1 + np.exp(f)) ** 2                                   # :424
            # :430 is grad(link_fn): d/df of the jittered probit
            return (1 - 2e-10) * np.exp(-0.5 * f * f) / np.sqrt(2.0 * pi)

    def evaluate_likelihood(self, y, f):
        return np.where(np.equal(y, 1), self.link_fn(f), 1 - self.link_fn(f))           # :445

    def evaluate_log_likelihood(self, y, f):
        with np.errstate(all="ignore"):
            return np.log(self.evaluate_likelihood(y, f))                                # :456

    def conditional_moments(self, f):
        p = self.link_fn(f)
        return p, p - (p ** 2)                                                           # :465

    def moment_match(self, y, m, v, power, cub):
        # likelihoods.py:467-505
        y = np.sign(np.asarray(y, dtype=float))                                          # :488
        if power == 1 and self.link == "probit":
            m, v = np.asarray(m, dtype=float), np.asarray(v, dtype=float)
            with np.errstate(all="ignore"):
                y = np.sign(y - 0.01)                                                    # :490
                z = m / np.sqrt(1.0 + v)
                z = z * y
                lZ, dlp = logphi(z)
                dlZ = y * dlp / np.sqrt(1.0 + v)
                d2lZ = -dlp * (z + dlp) / (1.0 + v)
                site_mean = m - dlZ / d2lZ
                site_var = - (v + 1 / d2lZ)
            return lZ, site_mean, site_var
        return self.moment_match_cubature(y, m, v, power, cub)                           # :505

    def analytical_linearisation(self, m):
        # likelihoods.py:507-518 evaluated at sigma = 0 (the only call site: approximate_inference.py:112)
        m = np.asarray(m, dtype=float)
        with np.errstate(all="ignore"):
            p = self.link_fn(m)
            Jf = self.dlink_fn(m) + (0.5 * (p * (1 - p)) ** -0.5 * self.dlink_fn(m) * (1 - 2 * p)) * 0.0
            Jsigma = (p * (1 - p)) ** 0.5
        return Jf, Jsigma


class Probit(Bernoulli):
    def __init__(self):
        super().__init__("probit")


class Logit(Bernoulli):
    def __init__(self):
        super().__init__("logit")


class Poisson(Likelihood):
    """likelihoods.py:551-644."""
    name = "Poisson"

    def __init__(self, link="exp"):
        if link not in ("exp", "logistic"):
            raise NotImplementedError("link function not implemented")
        self.link = link

    def link_fn(self, f):
        with np.errstate(all="ignore"):
            return np.exp(f) if self.link == "exp" else softplus(f)                      # :571-576

    def dlink_fn(self, f):
        with np.errstate(all="ignore"):
            return np.exp(f) if self.link == "exp" else sigmoid(f)

    def evaluate_likelihood(self, y, f):
        with np.errstate(all="ignore"):
            mu = self.link_fn(f)
            return mu ** y * np.exp(-mu) / np.exp(gammaln(y + 1))                        # :595-596

    def evaluate_log_likelihood(self, y, f):
        with np.errstate(all="ignore"):
            mu = self.link_fn(f)
            return y * np.log(mu) - mu - gammaln(y + 1)                                  # :612-613

    def conditional_moments(self, f):
        # :625-632. HEAD returns the variance as a [1,1,Q] array which breaks the SLR family downstream
        # (see tests/golden/make_golden.py:poisson_slr_patch); the scalar semantics are E = Var = link(f).
        mu = self.link_fn(f)
        return mu, mu

    def analytical_linearisation(self, m):
        # likelihoods.py:636-644 at sigma = 0. NB: Jf uses link_fn (not dlink_fn) exactly as written (:642).
        m = np.asarray(m, dtype=float)
        with np.errstate(all="ignore"):
            mu = self.link_fn(m)
            Jf = mu + 0.5 * mu ** -0.5 * self.dlink_fn(m) * 0.0
            Jsigma = mu ** 0.5
        return Jf, Jsigma
