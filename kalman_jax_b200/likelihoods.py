"""Likelihood models — host-side mirror of /root/reference/kalmanjax/likelihoods.py for the likelihoods
named by BASELINE.json (Gaussian 331-404, Bernoulli probit/logit 407-548, Poisson 551-644).

The classes only *describe* the likelihood (kind + link + hyperparameter, same constructor arguments and
`hyp` convention as the reference); the moment primitives themselves (moment_match, statistical linear
regression, variational expectation, analytical linearisation) are evaluated on the GPU inside the
filter / smoother / site-update kernels (csrc/kjx_sites.cuh).
"""
import numpy as np

from . import _lib
from .utils import softplus, softplus_inv


class Likelihood(object):
    """likelihoods.py:20-47."""
    kjx_code = None

    def __init__(self, hyp=None):
        hyp = [] if hyp is None else hyp
        self.hyp = softplus_inv(np.array(hyp, dtype=np.float64))

    def lik_hyp(self, hyp=None):
        """already-transformed hyperparameter handed to the kernels (0 when the likelihood has none)."""
        h = softplus(self.hyp) if hyp is None else np.asarray(hyp, dtype=np.float64)
        return float(np.reshape(h, -1)[0]) if np.size(h) else 0.0

    def moment_match(self, y, m, v, hyp=None, power=1.0, cubature_func=None):
        """(log Z, site_mean, site_var) per entry of the batch, on the GPU (likelihoods.py:122-192): the likelihood is
        raised to `power` and the site variance scaled by it, exactly as the reference's moment_match does.
        cubature_func=None means 20-point Gauss-Hermite, like likelihoods.py:141-142."""
        from .sde_gp import site_update
        from .utils import gauss_hermite

        class _MomentMatch(object):          # no update rule: kjx_site_update evaluates moment_match itself
            kjx_code = _lib.INF_MOMENT_MATCH
            damping = 1.0
        rule = _MomentMatch()
        rule.power = float(power)
        rule.cubature_func = gauss_hermite if cubature_func is None else cubature_func
        return site_update(self, rule, y, m, v, hyp, None)


class Gaussian(Likelihood):
    """likelihoods.py:331-404."""
    kjx_code = _lib.LIK_GAUSSIAN

    def __init__(self, variance=0.1):
        super().__init__(hyp=variance)
        self.name = 'Gaussian'

    @property
    def variance(self):
        return softplus(self.hyp)


class Bernoulli(Likelihood):
    """likelihoods.py:407-518."""
    def __init__(self, link):
        super().__init__(hyp=None)
        if link == 'logit':
            self.kjx_code = _lib.LIK_BERNOULLI_LOGIT
        elif link == 'probit':
            self.kjx_code = _lib.LIK_BERNOULLI_PROBIT
        else:
            raise NotImplementedError('link function not implemented')
        self.link = link
        self.name = 'Bernoulli'


class Probit(Bernoulli):
    def __init__(self):
        super().__init__(link='probit')


class Erf(Probit):
    pass


class Logit(Bernoulli):
    def __init__(self):
        super().__init__(link='logit')


class Logistic(Logit):
    pass


class Poisson(Likelihood):
    """likelihoods.py:551-644."""
    def __init__(self, link='exp'):
        super().__init__(hyp=None)
        if link == 'exp':
            self.kjx_code = _lib.LIK_POISSON_EXP
        elif link == 'logistic':
            self.kjx_code = _lib.LIK_POISSON_LOGISTIC
        else:
            raise NotImplementedError('link function not implemented')
        self.link = link
        self.name = 'Poisson'


class HeteroscedasticNoise(Likelihood):
    """likelihoods.py:647-850: p(y | f1, f2) = N(y | f1, link(f2)^2) — two latent functions (func_dim = 2), used with an
    `Independent` prior and any of the update rules: (power-)EP and VI integrate f1 analytically and f2 with the
    one-dimensional cubature of the inference method, the SLR family (SLEP / GHEP / UEP / GHKS / UKS / PL) uses its
    two-dimensional rule, EEP / EKS the analytical linearisation.  All of it runs in csrc/kjx_multi_kernels.cuh."""
    def __init__(self, link='softplus'):
        super().__init__(hyp=None)
        if link == 'softplus':
            self.kjx_code = _lib.LIK_HETERO_SOFTPLUS
            self.link_fn = lambda mu: np.log(1 + np.exp(-np.abs(np.asarray(mu) - 0.5))) + np.maximum(np.asarray(mu) - 0.5, 0) + 1e-10
        elif link == 'exp':
            self.kjx_code = _lib.LIK_HETERO_EXP
            self.link_fn = lambda mu: np.exp(np.asarray(mu) - 0.5)
        else:
            raise NotImplementedError('link function not implemented')
        self.link = link
        self.func_dim = 2
        self.name = 'Heteroscedastic Noise'
