"""Approximate-inference (site update) rules — host-side mirror of
/root/reference/kalmanjax/approximate_inference.py.  Same class names, aliases and constructor arguments;
each object carries the rule's parameters (kind, power, damping, cubature rule) which select the update
that the GPU kernels apply per time step (csrc/kjx_sites.cuh).  `update()` evaluates the rule for a whole
batch of steps on the GPU through kjx_site_update (include/kjx.h).
"""
from . import _lib
from .utils import gauss_hermite, symmetric_cubature_third_order, symmetric_cubature_fifth_order


class ApproxInf(object):
    """approximate_inference.py:18-37."""
    kjx_code = None
    power = 1.0
    damping = 1.0

    def __init__(self, site_params=None, intmethod='GH', num_cub_pts=20):
        self.site_params = site_params
        self.intmethod = intmethod
        if intmethod == 'GH':
            self.cubature_func = lambda dim: gauss_hermite(dim, num_cub_pts)
        elif intmethod == 'UT3':
            self.cubature_func = lambda dim: symmetric_cubature_third_order(dim)
        elif (intmethod == 'UT5') or (intmethod == 'UT'):
            self.cubature_func = lambda dim: symmetric_cubature_fifth_order(dim)
        else:
            raise NotImplementedError('integration method not recognised')

    def update(self, likelihood, y, post_mean, post_cov, hyp=None, site_params=None):
        """(log_lik, site_mean, site_cov) for every entry of the batch; site_params=None is the
        "initialise from the predictive" branch the filter uses (sde_gp.py:287)."""
        if self.kjx_code is None:
            raise NotImplementedError('the update function for this approximate inference method is not implemented')
        from .sde_gp import site_update
        return site_update(likelihood, self, y, post_mean, post_cov, hyp, site_params)


class ExpectationPropagation(ApproxInf):
    """approximate_inference.py:40-73."""
    kjx_code = _lib.INF_EP

    def __init__(self, site_params=None, damping=1.0, power=1.0, intmethod='GH', num_cub_pts=20):
        self.damping = damping
        self.power = power
        super().__init__(site_params=site_params, intmethod=intmethod, num_cub_pts=num_cub_pts)
        self.name = 'Expectation Propagation (EP)'


class PowerExpectationPropagation(ExpectationPropagation):
    pass


class EP(ExpectationPropagation):
    pass


class PEP(ExpectationPropagation):
    pass


class ExtendedEP(ApproxInf):
    """approximate_inference.py:88-133."""
    kjx_code = _lib.INF_EEP

    def __init__(self, site_params=None, power=1.0, damping=1.0):
        self.power = power
        self.damping = damping
        super().__init__(site_params=site_params)
        self.name = 'Extended Expectation Propagation (EEP)'


class EEP(ExtendedEP):
    pass


class ExtendedKalmanSmoother(ExtendedEP):
    def __init__(self, site_params=None, damping=1.0):
        super().__init__(site_params=site_params, power=0, damping=damping)
        self.name = 'Extended Kalman Smoother (EKS)'


class EKS(ExtendedKalmanSmoother):
    pass


class ExtendedKalmanFilter(ExtendedKalmanSmoother):
    pass


class EKF(ExtendedKalmanFilter):
    pass


class StatisticallyLinearisedEP(ApproxInf):
    """approximate_inference.py:166-210."""
    kjx_code = _lib.INF_SLEP

    def __init__(self, site_params=None, power=1.0, damping=1.0, intmethod='GH', num_cub_pts=20):
        self.power = power
        self.damping = damping
        super().__init__(site_params=site_params, intmethod=intmethod, num_cub_pts=num_cub_pts)
        self.name = 'Statistically Linearised Expectation Propagation (SLEP)'


class SLEP(StatisticallyLinearisedEP):
    pass


class GaussHermiteEP(StatisticallyLinearisedEP):
    def __init__(self, site_params=None, power=1, damping=1, num_cub_pts=20):
        super().__init__(site_params=site_params, power=power, damping=damping, intmethod='GH', num_cub_pts=num_cub_pts)
        self.name = 'Gauss-Hermite Expectation Propagation (GHEP)'


class GHEP(GaussHermiteEP):
    pass


class GaussHermiteKalmanSmoother(GaussHermiteEP):
    def __init__(self, site_params=None, damping=1, num_cub_pts=20):
        super().__init__(site_params=site_params, power=0, damping=damping, num_cub_pts=num_cub_pts)
        self.name = 'Gauss-Hermite Kalman Smoother (GHKS)'


class GHKS(GaussHermiteKalmanSmoother):
    pass


class UnscentedEP(StatisticallyLinearisedEP):
    def __init__(self, site_params=None, power=1, damping=1):
        super().__init__(site_params=site_params, power=power, damping=damping, intmethod='UT')
        self.name = 'Unscented Expectation Propagation (UEP)'


class UEP(UnscentedEP):
    pass


class UnscentedKalmanSmoother(UnscentedEP):
    def __init__(self, site_params=None, damping=1):
        super().__init__(site_params=site_params, power=0, damping=damping)
        self.name = 'Unscented Kalman Smoother (UKS)'


class UKS(UnscentedKalmanSmoother):
    pass


class PosteriorLinearisation(StatisticallyLinearisedEP):
    pass


class PL(PosteriorLinearisation):
    pass


class GaussHermiteKalmanFilter(GaussHermiteKalmanSmoother):
    pass


class GHKF(GaussHermiteKalmanFilter):
    pass


class UnscentedKalmanFilter(UnscentedKalmanSmoother):
    pass


class UKF(UnscentedKalmanFilter):
    pass


class VariationalInference(ApproxInf):
    """approximate_inference.py:289-323."""
    kjx_code = _lib.INF_VI

    def __init__(self, site_params=None, damping=1., intmethod='GH', num_cub_pts=20):
        self.damping = damping
        super().__init__(site_params=site_params, intmethod=intmethod, num_cub_pts=num_cub_pts)
        self.name = 'Variational Inference (VI)'


class VI(VariationalInference):
    pass


class CVI(VariationalInference):
    pass
