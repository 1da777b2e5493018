"""Approximate-inference (site update) rules — host-side mirror of
/root/reference/kalmanjax/approximate_inference.py.  Same class names, aliases and constructor arguments;
each object is a DESCRIPTOR: it carries the rule's parameters (kind, power, damping, cubature rule) which select the
update that the GPU kernels apply per time step (csrc/kjx_sites.cuh, csrc/kjx_multi_kernels.cuh).  `update()`
evaluates the rule for a whole batch of steps on the GPU through kjx_site_update (include/kjx.h).

The reference spells its 25 further class names out as a chain of subclasses; here the four rules are classes and
every other name is generated from one table (`_VARIANTS`): which rule it specialises, what it pins, what its
constructor accepts.
"""
from . import _lib
from .utils import gauss_hermite, symmetric_cubature_third_order, symmetric_cubature_fifth_order

_CUBATURE = {                      # approximate_inference.py:25-34
    'GH': lambda n: (lambda dim: gauss_hermite(dim, n)),
    'UT3': lambda n: (lambda dim: symmetric_cubature_third_order(dim)),
    'UT5': lambda n: (lambda dim: symmetric_cubature_fifth_order(dim)),
    'UT': lambda n: (lambda dim: symmetric_cubature_fifth_order(dim)),
}


class ApproxInf(object):
    """approximate_inference.py:18-37: the site parameters and the cubature rule; subclasses name the update."""
    kjx_code = None                # kjx_method of include/kjx.h
    display = None
    power = 1.0
    damping = 1.0

    def __init__(self, site_params=None, intmethod='GH', num_cub_pts=20):
        if intmethod not in _CUBATURE:
            raise NotImplementedError('integration method not recognised')
        self.site_params = site_params
        self.intmethod = intmethod
        self.cubature_func = _CUBATURE[intmethod](num_cub_pts)
        if self.display is not None:
            self.name = self.display

    def update(self, likelihood, y, post_mean, post_cov, hyp=None, site_params=None):
        """(log_lik, site_mean, site_cov) for every entry of the batch; site_params=None is the
        "initialise from the predictive" branch the filter uses (sde_gp.py:287)."""
        if self.kjx_code is None:
            raise NotImplementedError('the update function for this approximate inference method is not implemented')
        from .sde_gp import site_update
        return site_update(likelihood, self, y, post_mean, post_cov, hyp, site_params)


class ExpectationPropagation(ApproxInf):
    """approximate_inference.py:40-73 (power EP with damping; moment matching by cubature)."""
    kjx_code = _lib.INF_EP
    display = 'Expectation Propagation (EP)'

    def __init__(self, site_params=None, damping=1.0, power=1.0, intmethod='GH', num_cub_pts=20):
        self.damping, self.power = damping, power
        super().__init__(site_params, intmethod, num_cub_pts)


class ExtendedEP(ApproxInf):
    """approximate_inference.py:88-133 (analytical linearisation at the cavity mean; power 0 is the EKS)."""
    kjx_code = _lib.INF_EEP
    display = 'Extended Expectation Propagation (EEP)'

    def __init__(self, site_params=None, power=1.0, damping=1.0):
        self.power, self.damping = power, damping
        super().__init__(site_params)


class StatisticallyLinearisedEP(ApproxInf):
    """approximate_inference.py:166-210 (statistical linear regression in the cavity; power 0 is posterior linearisation)."""
    kjx_code = _lib.INF_SLEP
    display = 'Statistically Linearised Expectation Propagation (SLEP)'

    def __init__(self, site_params=None, power=1.0, damping=1.0, intmethod='GH', num_cub_pts=20):
        self.power, self.damping = power, damping
        super().__init__(site_params, intmethod, num_cub_pts)


class VariationalInference(ApproxInf):
    """approximate_inference.py:289-323 (conjugate-computation VI, natural-gradient step of size `damping`)."""
    kjx_code = _lib.INF_VI
    display = 'Variational Inference (VI)'

    def __init__(self, site_params=None, damping=1., intmethod='GH', num_cub_pts=20):
        self.damping = damping
        super().__init__(site_params, intmethod, num_cub_pts)


def _variant(cls_name, base, display=None, accepts=None, **pinned):
    """Subclass `base` under the reference's name `cls_name`.  accepts: the constructor's arguments after site_params, in
    the reference's order (None: the base's own constructor, i.e. a plain alias); pinned: what the variant fixes."""
    body = {'__doc__': '%s: %s%s' % (cls_name, base.__name__, ''.join(', %s=%r' % kv for kv in sorted(pinned.items())))}
    if display is not None:
        body['display'] = display
    if accepts is not None:
        def __init__(self, site_params=None, *args, **kw):
            if len(args) > len(accepts):
                raise TypeError('%s() takes at most %d arguments after site_params' % (cls_name, len(accepts)))
            kw.update(zip(accepts, args))
            unknown = sorted(set(kw) - set(accepts))
            if unknown:
                raise TypeError('%s() got an unexpected keyword argument %r' % (cls_name, unknown[0]))
            kw.update(pinned)
            base.__init__(self, site_params, **kw)
        body['__init__'] = __init__
    return type(cls_name, (base,), body)


# name, what it specialises, display name, constructor arguments, pinned arguments          (approximate_inference.py)
_VARIANTS = [
    ('PowerExpectationPropagation', 'ExpectationPropagation', None, None, {}),                              # :76
    ('EP', 'ExpectationPropagation', None, None, {}),                                                       # :80
    ('PEP', 'ExpectationPropagation', None, None, {}),                                                      # :84
    ('EEP', 'ExtendedEP', None, None, {}),                                                                  # :136
    ('ExtendedKalmanSmoother', 'ExtendedEP', 'Extended Kalman Smoother (EKS)', ('damping',), dict(power=0)),   # :140-147
    ('EKS', 'ExtendedKalmanSmoother', None, None, {}),
    ('ExtendedKalmanFilter', 'ExtendedKalmanSmoother', None, None, {}),                                     # :153-158
    ('EKF', 'ExtendedKalmanFilter', None, None, {}),
    ('SLEP', 'StatisticallyLinearisedEP', None, None, {}),                                                  # :213
    ('GaussHermiteEP', 'StatisticallyLinearisedEP', 'Gauss-Hermite Expectation Propagation (GHEP)',
     ('power', 'damping', 'num_cub_pts'), dict(intmethod='GH')),                                            # :217-220
    ('GHEP', 'GaussHermiteEP', None, None, {}),
    ('GaussHermiteKalmanSmoother', 'GaussHermiteEP', 'Gauss-Hermite Kalman Smoother (GHKS)',
     ('damping', 'num_cub_pts'), dict(power=0)),                                                            # :227-230
    ('GHKS', 'GaussHermiteKalmanSmoother', None, None, {}),
    ('UnscentedEP', 'StatisticallyLinearisedEP', 'Unscented Expectation Propagation (UEP)',
     ('power', 'damping'), dict(intmethod='UT')),                                                           # :237-240
    ('UEP', 'UnscentedEP', None, None, {}),
    ('UnscentedKalmanSmoother', 'UnscentedEP', 'Unscented Kalman Smoother (UKS)', ('damping',), dict(power=0)),   # :247-250
    ('UKS', 'UnscentedKalmanSmoother', None, None, {}),
    ('PosteriorLinearisation', 'StatisticallyLinearisedEP', None, None, {}),                                # :257
    ('PL', 'PosteriorLinearisation', None, None, {}),
    ('GaussHermiteKalmanFilter', 'GaussHermiteKalmanSmoother', None, None, {}),                             # :265-270
    ('GHKF', 'GaussHermiteKalmanFilter', None, None, {}),
    ('UnscentedKalmanFilter', 'UnscentedKalmanSmoother', None, None, {}),                                   # :277-282
    ('UKF', 'UnscentedKalmanFilter', None, None, {}),
    ('VI', 'VariationalInference', None, None, {}),                                                         # :326
    ('CVI', 'VariationalInference', None, None, {}),                                                        # :330
]
for _name, _base, _display, _accepts, _pinned in _VARIANTS:
    globals()[_name] = _variant(_name, globals()[_base], _display, _accepts, **_pinned)
del _name, _base, _display, _accepts, _pinned
