"""GP priors as LTI SDEs — host-side mirror of /root/reference/kalmanjax/priors.py for the priors named
by BASELINE.json (Matern-3/2, Matern-5/2, QuasiPeriodicMatern32, Sum, SpatioTemporalMatern52) and, from the model
zoo of SURVEY 8(f)-4, Matern-1/2 (Exponential), Matern-7/2, Cosine, Periodic, QuasiPeriodicMatern12,
SubbandMatern12, SubbandMatern32, SpatialMatern52 and SpatialMatern32.

Same class names, constructor arguments and `hyp` convention (softplus-inverse of the positive
hyperparameters) as the reference.  What is computed here on the host is only what the reference
computes once per call (`kernel_to_state_space`: Pinf, H — tiny d x d NumPy work); the per-step
discretisation A(dt) runs on the GPU from the `blocks()` description (include/kjx.h kjx_block_t).
`state_transition()` is kept as a NumPy convenience with the reference's closed forms.
"""
import numpy as np
import scipy.linalg as sl

from . import _lib
from .utils import softplus, softplus_inv, softplus_list


class Prior(object):
    """priors.py:8-47."""
    def __init__(self, hyp=None):
        self.hyp = softplus_inv(np.array(hyp, dtype=np.float64))

    def _theta(self, hyperparams):
        return softplus(self.hyp) if hyperparams is None else np.asarray(hyperparams, dtype=np.float64)


class Matern32(Prior):
    """priors.py:111-175."""
    state_dim = 2

    def __init__(self, variance=1.0, lengthscale=1.0):
        super().__init__(hyp=[variance, lengthscale])
        self.name = 'Matern-3/2'

    @property
    def variance(self):
        return softplus(self.hyp[0])

    @property
    def lengthscale(self):
        return softplus(self.hyp[1])

    def kernel_to_state_space(self, hyperparams=None):
        var, ell = self._theta(hyperparams)
        lam = 3.0 ** 0.5 / ell
        F = np.array([[0.0, 1.0], [-lam ** 2, -2 * lam]])
        L = np.array([[0], [1]])
        Qc = np.array([[12.0 * 3.0 ** 0.5 / ell ** 3.0 * var]])
        H = np.array([[1.0, 0.0]])
        Pinf = np.array([[var, 0.0], [0.0, 3.0 * var / ell ** 2.0]])
        return F, L, Qc, H, Pinf

    def measurement_model(self, r=None, hyperparams=None):
        return np.array([[1.0, 0.0]])

    def state_transition(self, dt, hyperparams=None):
        ell = self._theta(hyperparams)[1]
        lam = np.sqrt(3.0) / ell
        return np.exp(-dt * lam) * (dt * np.array([[lam, 1.0], [-lam ** 2.0, -lam]]) + np.eye(2))

    def blocks(self, hyperparams=None):
        return [(_lib.BLOCK_MATERN32, 1, float(np.sqrt(3.0) / self._theta(hyperparams)[1]), 0.0)]


class Matern52(Prior):
    """priors.py:178-255."""
    state_dim = 3

    def __init__(self, variance=1.0, lengthscale=1.0):
        super().__init__(hyp=[variance, lengthscale])
        self.name = 'Matern-5/2'

    @property
    def variance(self):
        return softplus(self.hyp[0])

    @property
    def lengthscale(self):
        return softplus(self.hyp[1])

    def kernel_to_state_space(self, hyperparams=None):
        var, ell = self._theta(hyperparams)
        lam = 5.0 ** 0.5 / ell
        F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3.0, -3.0 * lam ** 2.0, -3.0 * lam]])
        L = np.array([[0.0], [0.0], [1.0]])
        Qc = np.array([[var * 400.0 * 5.0 ** 0.5 / 3.0 / ell ** 5.0]])
        H = np.array([[1.0, 0.0, 0.0]])
        kappa = 5.0 / 3.0 * var / ell ** 2.0
        Pinf = np.array([[var, 0.0, -kappa], [0.0, kappa, 0.0], [-kappa, 0.0, 25.0 * var / ell ** 4.0]])
        return F, L, Qc, H, Pinf

    def measurement_model(self, r=None, hyperparams=None):
        return np.array([[1.0, 0.0, 0.0]])

    def state_transition(self, dt, hyperparams=None):
        return _matern52_transition(dt, self._theta(hyperparams)[1])

    def blocks(self, hyperparams=None):
        return [(_lib.BLOCK_MATERN52, 1, float(np.sqrt(5.0) / self._theta(hyperparams)[1]), 0.0)]


def _matern52_transition(dt, ell):
    lam = np.sqrt(5.0) / ell
    dtlam = dt * lam
    return np.exp(-dtlam) * (
        dt * np.array([[lam * (0.5 * dtlam + 1.0), dtlam + 1.0, 0.5 * dt],
                       [-0.5 * dtlam * lam ** 2, lam * (1.0 - dtlam), 1.0 - 0.5 * dtlam],
                       [lam ** 3 * (0.5 * dtlam - 1.0), lam ** 2 * (dtlam - 3), lam * (0.5 * dtlam - 2.0)]])
        + np.eye(3))


class QuasiPeriodicMatern32(Prior):
    """priors.py:946-1084 (A and Pinf harmonic-major, H = kron(H_m, H_p): kept exactly as written)."""
    state_dim = 28

    def __init__(self, variance=1.0, lengthscale_periodic=1.0, period=1.0, lengthscale_matern=1.0, order=6):
        super().__init__(hyp=[variance, lengthscale_periodic, period, lengthscale_matern])
        self.name = 'Periodic'
        if order != 6:
            raise NotImplementedError("the reference hard-codes order 6 (priors.py:963-976)")
        self.order = order
        self.K = np.meshgrid(np.arange(self.order + 1), np.arange(self.order + 1))[1]
        factorial_mesh_K = np.array([[1.] * 7, [1.] * 7, [2.] * 7, [6.] * 7, [24.] * 7, [120.] * 7, [720.] * 7])
        b = np.array([[1., 0., 0., 0., 0., 0., 0.],
                      [0., 2., 0., 0., 0., 0., 0.],
                      [2., 0., 2., 0., 0., 0., 0.],
                      [0., 6., 0., 2., 0., 0., 0.],
                      [6., 0., 8., 0., 2., 0., 0.],
                      [0., 20., 0., 10., 0., 2., 0.],
                      [20., 0., 30., 0., 12., 0., 2.]])
        self.b_fmK_2K = b * (1. / factorial_mesh_K) * (2. ** -self.K)

    def kernel_to_state_space(self, hyperparams=None):
        var, ell_p, period, ell_m = self._theta(hyperparams)
        a = self.b_fmK_2K * ell_p ** (-2. * self.K) * np.exp(-1. / ell_p ** 2.) * 1.
        q2 = np.sum(a, axis=0)
        omega = 2 * np.pi / period
        F_p = np.kron(np.diag(np.arange(self.order + 1)), np.array([[0., -omega], [omega, 0.]]))
        L_p = np.eye(2 * (self.order + 1))
        Pinf_p = np.kron(np.diag(q2), np.eye(2))
        lam = 3.0 ** 0.5 / ell_m
        F_m = np.array([[0.0, 1.0], [-lam ** 2, -2 * lam]])
        L_m = np.array([[0], [1]])
        Qc_m = np.array([[12.0 * 3.0 ** 0.5 / ell_m ** 3.0 * var]])
        Pinf_m = np.array([[var, 0.0], [0.0, 3.0 * var / ell_m ** 2.0]])
        F = np.kron(F_m, np.eye(2 * (self.order + 1))) + np.kron(np.eye(2), F_p)
        L = np.kron(L_m, L_p)
        Qc = np.kron(Qc_m, Pinf_p)
        H = self.measurement_model()
        Pinf = sl.block_diag(*[np.kron(Pinf_m, q2[j] * np.eye(2)) for j in range(self.order + 1)])
        return F, L, Qc, H, Pinf

    def measurement_model(self, r=None, hyperparams=None):
        H_p = np.kron(np.ones([1, self.order + 1]), np.array([1., 0.]))
        return np.kron(np.array([[1.0, 0.0]]), H_p)

    def state_transition(self, dt, hyperparams=None):
        theta = self._theta(hyperparams)
        period, ell_m = theta[2], theta[3]
        lam = np.sqrt(3.0) / ell_m
        omega = 2 * np.pi / period
        blocks = []
        for j in range(self.order + 1):
            w = j * omega
            R = np.array([[np.cos(w * dt), -np.sin(w * dt)], [np.sin(w * dt), np.cos(w * dt)]])
            blocks.append(np.block([[(1. + dt * lam) * R, dt * R], [-dt * lam ** 2 * R, (1. - dt * lam) * R]]))
        return np.exp(-dt * lam) * sl.block_diag(*blocks)

    def blocks(self, hyperparams=None):
        theta = self._theta(hyperparams)
        lam = float(np.sqrt(3.0) / theta[3])
        omega = 2 * np.pi / theta[2]
        harmonics = np.arange(self.order + 1) * omega        # priors.py:1057
        return [(_lib.BLOCK_QP_MATERN32, 1, lam, float(harmonics[j])) for j in range(self.order + 1)]


class Exponential(Prior):
    """priors.py:50-104."""
    state_dim = 1

    def __init__(self, variance=1.0, lengthscale=1.0):
        super().__init__(hyp=[variance, lengthscale])
        self.name = 'Exponential'

    @property
    def variance(self):
        return softplus(self.hyp[0])

    @property
    def lengthscale(self):
        return softplus(self.hyp[1])

    def kernel_to_state_space(self, hyperparams=None):
        var, ell = self._theta(hyperparams)
        return np.array([[-1.0 / ell]]), np.array([[1.0]]), np.array([[2.0 * var / ell]]), np.array([[1.0]]), np.array([[var]])

    def measurement_model(self, r=None, hyperparams=None):
        return np.array([[1.0]])

    def state_transition(self, dt, hyperparams=None):
        return np.broadcast_to(np.exp(-dt / self._theta(hyperparams)[1]), [1, 1])

    def blocks(self, hyperparams=None):
        return [(_lib.BLOCK_MATERN12, 1, float(1.0 / self._theta(hyperparams)[1]), 0.0)]


class Matern12(Exponential):
    """priors.py:107-108."""
    pass


class Matern72(Prior):
    """priors.py:258-350."""
    state_dim = 4

    def __init__(self, variance=1.0, lengthscale=1.0):
        super().__init__(hyp=[variance, lengthscale])
        self.name = 'Matern-7/2'

    @property
    def variance(self):
        return softplus(self.hyp[0])

    @property
    def lengthscale(self):
        return softplus(self.hyp[1])

    def kernel_to_state_space(self, hyperparams=None):
        var, ell = self._theta(hyperparams)
        lam = 7.0 ** 0.5 / ell
        F = np.array([[0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0],
                      [-lam ** 4.0, -4.0 * lam ** 3.0, -6.0 * lam ** 2.0, -4.0 * lam]])
        L = np.array([[0.0], [0.0], [0.0], [1.0]])
        Qc = np.array([[var * 10976.0 * 7.0 ** 0.5 / 5.0 / ell ** 7.0]])
        H = np.array([[1.0, 0.0, 0.0, 0.0]])
        kappa = 7.0 / 5.0 * var / ell ** 2.0
        kappa2 = 9.8 * var / ell ** 4.0
        Pinf = np.array([[var, 0.0, -kappa, 0.0], [0.0, kappa, 0.0, -kappa2], [-kappa, 0.0, kappa2, 0.0],
                         [0.0, -kappa2, 0.0, 343.0 * var / ell ** 6.0]])
        return F, L, Qc, H, Pinf

    def measurement_model(self, r=None, hyperparams=None):
        return np.array([[1.0, 0.0, 0.0, 0.0]])

    def state_transition(self, dt, hyperparams=None):
        lam = np.sqrt(7.0) / self._theta(hyperparams)[1]
        lam2, dtlam = lam * lam, dt * lam
        lam3, dtlam2 = lam2 * lam, dtlam ** 2
        return np.exp(-dtlam) * (
            dt * np.array([[lam * (1.0 + 0.5 * dtlam + dtlam2 / 6.0), 1.0 + dtlam + 0.5 * dtlam2, 0.5 * dt * (1.0 + dtlam), dt ** 2 / 6],
                           [-dtlam2 * lam ** 2.0 / 6.0, lam * (1.0 + 0.5 * dtlam - 0.5 * dtlam2), 1.0 + dtlam - 0.5 * dtlam2,
                            dt * (0.5 - dtlam / 6.0)],
                           [lam3 * dtlam * (dtlam / 6.0 - 0.5), dtlam * lam2 * (0.5 * dtlam - 2.0),
                            lam * (1.0 - 2.5 * dtlam + 0.5 * dtlam2), 1.0 - dtlam + dtlam2 / 6.0],
                           [lam2 ** 2 * (dtlam - 1.0 - dtlam2 / 6.0), lam3 * (3.5 * dtlam - 4.0 - 0.5 * dtlam2),
                            lam2 * (4.0 * dtlam - 6.0 - 0.5 * dtlam2), lam * (1.5 * dtlam - 3.0 - dtlam2 / 6.0)]])
            + np.eye(4))

    def blocks(self, hyperparams=None):
        return [(_lib.BLOCK_MATERN72, 1, float(np.sqrt(7.0) / self._theta(hyperparams)[1]), 0.0)]


def _rotation(dt, omega):
    """utils.py:253-265."""
    return np.array([[np.cos(omega * dt), -np.sin(omega * dt)], [np.sin(omega * dt), np.cos(omega * dt)]])


class Cosine(Prior):
    """priors.py:353-408 (hyp is the scalar frequency)."""
    state_dim = 2

    def __init__(self, frequency=1.0):
        super().__init__(hyp=frequency)
        self.name = 'Cosine'

    @property
    def frequency(self):
        return softplus(self.hyp)

    def kernel_to_state_space(self, hyperparams=None):
        omega = float(np.reshape(self._theta(hyperparams), -1)[0])
        return np.array([[0.0, -omega], [omega, 0.0]]), [], [], np.array([[1.0, 0.0]]), np.eye(2)

    def measurement_model(self, r=None, hyperparams=None):
        return np.array([[1.0, 0.0]])

    def state_transition(self, dt, hyperparams=None):
        return _rotation(dt, float(np.reshape(self._theta(hyperparams), -1)[0]))

    def blocks(self, hyperparams=None):
        return [(_lib.BLOCK_ROTATION, 1, 0.0, float(np.reshape(self._theta(hyperparams), -1)[0]))]


class Periodic(QuasiPeriodicMatern32):
    """priors.py:724-820 (the series tables b_fmK_2K are the ones QuasiPeriodicMatern32 builds, :741-757)."""
    state_dim = 14

    def __init__(self, variance=1.0, lengthscale=1.0, period=1.0, order=6):
        QuasiPeriodicMatern32.__init__(self, order=order)
        Prior.__init__(self, hyp=[variance, lengthscale, period])
        self.name = 'Periodic'

    def _q2(self, var, ell):
        return np.sum(self.b_fmK_2K * ell ** (-2. * self.K) * np.exp(-1. / ell ** 2.) * var, axis=0)

    def kernel_to_state_space(self, hyperparams=None):
        var, ell, period = self._theta(hyperparams)
        omega = 2 * np.pi / period
        F = np.kron(np.diag(np.arange(self.order + 1)), np.array([[0., -omega], [omega, 0.]]))
        n = 2 * (self.order + 1)
        return F, np.eye(n), np.zeros(n), self.measurement_model(), np.kron(np.diag(self._q2(var, ell)), np.eye(2))

    def measurement_model(self, r=None, hyperparams=None):
        return np.kron(np.ones([1, self.order + 1]), np.array([1., 0.]))

    def state_transition(self, dt, hyperparams=None):
        omega = 2 * np.pi / self._theta(hyperparams)[2]
        return sl.block_diag(*[_rotation(dt, j * omega) for j in range(self.order + 1)])

    def blocks(self, hyperparams=None):
        omega = 2 * np.pi / self._theta(hyperparams)[2]
        return [(_lib.BLOCK_ROTATION, 1, 0.0, float(j * omega)) for j in range(self.order + 1)]


class QuasiPeriodicMatern12(QuasiPeriodicMatern32):
    """priors.py:823-939."""
    state_dim = 14

    def __init__(self, variance=1.0, lengthscale_periodic=1.0, period=1.0, lengthscale_matern=1.0, order=6):
        super().__init__(variance, lengthscale_periodic, period, lengthscale_matern, order)
        self.name = 'Quasi-periodic Matern-1/2'

    def kernel_to_state_space(self, hyperparams=None):
        var, ell_p, period, ell_m = self._theta(hyperparams)
        q2 = np.sum(self.b_fmK_2K * ell_p ** (-2. * self.K) * np.exp(-1. / ell_p ** 2.) * 1., axis=0)
        omega = 2 * np.pi / period
        n = 2 * (self.order + 1)
        F_p = np.kron(np.diag(np.arange(self.order + 1)), np.array([[0., -omega], [omega, 0.]]))
        Pinf_p = np.kron(np.diag(q2), np.eye(2))
        F = np.kron(np.array([[-1.0 / ell_m]]), np.eye(n)) + F_p
        return F, np.eye(n), np.kron(Pinf_p, np.array([[2.0 * var / ell_m]])), self.measurement_model(), np.kron(np.array([[var]]), Pinf_p)

    def measurement_model(self, r=None, hyperparams=None):
        return np.kron(np.array([[1.0]]), np.kron(np.ones([1, self.order + 1]), np.array([1., 0.])))

    def state_transition(self, dt, hyperparams=None):
        theta = self._theta(hyperparams)
        omega = 2 * np.pi / theta[2]
        return np.exp(-dt / theta[3]) * sl.block_diag(*[_rotation(dt, j * omega) for j in range(self.order + 1)])

    def blocks(self, hyperparams=None):
        theta = self._theta(hyperparams)
        omega = 2 * np.pi / theta[2]
        return [(_lib.BLOCK_ROTATION, 1, float(1.0 / theta[3]), float(j * omega)) for j in range(self.order + 1)]


class SubbandMatern12(Prior):
    """priors.py:411-494."""
    state_dim = 2

    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        super().__init__(hyp=[variance, lengthscale, radial_frequency])
        self.name = 'Subband Matern-1/2'

    def kernel_to_state_space(self, hyperparams=None):
        var, ell, omega = self._theta(hyperparams)
        F = np.kron(np.array([[-1.0 / ell]]), np.eye(2)) + np.array([[0.0, -omega], [omega, 0.0]])
        return F, np.eye(2), np.kron(np.eye(2), np.array([[2.0 * var / ell]])), np.array([[1.0, 0.0]]), var * np.eye(2)

    def measurement_model(self, r=None, hyperparams=None):
        return np.array([[1.0, 0.0]])

    def state_transition(self, dt, hyperparams=None):
        theta = self._theta(hyperparams)
        return np.exp(-dt / theta[1]) * _rotation(dt, theta[2])

    def blocks(self, hyperparams=None):
        theta = self._theta(hyperparams)
        return [(_lib.BLOCK_ROTATION, 1, float(1.0 / theta[1]), float(theta[2]))]


class SubbandExponential(SubbandMatern12):
    """priors.py:497-498."""
    pass


class SubbandMatern32(Prior):
    """priors.py:501-602: its transition is the 4 x 4 block kind of QuasiPeriodicMatern32's harmonics."""
    state_dim = 4

    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        super().__init__(hyp=[variance, lengthscale, radial_frequency])
        self.name = 'Subband Matern-3/2'

    def kernel_to_state_space(self, hyperparams=None):
        var, ell, omega = self._theta(hyperparams)
        lam = 3.0 ** 0.5 / ell
        F_mat = np.array([[0.0, 1.0], [-lam ** 2, -2 * lam]])
        F = np.kron(F_mat, np.eye(2)) + np.kron(np.eye(2), np.array([[0.0, -omega], [omega, 0.0]]))
        L = np.kron(np.array([[0], [1]]), np.eye(2))
        Qc = np.kron(np.eye(2), np.array([[12.0 * 3.0 ** 0.5 / ell ** 3.0 * var]]))
        Pinf = np.kron(np.array([[var, 0.0], [0.0, 3.0 * var / ell ** 2.0]]), np.eye(2))
        return F, L, Qc, self.measurement_model(), Pinf

    def measurement_model(self, r=None, hyperparams=None):
        return np.kron(np.array([[1.0, 0.0]]), np.array([[1.0, 0.0]]))

    def state_transition(self, dt, hyperparams=None):
        theta = self._theta(hyperparams)
        lam = np.sqrt(3.0) / theta[1]
        R = _rotation(dt, theta[2])
        return np.exp(-dt * lam) * np.block([[(1. + dt * lam) * R, dt * R], [-dt * lam ** 2 * R, (1. - dt * lam) * R]])

    def blocks(self, hyperparams=None):
        theta = self._theta(hyperparams)
        return [(_lib.BLOCK_QP_MATERN32, 1, float(np.sqrt(3.0) / theta[1]), float(theta[2]))]


class SubbandMatern52(Prior):
    """priors.py:605-721 (product of Cosine and Matern-5/2): its transition is the 6 x 6 block kind — the Matern-5/2
    polynomial block times a rotation of every (cos, sin) pair."""
    state_dim = 6

    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        super().__init__(hyp=[variance, lengthscale, radial_frequency])
        self.name = 'Subband Matern-5/2'

    def kernel_to_state_space(self, hyperparams=None):
        var, ell, omega = self._theta(hyperparams)
        lam = 5.0 ** 0.5 / ell
        F_mat = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3.0, -3.0 * lam ** 2.0, -3.0 * lam]])
        F = np.kron(F_mat, np.eye(2)) + np.kron(np.eye(3), np.array([[0.0, -omega], [omega, 0.0]]))
        L = np.kron(np.array([[0.0], [0.0], [1.0]]), np.eye(2))
        Qc = np.kron(np.eye(2), np.array([[var * 400.0 * 5.0 ** 0.5 / 3.0 / ell ** 5.0]]))
        kappa = 5.0 / 3.0 * var / ell ** 2.0
        Pinf_mat = np.array([[var, 0.0, -kappa], [0.0, kappa, 0.0], [-kappa, 0.0, 25.0 * var / ell ** 4.0]])
        return F, L, Qc, self.measurement_model(), np.kron(Pinf_mat, np.eye(2))

    def measurement_model(self, r=None, hyperparams=None):
        return np.kron(np.array([[1.0, 0.0, 0.0]]), np.array([[1.0, 0.0]]))

    def state_transition(self, dt, hyperparams=None):
        theta = self._theta(hyperparams)
        lam = 5.0 ** 0.5 / theta[1]
        R = _rotation(dt, theta[2])
        dl = dt * lam
        return np.exp(-dl) * np.block([
            [0.5 * (2. + dl * (2. + dl)) * R, dt * (1. + dl) * R, 0.5 * dt ** 2 * R],
            [-0.5 * dt ** 2 * lam ** 3 * R, (1. + dl * (1. - dl)) * R, -0.5 * dt * (-2. + dl) * R],
            [0.5 * dt * lam ** 3 * (-2. + dl) * R, dt * lam ** 2 * (-3. + dl) * R, 0.5 * (2. + dl * (-4. + dl)) * R]])

    def blocks(self, hyperparams=None):
        theta = self._theta(hyperparams)
        return [(_lib.BLOCK_SUBBAND_MATERN52, 1, float(np.sqrt(5.0) / theta[1]), float(theta[2]))]


class Sum(object):
    """priors.py:1347-1414."""
    def __init__(self, components):
        self.components = components
        self.hyp = [c.hyp for c in components]
        self.name = 'Sum'
        self.state_dim = sum(c.state_dim for c in components)

    def _theta(self, hyperparams):
        return softplus_list(self.hyp) if hyperparams is None else hyperparams

    def kernel_to_state_space(self, hyperparams=None):
        hyperparams = self._theta(hyperparams)
        parts = [c.kernel_to_state_space(h) for c, h in zip(self.components, hyperparams)]
        F = sl.block_diag(*[p[0] for p in parts])
        L = sl.block_diag(*[p[1] for p in parts])
        Qc = sl.block_diag(*[p[2] for p in parts])
        H = np.block([p[3] for p in parts])
        Pinf = sl.block_diag(*[p[4] for p in parts])
        return F, L, Qc, H, Pinf

    def measurement_model(self, r=None, hyperparams=None):
        hyperparams = self._theta(hyperparams)
        return np.block([c.measurement_model(r, h) for c, h in zip(self.components, hyperparams)])

    def state_transition(self, dt, hyperparams=None):
        hyperparams = self._theta(hyperparams)
        return sl.block_diag(*[c.state_transition(dt, h) for c, h in zip(self.components, hyperparams)])

    def blocks(self, hyperparams=None):
        hyperparams = self._theta(hyperparams)
        out = []
        for c, h in zip(self.components, hyperparams):
            out += c.blocks(h)
        return out


class Independent(Sum):
    """priors.py:1417-1487: a stack of independent GP priors, each component fed to the likelihood — `Sum` with a
    block-diagonal measurement model (one function value per component: func_dim = number of components)."""
    def __init__(self, components):
        Sum.__init__(self, components)
        self.name = 'Independent'

    def kernel_to_state_space(self, hyperparams=None):
        hyperparams = self._theta(hyperparams)
        F, L, Qc, _, Pinf = Sum.kernel_to_state_space(self, hyperparams)
        H = sl.block_diag(*[c.kernel_to_state_space(h)[3] for c, h in zip(self.components, hyperparams)])
        return F, L, Qc, H, Pinf

    def measurement_model(self, r=None, hyperparams=None):
        hyperparams = self._theta(hyperparams)
        return sl.block_diag(*[c.measurement_model(r, h) for c, h in zip(self.components, hyperparams)])


class Separate(Independent):
    pass


def _matern52_spatial_kernel(X, X2, lengthscale):
    """kernels.py:29-36, 90-93 with utils.py:626-654 (Matern-5/2 covariance of the spatial factor)."""
    X, X2 = X / lengthscale, X2 / lengthscale
    Xs, X2s = np.sum(np.square(X), axis=-1), np.sum(np.square(X2), axis=-1)
    r2 = -2 * np.tensordot(X, X2, [[-1], [-1]]) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
    r = np.sqrt(np.maximum(r2, 1e-36))
    sqrt5 = np.sqrt(5.0)
    return (1.0 + sqrt5 * r + 5.0 / 3.0 * np.square(r)) * np.exp(-sqrt5 * r)


class SpatioTemporalMatern52(Prior):
    """priors.py:1087-1181."""
    def __init__(self, variance=1.0, lengthscale_time=1.0, lengthscale_space=1.0, spatial_dims=1, z=None,
                 fixed_grid=False):
        super().__init__(hyp=[variance, lengthscale_time, lengthscale_space])
        if spatial_dims != 1:
            raise NotImplementedError("one spatial dimension only")
        self.spatial_dims = spatial_dims
        if z is None:
            z = np.linspace(-3., 3., num=15)
        self.z = np.asarray(z, dtype=np.float64).reshape(-1, self.spatial_dims)
        self.M = self.z.shape[0]
        self.state_dim = 3 * self.M
        self.fixed_grid = fixed_grid
        self.name = 'Spatio-Temporal Matern-5/2'

    def kernel_to_state_space(self, hyperparams=None):
        var, ell_time, ell_space = self._theta(hyperparams)
        Kmm = _matern52_spatial_kernel(self.z, self.z, ell_space)
        lam = 5.0 ** 0.5 / ell_time
        F_time = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3.0, -3.0 * lam ** 2.0, -3.0 * lam]])
        F = np.kron(np.eye(self.M), F_time)
        L = np.array([[0.0], [0.0], [1.0]])
        Qc = np.array([[var * 400.0 * 5.0 ** 0.5 / 3.0 / ell_time ** 5.0]])
        kappa = 5.0 / 3.0 * var / ell_time ** 2.0
        Pinf_time = np.array([[var, 0.0, -kappa], [0.0, kappa, 0.0], [-kappa, 0.0, 25.0 * var / ell_time ** 4.0]])
        return F, L, Qc, None, np.kron(Kmm, Pinf_time)

    def spatial_weights(self, r, hyperparams=None):
        """Kx = Kxz Kzz^-1 for a batch of spatial locations r[N] -> [N, M] (priors.py:1154-1159)."""
        ell_space = self._theta(hyperparams)[2]
        r = np.asarray(r, dtype=np.float64).reshape(-1, 1)
        if self.fixed_grid:
            return np.tile(np.eye(self.M)[None], (r.shape[0], 1, 1))
        Kzz = _matern52_spatial_kernel(self.z, self.z, ell_space)
        Kxz = _matern52_spatial_kernel(r, self.z, ell_space)
        return sl.cho_solve(sl.cho_factor(Kzz), Kxz.T).T

    def measurement_model(self, r, hyperparams=None):
        Kx = self.spatial_weights(r, hyperparams)
        return np.kron(Kx, np.array([[1.0, 0.0, 0.0]]))

    def state_transition(self, dt, hyperparams=None):
        return np.kron(np.eye(self.M), _matern52_transition(dt, self._theta(hyperparams)[1]))

    def blocks(self, hyperparams=None):
        return [(_lib.BLOCK_MATERN52, int(self.M), float(np.sqrt(5.0) / self._theta(hyperparams)[1]), 0.0)]


def _matern32_spatial_kernel(X, X2, lengthscale):
    """kernels.py:29-36, 72-75 with utils.py:626-654."""
    X, X2 = X / lengthscale, X2 / lengthscale
    Xs, X2s = np.sum(np.square(X), axis=-1), np.sum(np.square(X2), axis=-1)
    r2 = -2 * np.tensordot(X, X2, [[-1], [-1]]) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
    r = np.sqrt(np.maximum(r2, 1e-36))
    sqrt3 = np.sqrt(3.0)
    return (1.0 + sqrt3 * r) * np.exp(-sqrt3 * r)


class SpatialMatern52(SpatioTemporalMatern52):
    """priors.py:1184-1267: one lengthscale shared by time and space; hyp = [variance, lengthscale]."""
    def __init__(self, variance=1.0, lengthscale=1.0, z=None, fixed_grid=False):
        SpatioTemporalMatern52.__init__(self, variance, lengthscale, lengthscale, z=z, fixed_grid=fixed_grid)
        Prior.__init__(self, hyp=[variance, lengthscale])
        self.name = 'Spatial Matern-5/2'

    def _theta(self, hyperparams):
        th = Prior._theta(self, hyperparams)
        return np.array([th[0], th[1], th[1]]) if np.size(th) == 2 else th      # (variance, time = space lengthscale)


class SpatialMatern32(Prior):
    """priors.py:1270-1344."""
    def __init__(self, variance=1.0, lengthscale=1.0, z=None, fixed_grid=False):
        super().__init__(hyp=[variance, lengthscale])
        if z is None:
            z = np.linspace(-3., 3., num=15)
        self.z = np.asarray(z, dtype=np.float64).reshape(-1, 1)
        self.M = self.z.shape[0]
        self.state_dim = 2 * self.M
        self.fixed_grid = fixed_grid
        self.name = 'Spatial Matern-3/2'

    def kernel_to_state_space(self, hyperparams=None):
        var, ell = self._theta(hyperparams)
        Kmm = _matern32_spatial_kernel(self.z, self.z, ell)
        lam = 3.0 ** 0.5 / ell
        F = np.kron(np.eye(self.M), np.array([[0.0, 1.0], [-lam ** 2, -2 * lam]]))
        Pinf_time = np.array([[var, 0.0], [0.0, 3.0 * var / ell ** 2.0]])
        return F, np.array([[0], [1]]), np.array([[12.0 * 3.0 ** 0.5 / ell ** 3.0 * var]]), None, np.kron(Kmm, Pinf_time)

    def spatial_weights(self, r, hyperparams=None):
        ell = self._theta(hyperparams)[1]
        r = np.asarray(r, dtype=np.float64).reshape(-1, 1)
        if self.fixed_grid:
            return np.tile(np.eye(self.M)[None], (r.shape[0], 1, 1))
        Kzz = _matern32_spatial_kernel(self.z, self.z, ell)
        Kxz = _matern32_spatial_kernel(r, self.z, ell)
        return sl.cho_solve(sl.cho_factor(Kzz), Kxz.T).T

    def measurement_model(self, r, hyperparams=None):
        return np.kron(self.spatial_weights(r, hyperparams), np.array([[1.0, 0.0]]))

    def state_transition(self, dt, hyperparams=None):
        lam = np.sqrt(3.0) / self._theta(hyperparams)[1]
        return np.kron(np.eye(self.M), np.exp(-dt * lam) * (dt * np.array([[lam, 1.0], [-lam ** 2.0, -lam]]) + np.eye(2)))

    def blocks(self, hyperparams=None):
        return [(_lib.BLOCK_MATERN32, int(self.M), float(np.sqrt(3.0) / self._theta(hyperparams)[1]), 0.0)]


class Matern52FixedVar(Matern52):
    """priors.py:1552-1600: the variance is a constant of the object, hyp is the (scalar) lengthscale."""
    def __init__(self, variance=1.0, lengthscale=1.0):
        Prior.__init__(self, hyp=lengthscale)
        self.variance_fixed = variance
        self.name = 'Matern-5/2'

    @property
    def variance(self):
        return self.variance_fixed

    @property
    def lengthscale(self):
        return softplus(self.hyp)

    def _theta(self, hyperparams):
        ell = float(np.reshape(softplus(self.hyp) if hyperparams is None else np.asarray(hyperparams, dtype=np.float64), -1)[-1])
        return np.array([self.variance_fixed, ell])


class SubbandExponentialFixedVar(SubbandMatern12):
    """priors.py:1498-1549: constant variance, hyp = [lengthscale, radial_frequency]."""
    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        Prior.__init__(self, hyp=[lengthscale, radial_frequency])
        self.variance_fixed = variance
        self.name = 'Subband Matern-1/2'

    def _theta(self, hyperparams):
        th = softplus(self.hyp) if hyperparams is None else np.asarray(hyperparams, dtype=np.float64)
        th = np.reshape(th, -1)
        return np.array([self.variance_fixed, th[-2], th[-1]])
