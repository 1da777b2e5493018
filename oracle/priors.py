"""ORACLE (test infrastructure, not product): GP priors as LTI SDEs and their per-step discretisation.

NumPy fp64 restatement of the parts of /root/reference/kalmanjax/priors.py (and kernels.py, utils.py)
that the Kalman filter / RTS smoother touch: Pinf, H (measurement model) and the closed-form
A(dt) = expm(F dt).  Hyperparameters are the already-transformed (positive) values.
Restated: the priors named by BASELINE.json's configs (Matern-3/2, Matern-5/2, QuasiPeriodicMatern32, Sum,
SpatioTemporalMatern52) and, from SURVEY 8(f)-4's model zoo, Matern-1/2 (Exponential), Matern-7/2, Cosine, Periodic,
QuasiPeriodicMatern12, SubbandMatern12 / SubbandMatern32, SpatialMatern52 / SpatialMatern32.
"""
import numpy as np
import scipy.linalg as sl


class Matern32(object):
    """priors.py:111-175."""
    state_dim = 2

    def __init__(self, variance=1.0, lengthscale=1.0):
        self.hyp = np.array([variance, lengthscale], dtype=float)

    def kernel_to_state_space(self):
        var, ell = self.hyp
        H = np.array([[1.0, 0.0]])
        Pinf = np.array([[var, 0.0],
                         [0.0, 3.0 * var / ell ** 2.0]])                     # :154-155
        return H, Pinf

    def measurement_model(self, r=None):
        return np.array([[1.0, 0.0]])                                        # :160

    def state_transition(self, dt):
        lam = np.sqrt(3.0) / self.hyp[1]                                     # :172-174
        return np.exp(-dt * lam) * (dt * np.array([[lam, 1.0], [-lam ** 2.0, -lam]]) + np.eye(2))


class Matern52(object):
    """priors.py:178-255."""
    state_dim = 3

    def __init__(self, variance=1.0, lengthscale=1.0):
        self.hyp = np.array([variance, lengthscale], dtype=float)

    def kernel_to_state_space(self):
        var, ell = self.hyp
        H = np.array([[1.0, 0.0, 0.0]])
        kappa = 5.0 / 3.0 * var / ell ** 2.0                                 # :227
        Pinf = np.array([[var, 0.0, -kappa],
                         [0.0, kappa, 0.0],
                         [-kappa, 0.0, 25.0 * var / ell ** 4.0]])            # :228-230
        return H, Pinf

    def measurement_model(self, r=None):
        return np.array([[1.0, 0.0, 0.0]])                                   # :235

    def state_transition(self, dt):
        return matern52_transition(dt, self.hyp[1])


def matern52_transition(dt, ell):
    """priors.py:246-255 (also used by the spatio-temporal prior, priors.py:1171-1179)."""
    lam = np.sqrt(5.0) / ell
    dtlam = dt * lam
    return np.exp(-dtlam) * (
        dt * np.array([[lam * (0.5 * dtlam + 1.0), dtlam + 1.0, 0.5 * dt],
                       [-0.5 * dtlam * lam ** 2, lam * (1.0 - dtlam), 1.0 - 0.5 * dtlam],
                       [lam ** 3 * (0.5 * dtlam - 1.0), lam ** 2 * (dtlam - 3), lam * (0.5 * dtlam - 2.0)]])
        + np.eye(3))


def rotation_matrix(dt, omega):
    """utils.py:253-265."""
    return np.array([[np.cos(omega * dt), -np.sin(omega * dt)],
                     [np.sin(omega * dt), np.cos(omega * dt)]])


class QuasiPeriodicMatern32(object):
    """priors.py:946-1084.  State is harmonic-major 4x4 blocks for A and Pinf, while H is
    kron(H_m, H_p) = ones at indices 0,2,..,12 — reproduced exactly as written (SURVEY 8a trap ix)."""
    order = 6
    state_dim = 28

    def __init__(self, variance=1.0, lengthscale_periodic=1.0, period=1.0, lengthscale_matern=1.0):
        self.hyp = np.array([variance, lengthscale_periodic, period, lengthscale_matern], dtype=float)
        K = np.meshgrid(np.arange(self.order + 1), np.arange(self.order + 1))[1]          # :962
        factorial_mesh_K = np.array([[1.] * 7, [1.] * 7, [2.] * 7, [6.] * 7, [24.] * 7, [120.] * 7, [720.] * 7])
        b = np.array([[1., 0., 0., 0., 0., 0., 0.],
                      [0., 2., 0., 0., 0., 0., 0.],
                      [2., 0., 2., 0., 0., 0., 0.],
                      [0., 6., 0., 2., 0., 0., 0.],
                      [6., 0., 8., 0., 2., 0., 0.],
                      [0., 20., 0., 10., 0., 2., 0.],
                      [20., 0., 30., 0., 12., 0., 2.]])                                   # :970-976
        self.K = K
        self.b_fmK_2K = b * (1. / factorial_mesh_K) * (2. ** -K)                          # :977

    def q2(self):
        ell_p = self.hyp[1]
        a = self.b_fmK_2K * ell_p ** (-2. * self.K) * np.exp(-1. / ell_p ** 2.) * 1.      # :1001
        return np.sum(a, axis=0)                                                          # :1002

    def kernel_to_state_space(self):
        var, ell_p, period, ell_m = self.hyp
        q2 = self.q2()
        Pinf_m = np.array([[var, 0.0],
                           [0.0, 3.0 * var / ell_m ** 2.0]])                              # :1018-1019
        Pinf = sl.block_diag(*[np.kron(Pinf_m, q2[j] * np.eye(2)) for j in range(self.order + 1)])  # :1026-1034
        return self.measurement_model(), Pinf

    def measurement_model(self, r=None):
        H_p = np.kron(np.ones([1, self.order + 1]), np.array([1., 0.]))                   # :1039
        H_m = np.array([[1.0, 0.0]])
        return np.kron(H_m, H_p)                                                          # :1041

    def state_transition(self, dt):
        period, ell_m = self.hyp[2], self.hyp[3]
        lam = np.sqrt(3.0) / ell_m                                                        # :1054
        omega = 2 * np.pi / period
        blocks = []
        for j in range(self.order + 1):
            R = rotation_matrix(dt, j * omega)                                            # :1078
            blocks.append(np.block([[(1. + dt * lam) * R, dt * R],
                                    [-dt * lam ** 2 * R, (1. - dt * lam) * R]]))          # :1079-1082
        return np.exp(-dt * lam) * sl.block_diag(*blocks)                                 # :1066-1074


class Matern12(object):
    """priors.py:50-108 (Exponential / Matern-1/2)."""
    state_dim = 1

    def __init__(self, variance=1.0, lengthscale=1.0):
        self.hyp = np.array([variance, lengthscale], dtype=float)

    def kernel_to_state_space(self):
        return np.array([[1.0]]), np.array([[self.hyp[0]]])                               # :85-86

    def measurement_model(self, r=None):
        return np.array([[1.0]])                                                          # :90

    def state_transition(self, dt):
        return np.broadcast_to(np.exp(-dt / self.hyp[1]), [1, 1])                         # :103


Exponential = Matern12


class Matern72(object):
    """priors.py:258-350."""
    state_dim = 4

    def __init__(self, variance=1.0, lengthscale=1.0):
        self.hyp = np.array([variance, lengthscale], dtype=float)

    def kernel_to_state_space(self):
        var, ell = self.hyp
        kappa = 7.0 / 5.0 * var / ell ** 2.0                                              # :312
        kappa2 = 9.8 * var / ell ** 4.0                                                   # :313
        Pinf = np.array([[var, 0.0, -kappa, 0.0],
                         [0.0, kappa, 0.0, -kappa2],
                         [-kappa, 0.0, kappa2, 0.0],
                         [0.0, -kappa2, 0.0, 343.0 * var / ell ** 6.0]])                  # :314-317
        return np.array([[1.0, 0.0, 0.0, 0.0]]), Pinf

    def measurement_model(self, r=None):
        return np.array([[1.0, 0.0, 0.0, 0.0]])                                           # :322

    def state_transition(self, dt):
        lam = np.sqrt(7.0) / self.hyp[1]                                                  # :335-350
        lam2 = lam * lam
        lam3 = lam2 * lam
        dtlam = dt * lam
        dtlam2 = dtlam ** 2
        return np.exp(-dtlam) * (
            dt * np.array([[lam * (1.0 + 0.5 * dtlam + dtlam2 / 6.0), 1.0 + dtlam + 0.5 * dtlam2,
                            0.5 * dt * (1.0 + dtlam), dt ** 2 / 6],
                           [-dtlam2 * lam ** 2.0 / 6.0, lam * (1.0 + 0.5 * dtlam - 0.5 * dtlam2),
                            1.0 + dtlam - 0.5 * dtlam2, dt * (0.5 - dtlam / 6.0)],
                           [lam3 * dtlam * (dtlam / 6.0 - 0.5), dtlam * lam2 * (0.5 * dtlam - 2.0),
                            lam * (1.0 - 2.5 * dtlam + 0.5 * dtlam2), 1.0 - dtlam + dtlam2 / 6.0],
                           [lam2 ** 2 * (dtlam - 1.0 - dtlam2 / 6.0), lam3 * (3.5 * dtlam - 4.0 - 0.5 * dtlam2),
                            lam2 * (4.0 * dtlam - 6.0 - 0.5 * dtlam2), lam * (1.5 * dtlam - 3.0 - dtlam2 / 6.0)]])
            + np.eye(4))


class Cosine(object):
    """priors.py:353-408: a pure rotation, Pinf = I, no process noise."""
    state_dim = 2

    def __init__(self, frequency=1.0):
        self.hyp = np.array(frequency, dtype=float)

    def kernel_to_state_space(self):
        return np.array([[1.0, 0.0]]), np.eye(2)                                          # :386-389

    def measurement_model(self, r=None):
        return np.array([[1.0, 0.0]])                                                     # :393

    def state_transition(self, dt):
        return rotation_matrix(dt, float(np.reshape(self.hyp, -1)[0]))                    # :406


class Periodic(QuasiPeriodicMatern32):
    """priors.py:724-820: sum of order + 1 = 7 cosines (2 x 2 rotations), Pinf = kron(diag(q2), I2), no process noise.
    Inherits the series coefficients b_fmK_2K (identical tables, :741-757)."""
    state_dim = 14

    def __init__(self, variance=1.0, lengthscale=1.0, period=1.0):
        QuasiPeriodicMatern32.__init__(self)
        self.hyp = np.array([variance, lengthscale, period], dtype=float)

    def q2(self):
        var, ell = self.hyp[0], self.hyp[1]
        a = self.b_fmK_2K * ell ** (-2. * self.K) * np.exp(-1. / ell ** 2.) * var         # :773
        return np.sum(a, axis=0)                                                          # :774

    def kernel_to_state_space(self):
        return self.measurement_model(), np.kron(np.diag(self.q2()), np.eye(2))           # :781

    def measurement_model(self, r=None):
        return np.kron(np.ones([1, self.order + 1]), np.array([1., 0.]))                  # :787

    def state_transition(self, dt):
        omega = 2 * np.pi / self.hyp[2]                                                   # :802
        return sl.block_diag(*[rotation_matrix(dt, j * omega) for j in range(self.order + 1)])   # :803-819


class QuasiPeriodicMatern12(QuasiPeriodicMatern32):
    """priors.py:823-939: kron(Matern-1/2, Periodic): blocks exp(-dt / ell_m) R(j omega dt), Pinf = var kron(diag(q2), I2)."""
    state_dim = 14

    def kernel_to_state_space(self):
        var = self.hyp[0]
        return self.measurement_model(), np.kron(np.array([[var]]), np.kron(np.diag(self.q2()), np.eye(2)))   # :899

    def measurement_model(self, r=None):
        return np.kron(np.array([[1.0]]), np.kron(np.ones([1, self.order + 1]), np.array([1., 0.])))           # :904-907

    def state_transition(self, dt):
        period, ell_m = self.hyp[2], self.hyp[3]
        omega = 2 * np.pi / period                                                        # :921
        return np.exp(-dt / ell_m) * sl.block_diag(*[rotation_matrix(dt, j * omega) for j in range(self.order + 1)])   # :930-938


class SubbandMatern12(object):
    """priors.py:411-494: kron(Matern-1/2, Cosine)."""
    state_dim = 2

    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        self.hyp = np.array([variance, lengthscale, radial_frequency], dtype=float)

    def kernel_to_state_space(self):
        return np.array([[1.0, 0.0]]), np.kron(np.array([[self.hyp[0]]]), np.eye(2))      # :469-470

    def measurement_model(self, r=None):
        return np.array([[1.0, 0.0]])                                                     # :477

    def state_transition(self, dt):
        return np.exp(-dt / self.hyp[1]) * rotation_matrix(dt, self.hyp[2])               # :492-493


SubbandExponential = SubbandMatern12


class SubbandMatern32(object):
    """priors.py:501-602: kron(Matern-3/2, Cosine)."""
    state_dim = 4

    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        self.hyp = np.array([variance, lengthscale, radial_frequency], dtype=float)

    def kernel_to_state_space(self):
        var, ell = self.hyp[0], self.hyp[1]
        Pinf_mat = np.array([[var, 0.0], [0.0, 3.0 * var / ell ** 2.0]])                  # :561-562
        return self.measurement_model(), np.kron(Pinf_mat, np.eye(2))                     # :575

    def measurement_model(self, r=None):
        return np.kron(np.array([[1.0, 0.0]]), np.array([[1.0, 0.0]]))                    # :583

    def state_transition(self, dt):
        lam = np.sqrt(3.0) / self.hyp[1]                                                  # :596
        R = rotation_matrix(dt, self.hyp[2])
        return np.exp(-dt * lam) * np.block([[(1. + dt * lam) * R, dt * R],
                                             [-dt * lam ** 2 * R, (1. - dt * lam) * R]])  # :598-601


class SubbandMatern52(object):
    """priors.py:605-721: kron(Matern-5/2, Cosine)."""
    state_dim = 6

    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        self.hyp = np.array([variance, lengthscale, radial_frequency], dtype=float)

    def kernel_to_state_space(self):
        var, ell = self.hyp[0], self.hyp[1]
        kappa = 5.0 / 3.0 * var / ell ** 2.0                                              # :667
        Pinf_mat = np.array([[var, 0.0, -kappa], [0.0, kappa, 0.0], [-kappa, 0.0, 25.0 * var / ell ** 4.0]])
        return self.measurement_model(), np.kron(Pinf_mat, np.eye(2))                     # :685

    def measurement_model(self, r=None):
        return np.kron(np.array([[1.0, 0.0, 0.0]]), np.array([[1.0, 0.0]]))               # :690-693

    def state_transition(self, dt):
        lam = 5.0 ** 0.5 / self.hyp[1]                                                    # :706
        R = rotation_matrix(dt, self.hyp[2])
        return np.exp(-dt * lam) * np.block([                                             # :708-712
            [0.5 * (2. + dt * lam * (2. + dt * lam)) * R, dt * (1. + dt * lam) * R, 0.5 * dt ** 2 * R],
            [-0.5 * dt ** 2 * lam ** 3 * R, (1. + dt * lam * (1. - dt * lam)) * R, -0.5 * dt * (-2. + dt * lam) * R],
            [0.5 * dt * lam ** 3 * (-2. + dt * lam) * R, dt * lam ** 2 * (-3. + dt * lam) * R, 0.5 * (2. + dt * lam * (-4. + dt * lam)) * R]])


class Sum(object):
    """priors.py:1347-1414: block-diagonal stacking, H concatenated."""
    def __init__(self, components):
        self.components = components
        self.state_dim = sum(c.state_dim for c in components)

    def kernel_to_state_space(self):
        parts = [c.kernel_to_state_space() for c in self.components]
        return np.block([p[0] for p in parts]), sl.block_diag(*[p[1] for p in parts])

    def measurement_model(self, r=None):
        return np.block([c.measurement_model(r) for c in self.components])

    def state_transition(self, dt):
        return sl.block_diag(*[c.state_transition(dt) for c in self.components])


def matern52_kernel(X, X2, lengthscale):
    """kernels.py:29-36 + 90-93 with utils.py:626-654 (GPflow-style squared distance, clipped)."""
    X, X2 = X / lengthscale, X2 / lengthscale
    Xs, X2s = np.sum(np.square(X), axis=-1), np.sum(np.square(X2), axis=-1)
    r2 = -2 * np.tensordot(X, X2, [[-1], [-1]]) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
    r = np.sqrt(np.maximum(r2, 1e-36))
    sqrt5 = np.sqrt(5.0)
    return (1.0 + sqrt5 * r + 5.0 / 3.0 * np.square(r)) * np.exp(-sqrt5 * r)


class SpatioTemporalMatern52(object):
    """priors.py:1087-1181: M inducing points z, state = M copies of the Matern-5/2 temporal state."""
    def __init__(self, variance=1.0, lengthscale_time=1.0, lengthscale_space=1.0, z=None, fixed_grid=False):
        self.hyp = np.array([variance, lengthscale_time, lengthscale_space], dtype=float)
        if z is None:
            z = np.linspace(-3., 3., num=15)                                              # :1099
        self.z = np.asarray(z, dtype=float).reshape(-1, 1)
        self.M = self.z.shape[0]
        self.state_dim = 3 * self.M
        self.fixed_grid = fixed_grid

    def kernel_to_state_space(self):
        var, ell_time, ell_space = self.hyp
        Kmm = matern52_kernel(self.z, self.z, ell_space)                                  # :1129
        kappa = 5.0 / 3.0 * var / ell_time ** 2.0
        Pinf_time = np.array([[var, 0.0, -kappa],
                              [0.0, kappa, 0.0],
                              [-kappa, 0.0, 25.0 * var / ell_time ** 4.0]])               # :1141-1144
        return None, np.kron(Kmm, Pinf_time)                                              # :1145

    def spatial_weights(self, r):
        """Kx = Kxz Kzz^-1 (priors.py:1154-1159); r is one step's spatial location(s)."""
        if self.fixed_grid:
            return np.eye(self.M)
        ell_space = self.hyp[2]
        Kzz = matern52_kernel(self.z, self.z, ell_space)
        Kxz = matern52_kernel(np.asarray(r, dtype=float).reshape(-1, 1), self.z, ell_space)
        return sl.cho_solve(sl.cho_factor(Kzz), Kxz.T).T

    def measurement_model(self, r):
        return np.kron(self.spatial_weights(r), np.array([[1.0, 0.0, 0.0]]))              # :1160

    def state_transition(self, dt):
        return np.kron(np.eye(self.M), matern52_transition(dt, self.hyp[1]))              # :1180


def matern32_kernel(X, X2, lengthscale):
    """kernels.py:29-36 + 72-75 with utils.py:626-654."""
    X, X2 = X / lengthscale, X2 / lengthscale
    Xs, X2s = np.sum(np.square(X), axis=-1), np.sum(np.square(X2), axis=-1)
    r2 = -2 * np.tensordot(X, X2, [[-1], [-1]]) + Xs.reshape(-1, 1) + X2s.reshape(1, -1)
    r = np.sqrt(np.maximum(r2, 1e-36))
    sqrt3 = np.sqrt(3.0)
    return (1.0 + sqrt3 * r) * np.exp(-sqrt3 * r)


class SpatialMatern52(SpatioTemporalMatern52):
    """priors.py:1184-1267: the spatio-temporal prior with ONE lengthscale shared by time and space."""
    def __init__(self, variance=1.0, lengthscale=1.0, z=None, fixed_grid=False):
        SpatioTemporalMatern52.__init__(self, variance, lengthscale, lengthscale, z=z, fixed_grid=fixed_grid)


class SpatialMatern32(object):
    """priors.py:1270-1344: M copies of the Matern-3/2 temporal state, Matern-3/2 spatial kernel, shared lengthscale."""
    def __init__(self, variance=1.0, lengthscale=1.0, z=None, fixed_grid=False):
        self.hyp = np.array([variance, lengthscale], dtype=float)
        if z is None:
            z = np.linspace(-3., 3., num=15)                                              # :1282
        self.z = np.asarray(z, dtype=float).reshape(-1, 1)
        self.M = self.z.shape[0]
        self.state_dim = 2 * self.M
        self.fixed_grid = fixed_grid

    def kernel_to_state_space(self):
        var, ell = self.hyp
        Kmm = matern32_kernel(self.z, self.z, ell)                                        # :1302
        Pinf_time = np.array([[var, 0.0], [0.0, 3.0 * var / ell ** 2.0]])                 # :1311-1312
        return None, np.kron(Kmm, Pinf_time)                                              # :1313

    def spatial_weights(self, r):
        if self.fixed_grid:
            return np.eye(self.M)
        ell = self.hyp[1]
        Kzz = matern32_kernel(self.z, self.z, ell)
        Kxz = matern32_kernel(np.asarray(r, dtype=float).reshape(-1, 1), self.z, ell)
        return sl.cho_solve(sl.cho_factor(Kzz), Kxz.T).T                                  # :1325-1327

    def measurement_model(self, r):
        return np.kron(self.spatial_weights(r), np.array([[1.0, 0.0]]))                   # :1328

    def state_transition(self, dt):
        lam = np.sqrt(3.0) / self.hyp[1]                                                  # :1341-1343
        A_time = np.exp(-dt * lam) * (dt * np.array([[lam, 1.0], [-lam ** 2.0, -lam]]) + np.eye(2))
        return np.kron(np.eye(self.M), A_time)


class Matern52FixedVar(Matern52):
    """priors.py:1552-1600: Matern-5/2 whose variance is a constant, not a hyperparameter (hyp = lengthscale)."""
    def __init__(self, variance=1.0, lengthscale=1.0):
        Matern52.__init__(self, variance, lengthscale)


class SubbandExponentialFixedVar(SubbandMatern12):
    """priors.py:1498-1549: SubbandMatern12 with a constant variance (hyp = [lengthscale, radial_frequency])."""
    def __init__(self, variance=1.0, lengthscale=1.0, radial_frequency=1.0):
        SubbandMatern12.__init__(self, variance, lengthscale, radial_frequency)
