"""ORACLE (test infrastructure, not product): cubature tables for one-dimensional sites.

Restates /root/reference/kalmanjax/utils.py:418-463 (Gauss-Hermite through numpy `hermgauss`),
utils.py:466-474 (3rd-order symmetric rule, dim 1) and utils.py:505-511 (5th-order rule, dim 1 —
note the reference's 4-digit literals; the weights sum to 1.0001 and that is reproduced on purpose).
"""
import numpy as np
from numpy.polynomial.hermite import hermgauss

GH, UT3, UT5 = "GH", "UT3", "UT5"


def gauss_hermite(num_pts=20):
    # utils.py:455-463 with dim=1:  x = sqrt(2)*hermgauss.x ,  w = hermgauss.w / sqrt(pi)
    x, w = hermgauss(num_pts)
    return np.sqrt(2.0) * x, w * np.pi ** -0.5


def symmetric_cubature_third_order():
    # utils.py:474-476 (dim == 1, kappa == 0)
    return np.array([0.0, 1.0, -1.0]), np.array([0.0, 0.5, 0.5])


def symmetric_cubature_fifth_order():
    # utils.py:509-511 (dim == 1): literals exactly as written in the reference
    return np.array([0.0, 1.7321, -1.7321]), np.array([0.6667, 0.1667, 0.1667])


def make(intmethod="GH", num_cub_pts=20):
    """approximate_inference.py:25-34 — select the rule by name; returns (x[Q], w[Q])."""
    if intmethod == "GH":
        return gauss_hermite(num_cub_pts)
    if intmethod == "UT3":
        return symmetric_cubature_third_order()
    if intmethod in ("UT5", "UT"):
        return symmetric_cubature_fifth_order()
    raise NotImplementedError("integration method not recognised")


# ---- two-dimensional rules (func_dim = 2: HeteroscedasticNoise with the SLR family) --------------------------------
def gauss_hermite_2d(num_pts=20):
    """utils.py:418-463 with dim = 2: the product rule in itertools.product order, x [2, Q^2], w [Q^2]."""
    gx, gw = hermgauss(num_pts)
    x = np.array([[a, b] for a in gx for b in gx])
    w = np.array([a * b for a in gw for b in gw])
    return np.sqrt(2.0) * x.T, w * np.pi ** -1.0


def symmetric_cubature_third_order_2d():
    """utils.py:484-488 (dim == 2, kappa == 0): 4-digit literals as written."""
    return (np.array([[0., 1.4142, 0., -1.4142, 0.], [0., 0., 1.4142, 0., -1.4142]]),
            np.array([0., 0.25, 0.25, 0.25, 0.25]))


def symmetric_cubature_fifth_order_2d():
    """utils.py:512-516 (dim == 2): 4-digit literals as written."""
    return (np.array([[0., 1.7321, -1.7321, 0., 0., 1.7321, -1.7321, 1.7321, -1.7321],
                      [0., 0., 0., 1.7321, -1.7321, 1.7321, -1.7321, -1.7321, 1.7321]]),
            np.array([0.4444, 0.1111, 0.1111, 0.1111, 0.1111, 0.0278, 0.0278, 0.0278, 0.0278]))


def make_2d(intmethod="GH", num_cub_pts=20):
    """approximate_inference.py:25-34 evaluated at dim = 2."""
    if intmethod == "GH":
        return gauss_hermite_2d(num_cub_pts)
    if intmethod == "UT3":
        return symmetric_cubature_third_order_2d()
    if intmethod in ("UT5", "UT"):
        return symmetric_cubature_fifth_order_2d()
    raise NotImplementedError("integration method not recognised")
