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
