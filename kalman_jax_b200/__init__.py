"""kalman_jax_b200 — B200 (sm_100a) implementation of kalman-jax's shared inference core: the Kalman
filter, the Rauch-Tung-Striebel smoother and the per-step site-update moments, behind the reference's
own SDEGP / approximate_inference API.

    from kalman_jax_b200 import SDEGP, priors, likelihoods, approximate_inference as approx_inf

The hot path is hand-written CUDA in csrc/ behind the C ABI of include/kjx.h (libkjx.so, loaded with
ctypes).  There is no CPU fallback: without the built library every compute call raises.
"""
from . import _lib, utils, priors, likelihoods, approximate_inference, optimizers  # noqa: F401
from .sde_gp import SDEGP, site_update  # noqa: F401

__all__ = ["SDEGP", "site_update", "priors", "likelihoods", "approximate_inference", "utils", "optimizers"]
