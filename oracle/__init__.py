"""ORACLE — CPU restatement of the reference's filter / smoother / site-update path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import anything from here; the product
package (`kalman_jax_b200`) never does and fails loudly when its CUDA library is missing.

Parity status: pinned.  JAX cannot be installed here, so the reference's own sources are executed on
NumPy through `oracle/jax_shim` by `tests/golden/make_golden.py`; the recorded fp64 outputs
(tests/golden/*.npz, 30 cases) pin this restatement to <= 1e-10 relative
(tests/test_oracle_golden.py), and four of them reproduce values printed in the reference's executed
notebooks (797.76, 295.62, 469.46, 178.81).  What stays unpinned: the arithmetic of the un-vendored
jax==0.1.67 / jaxlib==0.1.47 kernels (LAPACK potrf/potrs, XLA's erf/erfc/lgamma), replaced here by
SciPy's.
"""
from . import cubature, likelihoods, approx_inf, priors, sde_gp, multi  # noqa: F401
from .sde_gp import SDEGP as _SDEGPScalar
from .multi import SDEGPMulti, Independent as _Independent, HeteroscedasticNoise as _HeteroscedasticNoise

# the multi-latent classes under the reference's names, next to the scalar ones
priors.Independent = _Independent
likelihoods.HeteroscedasticNoise = _HeteroscedasticNoise


def SDEGP(prior, likelihood, t, y, r=None, approx_inf=None):
    """The scalar-site oracle (sde_gp.py) or, for an `Independent` prior (func_dim > 1), the matrix-site one (multi.py)."""
    if isinstance(prior, _Independent):
        return SDEGPMulti(prior, likelihood, t, y, r, approx_inf)
    return _SDEGPScalar(prior, likelihood, t, y, r, approx_inf)
