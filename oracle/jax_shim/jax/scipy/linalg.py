import numpy as _np
import scipy.linalg as _sl

expm = _sl.expm


def cholesky(a, lower=False):
    # jax.scipy.linalg.cholesky: clean triangular factor (other triangle zero), no exception on
    # non-PD input (NaN propagates instead).
    a = _np.atleast_2d(_np.asarray(a, dtype=float))
    try:
        l = _np.linalg.cholesky(a)
    except _np.linalg.LinAlgError:
        l = _np.full_like(a, _np.nan)
    return l if lower else l.T.copy()


def cho_factor(a, lower=False):
    return cholesky(a, lower=lower), lower


def cho_solve(c_and_lower, b):
    c, lower = c_and_lower
    b = _np.asarray(b, dtype=float)
    if not _np.all(_np.isfinite(c)):
        return _np.full(_np.broadcast_shapes(b.shape, (c.shape[0],) + b.shape[1:]), _np.nan)
    return _sl.cho_solve((c, lower), b, check_finite=False)
