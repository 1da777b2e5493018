import numpy as _np


def softplus(x):
    return _np.logaddexp(x, 0.0)
