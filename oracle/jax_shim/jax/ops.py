import numpy as _np

index = _np.s_


def index_add(x, idx, y):
    out = _np.array(x, copy=True)
    out[idx] += y
    return out


def index_update(x, idx, y):
    out = _np.array(x, copy=True)
    out[idx] = y
    return out
