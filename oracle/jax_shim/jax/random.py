import numpy as _np


def PRNGKey(seed):
    return int(seed)


def normal(key, shape=()):
    return _np.random.default_rng(int(key)).standard_normal(shape)
