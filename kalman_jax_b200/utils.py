"""Host-side helpers mirroring /root/reference/kalmanjax/utils.py (NumPy; nothing here is on the GPU path).

softplus / softplus_inv / softplus_list (utils.py:41-84), input_admin (87-114), test_input_admin (117-197),
gauss_hermite (455-463), symmetric_cubature_third_order / _fifth_order (466-511, one-dimensional rules).
"""
import numpy as np
from numpy.polynomial.hermite import hermgauss

pi = 3.141592653589793


def softplus(x_):
    x_ = np.asarray(x_, dtype=np.float64)
    return np.log(1 + np.exp(-np.abs(x_))) + np.maximum(x_, 0)


def softplus_inv(x_):
    if x_ is None:
        return x_
    x_ = np.asarray(x_, dtype=np.float64)
    return np.log(1 - np.exp(-np.abs(x_))) + np.maximum(x_, 0)


def softplus_list(x_):
    return [softplus(x) for x in x_]


def gauss_hermite(dim=1, num_quad_pts=20):
    if dim != 1:
        raise NotImplementedError("only one-dimensional (f = 1) cubature is built")
    x, w = hermgauss(num_quad_pts)
    return np.sqrt(2) * x, w * pi ** (-0.5)


def symmetric_cubature_third_order(dim=1, kappa=None):
    if dim != 1:
        raise NotImplementedError("only one-dimensional (f = 1) cubature is built")
    return np.array([0., 1., -1.]), np.array([0., 0.5, 0.5])


def symmetric_cubature_fifth_order(dim=1):
    if dim != 1:
        raise NotImplementedError("only one-dimensional (f = 1) cubature is built")
    # the reference's 4-digit literals (utils.py:510-511); weights sum to 1.0001 on purpose
    return np.array([0., 1.7321, -1.7321]), np.array([0.6667, 0.1667, 0.1667])


def input_admin(t, y, r):
    """Order the inputs (utils.py:87-114)."""
    t, y = np.asarray(t, dtype=np.float64), np.asarray(y, dtype=np.float64)
    assert t.shape[0] == y.shape[0]
    if t.ndim < 2:
        t = np.expand_dims(t, 1)
    if y.ndim < 2:
        y = np.expand_dims(y, 1)
    if r is None:
        r = np.nan * t
    r = np.asarray(r, dtype=np.float64)
    if r.ndim < 2:
        r = np.expand_dims(r, 1)
    ind = np.argsort(t[:, 0], axis=0)
    t_train, y_train, r_train = t[ind, ...], y[ind, ...], r[ind, ...]
    dt_train = np.concatenate([np.array([0.0]), np.diff(t_train[:, 0])])
    return t_train, y_train, r_train, dt_train


def test_input_admin(t, y, r, t_test, y_test, r_test):
    """Merge + sort train and test inputs, index them, build the mask (utils.py:117-197)."""
    t, y = np.asarray(t, dtype=np.float64), np.asarray(y, dtype=np.float64)
    assert t.shape[0] == y.shape[0]
    if t.ndim < 2:
        t = np.expand_dims(t, 1)
    if y.ndim < 2:
        y = np.expand_dims(y, 1)
    if r is None:
        r = np.nan * t
    r = np.asarray(r, dtype=np.float64)
    if r.ndim < 2:
        r = np.expand_dims(r, 1)
    ind = np.argsort(t[:, 0], axis=0)
    t_train, y_train, r_train = t[ind, ...], y[ind, ...], r[ind, ...]
    if t_test is None:
        t_test = np.empty((1,) + t_train.shape[1:]) * np.nan
        r_test = np.empty((1,) + t_train.shape[1:]) * np.nan
    else:
        t_test = np.asarray(t_test, dtype=np.float64)
        if t_test.ndim < 2:
            t_test = np.expand_dims(t_test, 1)
        test_sort_ind = np.argsort(t_test[:, 0], axis=0)
        t_test = t_test[test_sort_ind, ...]
        if y_test is not None:
            y_test = np.asarray(y_test, dtype=np.float64)[test_sort_ind, ...].reshape((-1,) + y.shape[1:])
        if r_test is not None:
            r_test = np.asarray(r_test, dtype=np.float64)
            if r_test.ndim < 2:
                r_test = np.expand_dims(r_test, 1)
            r_test = r_test[test_sort_ind, ...]
        else:
            r_test = np.nan * t_test
    if not (t_test.shape[1] == t_train.shape[1]):
        t_test = np.concatenate([t_test[:, 0][:, None],
                                 np.nan * np.empty([t_test.shape[0], t_train.shape[1] - 1])], axis=1)
    t_train_test = np.concatenate([t_train, t_test])
    keep_ind = ~np.isnan(t_train_test[:, 0])
    t_train_test = t_train_test[keep_ind, ...]
    if r_test.shape[1] != r_train.shape[1]:
        r_test_nan = np.nan * np.zeros([r_test.shape[0], r_train.shape[1]])
    else:
        r_test_nan = r_test
    r_train_test = np.concatenate([r_train, r_test_nan])[keep_ind, ...]
    t_ind = np.argsort(t_train_test[:, 0])
    t_all, r_all = t_train_test[t_ind], r_train_test[t_ind]
    reverse_ind = np.argsort(t_ind)
    n_train = t_train.shape[0]
    train_id, test_id = reverse_ind[:n_train], reverse_ind[n_train:]
    y_all = np.nan * np.zeros([t_all.shape[0], y_train.shape[1]])
    y_all[train_id] = y_train
    if y_test is not None:
        y_all[test_id] = y_test
    mask = np.ones_like(y_all, dtype=bool)
    mask[train_id] = False
    dt_all = np.concatenate([np.array([0.0]), np.diff(t_all[:, 0])])
    return (t_all, y_all, r_all, r_test, dt_all, np.array(train_id, dtype=np.int64),
            np.array(test_id, dtype=np.int64), mask)
