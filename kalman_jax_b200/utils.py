"""Host-side helpers mirroring /root/reference/kalmanjax/utils.py (NumPy; nothing here is on the GPU path).

softplus / softplus_inv / softplus_list (utils.py:41-84), input_admin / test_input_admin (own implementation of the
contract of utils.py:87-197: one stable merge of the training and test time columns),
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
    """utils.py:418-463: Gauss-Hermite points scaled for a unit Gaussian; dim > 1 is the product rule, x [dim, Q^dim]
    with the LAST coordinate varying fastest (the order of itertools.product), w [Q^dim]."""
    x, w = hermgauss(num_quad_pts)
    if dim == 1:
        return np.sqrt(2) * x, w * pi ** (-0.5)
    grid = np.stack(np.meshgrid(*([x] * dim), indexing="ij"), axis=0).reshape(dim, -1)
    wts = np.prod(np.stack(np.meshgrid(*([w] * dim), indexing="ij"), axis=0).reshape(dim, -1), axis=0)
    return np.sqrt(2) * grid, wts * pi ** (-0.5 * dim)


def symmetric_cubature_third_order(dim=1, kappa=None):
    """utils.py:466-503 with kappa = 0: the origin (weight 0) and +- sqrt(dim) along every axis; dims 1-3 carry the
    reference's 4-digit literals."""
    if kappa not in (None, 0):
        raise NotImplementedError("only kappa = 0 (the reference's default) is built")
    if dim == 1:
        return np.array([0., 1., -1.]), np.array([0., 0.5, 0.5])
    u, wm = {2: (1.4142, 0.25), 3: (1.7321, 0.1667)}.get(dim, (np.sqrt(dim), 1.0 / (2 * dim)))
    pts = np.concatenate([np.zeros([dim, 1]), u * np.eye(dim), -u * np.eye(dim)], axis=1)
    wts = np.concatenate([[0.], np.full(2 * dim, wm)])
    return pts, (wts if dim in (2, 3) else wts[None, :])      # the reference's general branch returns weights [1, 2 dim + 1]


def symmetric_cubature_fifth_order(dim=1):
    """utils.py:505-516: the reference's 4-digit literals (the weights do not sum to one exactly; kept on purpose).
    Dimensions 1 and 2 (scalar sites and the two-latent HeteroscedasticNoise); the reference's tables for 3 and 6 latent
    functions belong to likelihoods that are not built."""
    u = 1.7321
    if dim == 1:
        return np.array([0., u, -u]), np.array([0.6667, 0.1667, 0.1667])
    if dim == 2:
        return (np.array([[0., u, -u, 0., 0., u, -u, u, -u], [0., 0., 0., u, -u, u, -u, -u, u]]),
                np.array([0.4444, 0.1111, 0.1111, 0.1111, 0.1111, 0.0278, 0.0278, 0.0278, 0.0278]))
    raise NotImplementedError("fifth-order symmetric rule: dimensions 1 and 2 are built")


def _as_columns(x, like=None):
    """float64 array with the sample axis first and at least one trailing axis ([N] -> [N,1]); None -> NaNs like `like`."""
    if x is None:
        return np.full_like(like, np.nan)
    x = np.asarray(x, dtype=np.float64)
    return x[:, None] if x.ndim < 2 else x


def _time_ordered(t, *others):
    """Rows of t (and of the arrays that travel with it) by increasing first column."""
    order = np.argsort(t[:, 0], axis=0)
    return (t[order],) + tuple(o[order] for o in others)


def _steps(t_sorted):
    """dt[0] = 0, dt[k] = t[k] - t[k-1] (the filter's first step starts AT the first input)."""
    dt = np.zeros(t_sorted.shape[0])
    dt[1:] = t_sorted[1:, 0] - t_sorted[:-1, 0]
    return dt


def input_admin(t, y, r):
    """Training inputs as the filter wants them (the contract of the reference's utils.py:87-114): 2-D float64 arrays
    ordered by time, missing spatial inputs as NaN columns, and the step sizes."""
    t, y = _as_columns(t), _as_columns(y)
    assert t.shape[0] == y.shape[0]
    t, y, r = _time_ordered(t, y, _as_columns(r, like=t))
    return t, y, r, _steps(t)


def test_input_admin(t, y, r, t_test, y_test, r_test):
    """Training and test inputs on ONE time grid (the contract of the reference's utils.py:117-197).

    Returns (t_all, y_all, r_all, r_test, dt_all, train_id, test_id, mask): the merged grid ordered by time, the rows the
    (time-ordered) training / test points occupy on it, y_all holding the training targets (and the test targets when
    given, for the NLPD) with NaN elsewhere, and mask = True on the rows that are NOT training data (predict-only steps).
    """
    t_tr, y_tr, r_tr, _ = input_admin(t, y, r)
    n_tr, n_tcol, n_rcol = t_tr.shape[0], t_tr.shape[1], r_tr.shape[1]
    if t_test is None:                                          # nothing to predict: the grid is the training grid
        t_te = np.empty((0, n_tcol))
        r_test = np.full((1, n_tcol), np.nan)
        r_te = np.empty((0, n_rcol))
        y_te = None
    else:
        t_te = _as_columns(t_test)
        r_te = _as_columns(r_test, like=t_te)
        if y_test is None:
            t_te, r_te = _time_ordered(t_te, r_te)
            y_te = None
        else:
            t_te, r_te, y_te = _time_ordered(t_te, r_te, np.asarray(y_test, dtype=np.float64))
            y_te = y_te.reshape((-1,) + y_tr.shape[1:])
        r_test = r_te
        if t_te.shape[1] != n_tcol:                             # only the time column of the test inputs is used
            t_te = np.concatenate([t_te[:, :1], np.full((t_te.shape[0], n_tcol - 1), np.nan)], axis=1)
        if r_te.shape[1] != n_rcol:                             # different spatial dimensionality: handled by the caller
            r_te = np.full((r_te.shape[0], n_rcol), np.nan)
        known = ~np.isnan(t_te[:, 0])                           # test rows without a time are dropped
        t_te, r_te = t_te[known], r_te[known]
        if y_te is not None:
            y_te = y_te[known]
    # one stable ordering of the stacked time columns; rank[i] = row of stacked point i on the merged grid
    t_stack, r_stack = np.concatenate([t_tr, t_te]), np.concatenate([r_tr, r_te])
    order = np.argsort(t_stack[:, 0], kind="stable")
    rank = np.empty(order.size, dtype=np.int64)
    rank[order] = np.arange(order.size)
    train_id, test_id = rank[:n_tr], rank[n_tr:]
    t_all, r_all = t_stack[order], r_stack[order]
    y_all = np.full((order.size, y_tr.shape[1]), np.nan)
    y_all[train_id] = y_tr
    if y_te is not None:
        y_all[test_id] = y_te
    mask = np.ones(y_all.shape, dtype=bool)
    mask[train_id] = False
    return t_all, y_all, r_all, r_test, _steps(t_all), train_id, test_id, mask
