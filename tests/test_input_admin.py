"""Host-side input preparation (kalman_jax_b200/utils.py) against the oracle's restatement of utils.py:87-197."""
import numpy as np
import pytest

import oracle.sde_gp as O
from kalman_jax_b200 import utils as U


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("with_r", [False, True])
def test_input_admin_orders_by_time(with_r):
    rng = np.random.default_rng(0)
    t, y = rng.uniform(0, 10, 57), rng.standard_normal(57)
    r = rng.uniform(-1, 1, (57, 2)) if with_r else None
    for got, want in zip(U.input_admin(t, y, r), O.input_admin(t, y, r)):
        assert _eq(got, want)


@pytest.mark.parametrize("case", ["plain", "with_y", "with_r", "none", "nan_test", "r_dim_differs"])
def test_test_input_admin_matches_the_oracle(case):
    rng = np.random.default_rng(1)
    n, m = 41, 17
    t, y = rng.uniform(0, 10, n), rng.standard_normal(n)
    tt, yt, r, rt = rng.uniform(-1, 11, m), None, None, None
    if case == "with_y":
        yt = rng.standard_normal(m)
    if case == "with_r":
        r, rt = rng.uniform(-1, 1, (n, 1)), rng.uniform(-1, 1, (m, 1))
    if case == "none":
        tt = None
    if case == "nan_test":
        tt = tt.copy(); tt[[3, 8]] = np.nan
    if case == "r_dim_differs":
        r, rt = rng.uniform(-1, 1, (n, 1)), rng.uniform(-1, 1, (m, 2))
    t, y, r, _ = O.input_admin(t, y, r)                 # as SDEGP holds them (2-D, time-ordered)
    got = U.test_input_admin(t, y, r, tt, yt, rt)
    want = O.test_input_admin(t, y, r, tt, yt, rt)
    names = ["t_all", "y_all", "r_all", "r_test", "dt_all", "train_id", "test_id", "mask"]
    for name, g, w in zip(names, got, want):
        if name == "r_test" and case == "none":
            continue                                   # placeholder row, never read
        assert _eq(g, w), name


def test_ties_keep_training_points_first():
    t = np.array([0.0, 1.0, 2.0]); y = np.array([1.0, 2.0, 3.0])
    out = U.test_input_admin(t, y, None, np.array([1.0, 2.0]), None, None)
    assert out[0].shape == (5, 1) and np.array_equal(out[4], [0.0, 1.0, 0.0, 1.0, 0.0])
    assert list(out[5]) == [0, 1, 3] and list(out[6]) == [2, 4]
