"""Host side of the priors: the product's mirror (kalman_jax_b200/priors.py: Pinf, H, closed-form A(dt), block
description) against the oracle's restatement of the reference (oracle/priors.py), for every prior both implement."""
import numpy as np
import pytest

import oracle
from kalman_jax_b200 import priors as P

Z = np.linspace(-3.0, 3.0, 7)
ZOO = [
    ("Matern12", dict(variance=1.3, lengthscale=2.0)),
    ("Matern32", dict(variance=1.3, lengthscale=2.0)),
    ("Matern52", dict(variance=1.3, lengthscale=2.0)),
    ("Matern72", dict(variance=1.3, lengthscale=2.0)),
    ("Cosine", dict(frequency=[0.7])),
    ("Periodic", dict(variance=0.8, lengthscale=1.5, period=3.0)),
    ("QuasiPeriodicMatern12", dict(variance=1.2, lengthscale_periodic=1.5, period=20.0, lengthscale_matern=40.0)),
    ("QuasiPeriodicMatern32", dict(variance=1.2, lengthscale_periodic=1.5, period=20.0, lengthscale_matern=40.0)),
    ("SubbandMatern12", dict(variance=0.7, lengthscale=3.0, radial_frequency=1.3)),
    ("SubbandMatern32", dict(variance=1.1, lengthscale=4.0, radial_frequency=0.4)),
    ("SubbandMatern52", dict(variance=0.9, lengthscale=3.0, radial_frequency=0.7)),
    ("SpatioTemporalMatern52", dict(variance=0.5, lengthscale_time=0.8, lengthscale_space=0.5, z=Z)),
    ("SpatialMatern52", dict(variance=0.4, lengthscale=0.6, z=Z)),
    ("SpatialMatern32", dict(variance=0.4, lengthscale=0.8, z=Z)),
]
BLOCK_SIZE = {1: 2, 2: 3, 3: 4, 4: 1, 5: 2, 6: 4, 7: 6}


def _same(a, b):
    """equal up to the softplus(softplus_inv(.)) round trip of the hyperparameters, which the mirror (like the reference)
    stores in softplus-inverse space and the oracle does not"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-14 * max(1.0, float(np.max(np.abs(b)))))


@pytest.mark.parametrize("name,kw", ZOO, ids=[z[0] for z in ZOO])
def test_prior_mirror_matches_oracle(name, kw):
    mine, ref = getattr(P, name)(**kw), getattr(oracle.priors, name)(**kw)
    Hr, Pr = ref.kernel_to_state_space()
    _, _, _, Hm, Pm = mine.kernel_to_state_space()
    assert _same(Pm, Pr)
    spatial = hasattr(mine, "spatial_weights")
    if spatial:
        r = np.array([0.37])
        assert np.allclose(mine.measurement_model(r), ref.measurement_model(r), rtol=0, atol=1e-15)
    else:
        assert _same(Hm, Hr)
    for dt in (0.0, 0.013, 0.7, 5.0):
        assert _same(mine.state_transition(dt), ref.state_transition(dt))
    # the block description handed to the kernels adds up to the state and matches A(dt)'s sparsity
    blocks = mine.blocks()
    d = sum(BLOCK_SIZE[k] * c for k, c, _, _ in blocks)
    assert d == Pm.shape[0] == mine.state_transition(0.3).shape[0]
    A = mine.state_transition(0.3)
    mask = np.zeros_like(A, dtype=bool)
    o = 0
    for k, c, _, _ in blocks:
        for _ in range(c):
            b = BLOCK_SIZE[k]
            mask[o:o + b, o:o + b] = True
            o += b
    assert np.all(A[~mask] == 0.0)


def test_sum_stacks_components():
    parts = [("Matern52", dict(variance=2.0, lengthscale=3.0)), ("QuasiPeriodicMatern32", ZOO[7][1]), ("Matern12", ZOO[0][1])]
    mine = P.Sum([getattr(P, n)(**k) for n, k in parts])
    ref = oracle.priors.Sum([getattr(oracle.priors, n)(**k) for n, k in parts])
    Hr, Pr = ref.kernel_to_state_space()
    _, _, _, Hm, Pm = mine.kernel_to_state_space()
    assert _same(Pm, Pr) and _same(Hm, Hr)
    assert _same(mine.state_transition(0.4), ref.state_transition(0.4))
    assert len(mine.blocks()) == 1 + 7 + 1
