"""Host-side cubature tables of the mirror (kalman_jax_b200/utils.py) against the oracle's restatement of
utils.py:418-516 — dimensions 1 and 2 — and the inference-method descriptors built on them (CPU only)."""
import numpy as np
import pytest

from oracle import cubature as ocub


def _utils():
    from kalman_jax_b200 import utils
    return utils


@pytest.mark.parametrize("n", [3, 10, 20])
def test_gauss_hermite_tables(n):
    U = _utils()
    x, w = U.gauss_hermite(1, n)
    xo, wo = ocub.gauss_hermite(n)
    assert np.array_equal(x, xo) and np.array_equal(w, wo)
    x2, w2 = U.gauss_hermite(2, n)
    xo2, wo2 = ocub.gauss_hermite_2d(n)
    assert x2.shape == (2, n * n) and np.array_equal(x2, xo2) and np.allclose(w2, wo2, rtol=1e-15, atol=0)
    # the product rule in itertools.product order: the LAST coordinate varies fastest; weights multiply
    assert np.array_equal(x2[1, :n], x) and np.all(x2[0, :n] == x[0])
    assert abs(w2.sum() - 1.0) < 1e-12


def test_symmetric_rules_carry_the_reference_literals():
    U = _utils()
    for mine, theirs in ((U.symmetric_cubature_third_order(1), ocub.symmetric_cubature_third_order()),
                         (U.symmetric_cubature_fifth_order(1), ocub.symmetric_cubature_fifth_order()),
                         (U.symmetric_cubature_third_order(2), ocub.symmetric_cubature_third_order_2d()),
                         (U.symmetric_cubature_fifth_order(2), ocub.symmetric_cubature_fifth_order_2d())):
        assert np.array_equal(np.asarray(mine[0]), np.asarray(theirs[0]))
        assert np.array_equal(np.asarray(mine[1]), np.asarray(theirs[1]))
    assert abs(U.symmetric_cubature_fifth_order(1)[1].sum() - 1.0001) < 1e-12      # 4-digit literals, kept on purpose
    x3, w3 = U.symmetric_cubature_third_order(3)
    assert x3.shape == (3, 7) and w3.shape == (7,) and x3[0, 1] == 1.7321
    with pytest.raises(NotImplementedError):
        U.symmetric_cubature_fifth_order(6)


def test_descriptors_select_rule_and_cub_rule():
    """The reference's names resolve to the four rules with the pinned arguments, and the two-dimensional rule a
    two-latent SLR update needs is announced through kjx_model_t.cub_rule (0 Gauss-Hermite product, 3 / 5 symmetric)."""
    from kalman_jax_b200 import _lib, approximate_inference as ai, likelihoods, priors
    from kalman_jax_b200.sde_gp import _Model
    assert ai.UKS().kjx_code == _lib.INF_SLEP and ai.UKS().power == 0 and ai.UKS().intmethod == "UT"
    assert ai.GHKS(damping=0.3).power == 0 and ai.GHKS(damping=0.3).damping == 0.3 and ai.GHKS().name.endswith("(GHKS)")
    assert ai.EKF().kjx_code == _lib.INF_EEP and ai.EKF().power == 0 and issubclass(ai.EKF, ai.ExtendedKalmanSmoother)
    assert ai.PEP(power=0.5).power == 0.5 and ai.CVI(damping=0.2).kjx_code == _lib.INF_VI
    with pytest.raises(TypeError):
        ai.UKS(power=0.5)            # the reference's UKS constructor has no `power` either
    with pytest.raises(NotImplementedError):
        ai.EP(intmethod="Simpson")
    prior = priors.Independent([priors.Matern32(1.0, 1.0), priors.Matern32(1.0, 1.0)])
    lik = likelihoods.HeteroscedasticNoise()
    for rule, want in ((ai.SLEP(), 0), (ai.UKS(), 5), (ai.SLEP(intmethod="UT3"), 3), (ai.GHEP(num_cub_pts=9), 0)):
        m = _Model(prior, lik, rule, r0=np.zeros(1))
        assert m.c.cub_rule == want and m.c.func_dim == 2 and m.c.method == _lib.INF_SLEP
        assert m.c.n_cub == len(np.atleast_1d(rule.cubature_func(1)[0]))
