"""Adam mirror of jax.experimental.optimizers.adam as the reference's drivers call it (notebooks/regression.py:47-73)."""
import numpy as np

from kalman_jax_b200 import optimizers


def test_adam_first_step_and_convergence_on_nested_params():
    init, update, get_params = optimizers.adam(step_size=0.1)
    x0 = [np.array([1.0, -2.0]), np.array(0.5)]                    # [prior hyp, likelihood hyp]
    state = init(x0)
    grad = lambda p: [2.0 * (p[0] - np.array([3.0, 1.0])), 2.0 * (p[1] + 1.0)]     # noqa: E731  quadratic bowl
    state1 = update(0, grad(get_params(state)), state)
    p1 = get_params(state1)
    # the first bias-corrected Adam step moves every coordinate by step * sign(g) (up to eps)
    assert np.allclose(p1[0], x0[0] - 0.1 * np.sign(grad(x0)[0]), atol=1e-6)
    assert np.allclose(p1[1], x0[1] - 0.1 * np.sign(grad(x0)[1]), atol=1e-6)
    assert np.array_equal(x0[0], np.array([1.0, -2.0]))           # the caller's arrays are not modified
    for i in range(1, 600):
        state1 = update(i, grad(get_params(state1)), state1)
    p = get_params(state1)
    assert np.allclose(p[0], [3.0, 1.0], atol=1e-3) and abs(float(p[1]) + 1.0) < 1e-3


def test_adam_matches_a_hand_rolled_reference_sequence():
    init, update, get_params = optimizers.adam(step_size=0.05, b1=0.8, b2=0.95, eps=1e-7)
    x = np.array([0.3, -0.7])
    state = init([x])
    m = v = np.zeros(2)
    rng = np.random.default_rng(0)
    for i in range(20):
        g = rng.standard_normal(2)
        state = update(i, [g], state)
        m = 0.2 * g + 0.8 * m
        v = 0.05 * g * g + 0.95 * v
        x = x - 0.05 * (m / (1 - 0.8 ** (i + 1))) / (np.sqrt(v / (1 - 0.95 ** (i + 1))) + 1e-7)
        assert np.allclose(get_params(state)[0], x, rtol=0, atol=1e-14)
