"""Adam with the calling convention of `jax.experimental.optimizers.adam` (jax 0.1.67), which the reference's
drivers use (notebooks/regression.py:47-73):

    opt_init, opt_update, get_params = optimizers.adam(step_size=5e-1)
    opt_state = opt_init([model.prior.hyp, model.likelihood.hyp])
    ...
    opt_state = opt_update(i, gradients, opt_state)
    params = get_params(opt_state)

Parameters and gradients are nested lists of NumPy arrays / scalars (the [prior hyp, likelihood hyp] structure of
SDEGP.run()).  Host-side: a handful of scalars per step.
"""
import numpy as np


def _map(f, *trees):
    if isinstance(trees[0], (list, tuple)):
        return [_map(f, *xs) for xs in zip(*trees)]
    return f(*[np.asarray(t, dtype=np.float64) for t in trees])


def adam(step_size, b1=0.9, b2=0.999, eps=1e-8):
    """x <- x - step * m_hat / (sqrt(v_hat) + eps), bias-corrected first and second moments (Kingma & Ba 2015)."""
    step = step_size if callable(step_size) else (lambda i: step_size)

    def init(x0):
        x0 = _map(lambda x: x.copy(), x0)
        return x0, _map(np.zeros_like, x0), _map(np.zeros_like, x0)

    def update(i, g, state):
        x, m, v = state
        m = _map(lambda g_, m_: (1 - b1) * g_ + b1 * m_, g, m)
        v = _map(lambda g_, v_: (1 - b2) * np.square(g_) + b2 * v_, g, v)
        x = _map(lambda x_, m_, v_: x_ - step(i) * (m_ / (1 - b1 ** (i + 1))) / (np.sqrt(v_ / (1 - b2 ** (i + 1))) + eps), x, m, v)
        return x, m, v

    def get_params(state):
        return state[0]

    return init, update, get_params
