"""The reference's notebooks/regression.py:23-83 with only the imports changed: Matern-5/2 GP regression on an unevenly
spaced series, hyperparameters fitted by Adam on the filter energy, then NLPD and posterior prediction.  Needs a GPU."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kalman_jax_b200 import SDEGP, approximate_inference as approx_inf, likelihoods, optimizers, priors  # noqa: E402
from kalman_jax_b200.utils import softplus, softplus_list  # noqa: E402

pi = 3.141592653589793


def wiggly_time_series(x_):
    noise_var = 0.15  # true observation noise
    return np.cos(0.04 * x_ + 0.33 * pi) * np.sin(0.2 * x_) + np.sqrt(noise_var) * np.random.normal(0, 1, x_.shape)


np.random.seed(12345)
N = 1000
x = np.random.permutation(np.linspace(-25.0, 150.0, num=N) + 0.5 * np.random.randn(N))  # unevenly spaced
y = wiggly_time_series(x)
x_test = np.linspace(np.min(x) - 15.0, np.max(x) + 15.0, num=500)
y_test = wiggly_time_series(x_test)

prior = priors.Matern52(variance=1.0, lengthscale=5.0)
lik = likelihoods.Gaussian(variance=0.5)
model = SDEGP(prior=prior, likelihood=lik, t=x, y=y, approx_inf=approx_inf.EP(power=0.5))

opt_init, opt_update, get_params = optimizers.adam(step_size=5e-1)
opt_state = opt_init([model.prior.hyp, model.likelihood.hyp])


def gradient_step(i, state, mod):
    params = get_params(state)
    mod.prior.hyp = params[0]
    mod.likelihood.hyp = params[1]
    neg_log_marg_lik, gradients = mod.run()
    prior_params, lik_param = softplus_list(params[0]), softplus(params[1])
    print('iter %2d: var_f=%1.2f len_f=%1.2f var_y=%1.2f, nlml=%2.2f' %
          (i, prior_params[0], prior_params[1], lik_param, neg_log_marg_lik))
    return opt_update(i, gradients, state), neg_log_marg_lik


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    t0 = time.time()
    for j in range(iters):
        opt_state, nlml = gradient_step(j, opt_state, model)
    print('optimisation time: %2.2f secs' % (time.time() - t0))
    t0 = time.time()
    nlpd = model.negative_log_predictive_density(t=x_test, y=y_test)
    posterior_mean, posterior_cov = model.predict(t=np.linspace(np.min(x) - 20.0, np.max(x) + 20.0, 200))
    print('prediction time: %2.2f secs' % (time.time() - t0))
    print('test NLPD: %1.2f' % nlpd)
