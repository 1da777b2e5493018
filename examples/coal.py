"""The reference's experiments/coal/coal.py:15-130 (log-Gaussian Cox process on the coal-mining data, 10-fold CV) with
only the imports changed and the plotting dropped: 250 Adam steps on the hyperparameters, every step one SDEGP.run()
(filter energy + gradient, then smoother + site update), then the test NLPD of the held-out fold.

    python examples/coal.py [method] [fold]        (method 0 = EEP(power=1), ... as in the reference)

KAT K7: the reference's stored result for method 0, fold 0 is 0.8147043424175697 (experiments/coal/output/0_0_nlpd.txt).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kalman_jax_b200 import SDEGP, optimizers, priors, likelihoods  # noqa: E402
from kalman_jax_b200 import approximate_inference as approx_inf  # noqa: E402
from kalman_jax_b200.utils import softplus_list  # noqa: E402


def main(method=0, fold=0, steps=250, verbose=True):
    data = os.path.join(ROOT, "tests", "data")
    cvind = np.loadtxt(os.path.join(data, "cvind.csv")).astype(int)
    nt = np.floor(cvind.shape[0] / 10).astype(int)                      # 10-fold cross-validation
    cvind = np.reshape(cvind[:10 * nt], (10, nt))
    D = np.loadtxt(os.path.join(data, "binned.csv"))
    x, y = D[:, 0:1], D[:, 1:]
    ind_test = cvind[fold, :]
    ind_train = np.setdiff1d(cvind, ind_test)
    x_train, x_test, y_train, y_test = x[ind_train, ...], x[ind_test, ...], y[ind_train, ...], y[ind_test, ...]

    prior = priors.Matern52(variance=1.0, lengthscale=1.0)
    lik = likelihoods.Poisson()
    inf_method = [lambda: approx_inf.EEP(power=1), lambda: approx_inf.EEP(power=0.5), lambda: approx_inf.EKS(),
                  lambda: approx_inf.UEP(power=1), lambda: approx_inf.UEP(power=0.5), lambda: approx_inf.UKS(),
                  lambda: approx_inf.GHEP(power=1), lambda: approx_inf.GHEP(power=0.5), lambda: approx_inf.GHKS(),
                  lambda: approx_inf.EP(power=1, intmethod='UT'), lambda: approx_inf.EP(power=0.5, intmethod='UT'),
                  lambda: approx_inf.EP(power=0.01, intmethod='UT'), lambda: approx_inf.EP(power=1, intmethod='GH'),
                  lambda: approx_inf.EP(power=0.5, intmethod='GH'), lambda: approx_inf.EP(power=0.01, intmethod='GH'),
                  lambda: approx_inf.VI(intmethod='UT'), lambda: approx_inf.VI(intmethod='GH')][method]()
    model = SDEGP(prior=prior, likelihood=lik, t=x_train, y=y_train, approx_inf=inf_method)

    opt_init, opt_update, get_params = optimizers.adam(step_size=2.5e-1)
    opt_state = opt_init([model.prior.hyp, model.likelihood.hyp])

    def gradient_step(i, state, mod):
        params = get_params(state)
        mod.prior.hyp = params[0]
        mod.likelihood.hyp = params[1]
        neg_log_marg_lik, gradients = mod.run()                          # grad(Filter) + Smoother
        if verbose and i % 25 == 0:
            prior_params = softplus_list(params[0])
            print('iter %2d: var_f=%1.2f len_f=%1.2f, nlml=%2.2f' % (i, prior_params[0], prior_params[1], neg_log_marg_lik))
        return opt_update(i, gradients, state)

    t0 = time.time()
    for j in range(steps):
        opt_state = gradient_step(j, opt_state, model)
    t1 = time.time()
    nlpd = model.negative_log_predictive_density(t=x_test, y=y_test)
    if verbose:
        print('optimisation time: %2.2f secs' % (t1 - t0))
        print('NLPD: %.16f   (reference output/0_0_nlpd.txt for method 0, fold 0: 0.8147043424175697)' % nlpd)
    return nlpd


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
