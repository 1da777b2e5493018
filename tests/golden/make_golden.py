#!/usr/bin/env python
"""Generate golden fixtures by executing the UNMODIFIED reference in place.

Test infrastructure only.  Runs in the build container (needs /root/reference); the fixtures it
writes (tests/golden/*.npz) are committed and are all the GPU box ever sees.

How: `oracle/jax_shim` (a NumPy/SciPy stand-in for the jax==0.1.67 API surface) is put first on
sys.path, then `/root/reference/kalmanjax` — so `from sde_gp import SDEGP` imports the reference's
own sources, which then run on NumPy in fp64.  Each case replays exactly what
`SDEGP.run()` does (`sde_gp.py:179-187`) minus the autodiff gradient: filter(store=True, sites)
-> smoother(site update), for a few iterations, then `predict()` / `negative_log_predictive_density()`
on held-out inputs.  Known-answer values printed in the reference's executed notebooks
 are asserted here at their printed precision.

usage:  python tests/golden/make_golden.py [case ...]
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KJX_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "kalmanjax"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

from sde_gp import SDEGP  # noqa: E402  (reference)
import approximate_inference as approx_inf  # noqa: E402  (reference)
import priors  # noqa: E402  (reference)
import likelihoods  # noqa: E402  (reference)

pi = 3.141592653589793
DATA = os.path.join(REF, "data")


# ----------------------------------------------------------------------------- data sets
def data_regression(N=1000):
    # notebooks/regression.py:16-29
    def wiggly_time_series(x_):
        noise_var = 0.15
        return (np.cos(0.04 * x_ + 0.33 * pi) * np.sin(0.2 * x_) +
                np.sqrt(noise_var) * np.random.normal(0, 1, x_.shape))
    np.random.seed(12345)
    x = np.random.permutation(np.linspace(-25.0, 150.0, num=N) + 0.5 * np.random.randn(N))
    y = wiggly_time_series(x)
    x_test = np.linspace(np.min(x) - 15.0, np.max(x) + 15.0, num=500)
    y_test = wiggly_time_series(x_test)
    return x, y, x_test[::10], y_test[::10]


def data_classification(N=1000):
    # notebooks/classification.py:17-27
    np.random.seed(99)
    x = 100 * np.random.rand(N)
    f = lambda x_: 6 * np.sin(pi * x_ / 10.0) / (pi * x_ / 10.0 + 1)  # noqa: E731
    y_ = f(x) + np.sqrt(0.05) * np.random.randn(x.shape[0])
    y = np.sign(y_)
    y[y == -1] = 0
    x_test = np.linspace(np.min(x) - 5.0, np.max(x) + 5.0, num=500)
    y_test = np.sign(f(x_test) + np.sqrt(0.05) * np.random.randn(x_test.shape[0]))
    y_test[y_test == -1] = 0
    return x, y, x_test[::10], y_test[::10]


def data_coal_200():
    # notebooks/log_gaussian_cox_process.py:17-23
    disaster_timings = pd.read_csv(os.path.join(DATA, "coal.txt"), header=None).values[:, 0]
    num_time_bins = 200
    x = np.linspace(min(disaster_timings), max(disaster_timings), num_time_bins).T
    y = np.histogram(disaster_timings, np.concatenate([[-1e10], x[:-1] + np.diff(x) / 2, [1e10]]))[0][:, None]
    return x, y.astype(float)


def data_coal_333():
    # experiments/coal/binned.csv (t, count), used by experiments/coal/coal.py
    D = np.loadtxt(os.path.join(REF, "kalmanjax", "experiments", "coal", "binned.csv"))
    return D[:, 0], D[:, 1:2]


def data_aircraft(n_keep=160):
    # experiments/aircraft/aircraft.py:19-35 (daily bins), thinned to a small fixture
    from datetime import date
    acc = pd.read_csv(os.path.join(DATA, "aircraft_accidents.txt"), sep="-", header=None).values
    xx = np.zeros([acc.shape[0], 1])
    for j in range(acc.shape[0]):
        xx[j] = date.toordinal(date(acc[j, 0], acc[j, 1], acc[j, 2])) + 366
    x_min, x_max = np.floor(np.min(xx)), np.ceil(np.max(xx))
    x = np.linspace(x_min, x_max, num=int((x_max - x_min) + 1))
    x = np.concatenate([np.min(x) - np.linspace(61, 1, num=61), x])
    y, _ = np.histogram(xx, np.concatenate([[-1e10], x[1:] - np.diff(x) / 2, [1e10]]))
    rng = np.random.RandomState(123)
    keep = np.sort(rng.permutation(2000)[:n_keep]) + 15000  # irregular gaps of a few days
    return x[keep], y[keep].astype(float)[:, None]


def data_banana():
    # notebooks/2d_classification.py:18-21
    inputs = np.loadtxt(os.path.join(DATA, "banana_X_train"), delimiter=",")
    X = inputs[:, :1]
    R = inputs[:, 1:]
    Y = np.loadtxt(os.path.join(DATA, "banana_Y_train"))[:, None]
    return X, Y, R


def data_mcycle():
    # notebooks/heteroscedastic_noise.py:16-46: motorcycle data, standardised, fold 0 of the 10-fold split
    D = np.loadtxt(os.path.join(DATA, "mcycle.csv"), delimiter=",")
    X, Y = D[:, 1:2], D[:, 2:]
    Xall = (X - X.mean(0)) / X.std(0)                       # sklearn's StandardScaler: population standard deviation
    Yall = (Y - Y.mean(0)) / Y.std(0)
    cvind = np.loadtxt(os.path.join(REF, "kalmanjax", "experiments", "heteroscedastic", "cvind.csv")).astype(int)
    nt = np.floor(cvind.shape[0] / 10).astype(int)
    cvind = np.reshape(cvind[:10 * nt], (10, nt))
    test = cvind[0, :]
    train = np.setdiff1d(cvind, test)
    return Xall[train, :], Yall[train, :], Xall[test, :], Yall[test, :]


# ----------------------------------------------------------------------------- the replay
def replay(model, iters, keep_every=1, t_test=None, y_test=None, r_test=None):
    """What SDEGP.run() does (sde_gp.py:163-188) minus the gradient, recorded step by step."""
    out = {}
    for it in range(iters):
        params = [model.prior.hyp.copy() if not isinstance(model.prior.hyp, list) else list(model.prior.hyp),
                  model.likelihood.hyp.copy()]
        site_in = model.sites.site_params
        nlml, (fm, fc, sites_f) = model.kalman_filter(model.y, model.dt, params, True, None, site_in, model.r)
        model.sites.site_params = sites_f
        sites_s, sm, sc = model.rauch_tung_striebel_smoother(params, fm, fc, model.dt, True, False, model.y,
                                                            model.sites.site_params, model.r)
        model.sites.site_params = sites_s
        out["it%d_nlml" % it] = np.float64(nlml)
        out["it%d_filter_mean" % it] = np.asarray(fm)[::keep_every]
        out["it%d_filter_cov" % it] = np.asarray(fc)[::keep_every]
        out["it%d_site_mean_filter" % it] = np.asarray(sites_f[0])
        out["it%d_site_cov_filter" % it] = np.asarray(sites_f[1])
        out["it%d_smoothed_mean" % it] = np.asarray(sm)
        out["it%d_smoothed_cov" % it] = np.asarray(sc)
        out["it%d_site_mean" % it] = np.asarray(sites_s[0])
        out["it%d_site_cov" % it] = np.asarray(sites_s[1])
    if t_test is not None:
        pm, pv = model.predict(t=t_test, r=r_test)
        out["predict_mean"], out["predict_var"] = np.asarray(pm), np.asarray(pv)
        if y_test is not None:
            out["nlpd"] = np.float64(model.negative_log_predictive_density(t=t_test, y=y_test, r=r_test))
    out["keep_every"] = np.int64(keep_every)
    return out


def hyp_of(model):
    from utils import softplus, softplus_list
    ph = model.prior.hyp
    if isinstance(ph, list):
        theta = np.concatenate([np.atleast_1d(np.asarray(h, dtype=float)) for h in softplus_list(ph)])
    else:
        theta = np.atleast_1d(np.asarray(softplus(ph), dtype=float))
    lh = np.atleast_1d(np.asarray(softplus(model.likelihood.hyp), dtype=float)) if np.size(model.likelihood.hyp) else np.zeros(0)
    return theta, lh


sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_cases import CASES, KEEP_EVERY, build  # noqa: E402


class poisson_slr_patch(object):
    """Reference HEAD cannot run Poisson with the SLR family (SLEP/GHEP/UEP/GHKS/UKS/PL):
    `Poisson.conditional_moments` (likelihoods.py:632) returns a [1,1,Q] covariance, so `S` in
    `statistical_linear_regression_cubature` (likelihoods.py:229-231) becomes [1,1,1] and
    `np.diag(np.diag(om_sig_om))` (approximate_inference.py:199) raises "diag input must be 1d or 2d".
    For those cases only, this context manager swaps in the line the reference itself keeps commented
    out one line above (likelihoods.py:631: `return self.link_fn(f), self.link_fn(f)`), which is the
    evident scalar semantics.  Fixtures produced this way carry `reference_patched=1`."""
    def __init__(self, active):
        self.active = active

    def __enter__(self):
        if self.active:
            self._orig = likelihoods.Poisson.conditional_moments
            likelihoods.Poisson.conditional_moments = lambda self_, f, hyp=None: (self_.link_fn(f), self_.link_fn(f))
        return self

    def __exit__(self, *exc):
        if self.active:
            likelihoods.Poisson.conditional_moments = self._orig
        return False


def load_data(name):
    """-> t, y, r, t_test, y_test"""
    if name == "regression":
        x, y, xt, yt = data_regression()
        return x, y, None, xt, yt
    if name == "classification":
        x, y, xt, yt = data_classification()
        return x, y, None, xt, yt
    if name == "classification300":
        x, y, xt, yt = data_classification(300)
        return x, y, None, xt, yt
    if name == "coal200":
        x, y = data_coal_200()
        return x, y, None, np.linspace(x.min() - 2.0, x.max() + 2.0, 23), None
    if name == "coal333":
        x, y = data_coal_333()
        return x, y, None, x[::17] + 0.123, y[::17]
    if name == "aircraft160":
        x, y = data_aircraft()
        return x, y, None, x[5::20] + 0.5, y[5::20]
    if name == "mcycle":
        x, y, xt, yt = data_mcycle()
        return x, y, None, xt, yt
    if name == "banana400":
        X, Y, R = data_banana()
        return X, Y, R, None, None
    if name == "banana200":
        X, Y, R = data_banana()
        return X[:200], Y[:200], R[:200], None, None
    if name == "banana120":
        X, Y, R = data_banana()
        return X[:120], Y[:120], R[:120], None, None
    raise KeyError(name)


def make(name):
    data, prior_spec, lik_spec, (meth_name, meth_kw), iters, kat = CASES[name]
    t, y, r, t_test, y_test = load_data(data)
    prior = build(prior_spec, priors)
    lik = build(lik_spec, likelihoods)
    meth = getattr(approx_inf, meth_name)(**meth_kw)
    patched = isinstance(lik, likelihoods.Poisson) and isinstance(meth, approx_inf.StatisticallyLinearisedEP)
    model = SDEGP(prior=prior, likelihood=lik, t=t, y=y, r=r, approx_inf=meth)
    with poisson_slr_patch(patched):
        out = replay(model, iters, keep_every=KEEP_EVERY.get(data, 1), t_test=t_test, y_test=y_test)
    theta, lh = hyp_of(model)
    out.update(t=np.asarray(t, dtype=float), y=np.asarray(y, dtype=float), theta_prior=theta, theta_lik=lh,
               reference_patched=np.int64(patched))
    if r is not None:
        out["r"] = np.asarray(r, dtype=float)
        out["z"] = np.asarray(prior.z)
    if t_test is not None:
        out["t_test"] = np.asarray(t_test, dtype=float)
    if y_test is not None:
        out["y_test"] = np.asarray(y_test, dtype=float)
    if kat is not None:
        got = float(out["it0_nlml"])
        assert abs(got - kat) < 0.0051, (name, got, kat)
        out["kat_nlml_printed"] = np.float64(kat)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-38s nlml it0 = %.10f%s" % (name, float(out["it0_nlml"]),
                                         "" if kat is None else "   (notebook prints %.2f)" % kat))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        t0 = time.time()
        make(nm)
        print("    (%.1f s)" % (time.time() - t0))
