"""The NumPy oracle against fixtures recorded from the unmodified reference (CPU only)."""
import numpy as np
import pytest

import oracle
from golden_cases import CASES, KEEP_EVERY
from golden_util import load, build_model, rel_err

TOL = 1e-9


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference(name):
    g = load(name)
    model, iters = build_model(name, g, oracle.priors, oracle.likelihoods, oracle.approx_inf, oracle.SDEGP)
    ke = KEEP_EVERY.get(CASES[name][0], 1)
    for it in range(iters):
        nlml, (fm, fP, sites) = model.kalman_filter(model.y, model.dt, True, None, model.sites.site_params, model.r)
        model.sites.site_params = sites
        sites_s, sm, sc = model.rauch_tung_striebel_smoother(fm, fP, model.dt, True, False, model.y, sites, model.r)
        model.sites.site_params = sites_s
        pre = "it%d_" % it
        assert abs(nlml - g[pre + "nlml"]) <= TOL * abs(g[pre + "nlml"]), (it, nlml, g[pre + "nlml"])
        assert rel_err(fm[::ke], g[pre + "filter_mean"]) < TOL
        assert rel_err(fP[::ke], g[pre + "filter_cov"]) < TOL
        assert rel_err(sites[0], g[pre + "site_mean_filter"]) < TOL
        assert rel_err(sites[1], g[pre + "site_cov_filter"]) < TOL
        assert rel_err(sm, g[pre + "smoothed_mean"]) < TOL
        assert rel_err(sc, g[pre + "smoothed_cov"]) < TOL
        assert rel_err(sites_s[0], g[pre + "site_mean"]) < TOL
        assert rel_err(sites_s[1], g[pre + "site_cov"]) < TOL
    if "predict_mean" in g:
        pm, pv = model.predict(t=g["t_test"])
        assert rel_err(pm, g["predict_mean"]) < TOL
        assert rel_err(pv, g["predict_var"]) < TOL
    if "nlpd" in g:
        nlpd = model.negative_log_predictive_density(t=g["t_test"], y=g["y_test"])
        assert abs(nlpd - g["nlpd"]) <= TOL * abs(g["nlpd"])
    if "kat_nlml_printed" in g:
        assert abs(float(g["it0_nlml"]) - float(g["kat_nlml_printed"])) < 0.0051


def test_known_answers_from_reference_notebooks():
    # K1 regression.ipynb:176, K2 log_gaussian_cox_process.ipynb:173,
    # K3 classification.ipynb:176, K5 2d_classification.ipynb:166
    kats = {n: c[5] for n, c in CASES.items() if c[5] is not None}
    assert sorted(kats.values()) == [178.81, 295.62, 469.46, 797.76]
    for n, v in kats.items():
        assert round(float(load(n)["it0_nlml"]), 2) == v
