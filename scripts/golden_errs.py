"""Print the relative error of every recorded quantity of one golden case on the GPU path (debugging aid)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kalman_jax_b200 as kjx
from golden_util import load, build_model, rel_err
from golden_cases import CASES, KEEP_EVERY
for name in sys.argv[1:]:
    g = load(name)
    model, iters = build_model(name, g, kjx.priors, kjx.likelihoods, kjx.approximate_inference, kjx.SDEGP)
    ke = KEEP_EVERY.get(CASES[name][0], 1)
    params = model._params(None)
    npf = lambda t: t.detach().cpu().numpy()
    for it in range(iters):
        nlml, (fm, fP, sites) = model.kalman_filter(model.y, model.dt, params, True, None, model.sites.site_params, model.r)
        model.sites.site_params = sites
        sites_s, sm, sc = model.rauch_tung_striebel_smoother(params, fm, fP, model.dt, True, False, model.y, sites, model.r)
        model.sites.site_params = sites_s
        pre = "it%d_" % it
        print(name, it, "nlml %.2e" % (abs(float(nlml) - g[pre + "nlml"]) / abs(g[pre + "nlml"])),
              "fm %.2e" % rel_err(npf(fm)[::ke], g[pre + "filter_mean"]), "fP %.2e" % rel_err(npf(fP)[::ke], g[pre + "filter_cov"]),
              "sm %.2e" % rel_err(npf(sm), g[pre + "smoothed_mean"]), "sc %.2e" % rel_err(npf(sc), g[pre + "smoothed_cov"]),
              "site_m %.2e" % rel_err(npf(sites_s[0]), g[pre + "site_mean"]), "site_v %.2e" % rel_err(npf(sites_s[1]), g[pre + "site_cov"]))
