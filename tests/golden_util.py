import os

import numpy as np

from golden_cases import CASES, build

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def build_model(name, g, priors_mod, lik_mod, inf_mod, sdegp_cls):
    """Instantiate the golden case `name` from any module family with the reference's class names."""
    _, prior_spec, lik_spec, (meth_name, meth_kw), iters, _ = CASES[name]
    prior = build(prior_spec, priors_mod)
    lik = build(lik_spec, lik_mod)
    meth = getattr(inf_mod, meth_name)(**meth_kw)
    model = sdegp_cls(prior=prior, likelihood=lik, t=g["t"], y=g["y"], r=g.get("r"), approx_inf=meth)
    return model, iters


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)
