"""The plain-C restatement (oracle/seq_filter.c, the CPU baseline of bench.py) against the NumPy oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from golden_util import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libkjx_oracle.so")


def _lib():
    src = os.path.join(ROOT, "oracle", "seq_filter.c")
    if not os.path.exists(SO) or os.path.getmtime(src) > os.path.getmtime(SO):
        subprocess.check_call(["gcc", "-O3", "-shared", "-fPIC", "-o", SO, src, "-lm"])
    lib = C.CDLL(SO)
    lib.kjx_oracle_run.restype = C.c_double
    return lib


@pytest.mark.parametrize("kind,damping", [(2, 1.0), (1, 1.0), (2, 0.6)])
def test_c_oracle_matches_numpy_oracle(kind, damping):
    lib = _lib()
    rng = np.random.default_rng(3)
    N = 2000
    t = np.cumsum(rng.uniform(0.05, 0.15, N))
    y = np.cos(0.04 * t + 0.33 * np.pi) * np.sin(0.2 * t) + np.sqrt(0.15) * rng.standard_normal(N)
    prior = (oracle.priors.Matern32 if kind == 1 else oracle.priors.Matern52)(variance=1.0, lengthscale=5.0)
    model = oracle.SDEGP(prior, oracle.likelihoods.Gaussian(0.5), t, y, approx_inf=oracle.approx_inf.EP(power=0.5, damping=damping))
    d = prior.state_dim
    for it in range(2):
        sites = model.sites.site_params
        nlml, (fm, fP, s_f) = model.kalman_filter(model.y, model.dt, True, None, sites, model.r)
        s_new, pm, pv = model.rauch_tung_striebel_smoother(fm, fP, model.dt, True, False, model.y, s_f, model.r)
        out = dict(fm=np.zeros([N, d]), fP=np.zeros([N, d, d]), sm=np.zeros(N), sv=np.zeros(N), pm=np.zeros(N), pv=np.zeros(N))
        dp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        s_in = None if sites is None else (np.ascontiguousarray(sites[0].reshape(-1)), np.ascontiguousarray(sites[1].reshape(-1)))
        got = lib.kjx_oracle_run(kind, C.c_double(1.0), C.c_double(5.0), C.c_double(0.5), C.c_double(0.5), C.c_double(damping),
                                 C.c_long(N), dp(model.dt), dp(np.ascontiguousarray(model.y[:, 0])),
                                 dp(s_in[0]) if s_in else None, dp(s_in[1]) if s_in else None,
                                 dp(out["fm"]), dp(out["fP"]), dp(out["sm"]), dp(out["sv"]), dp(out["pm"]), dp(out["pv"]))
        assert abs(got - nlml) < 1e-10 * abs(nlml)
        assert rel_err(out["fm"], fm[:, :, 0]) < 1e-11 and rel_err(out["fP"], fP) < 1e-11
        assert rel_err(out["pm"], pm[:, 0, 0]) < 1e-11 and rel_err(out["pv"], pv[:, 0, 0]) < 1e-11
        assert rel_err(out["sm"], s_new[0][:, 0, 0]) < 1e-11 and rel_err(out["sv"], s_new[1][:, 0, 0]) < 1e-11
        model.sites.site_params = s_new
