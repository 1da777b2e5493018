"""CPU checks of the product's __host__ __device__ math (compiled for the host by
tests/host/kjx_hostemu.cu) against the NumPy oracle: the three-phase scan algebra and every
likelihood x inference-method site update.  No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from golden_util import load, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "host", "libkjx_hostemu.so")


def _lib():
    src = os.path.join(HERE, "host", "kjx_hostemu.cu")
    deps = [src, os.path.join(ROOT, "include", "kjx.h")] + [os.path.join(ROOT, "kalman_jax_b200", "csrc", f) for f in ("kjx_small_math.cuh", "kjx_sites.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "--shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-o", SO, src])
    return C.CDLL(SO)


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("two_filter", [0, 1])
@pytest.mark.parametrize("kind,L", [("Matern52", 7), ("Matern52", 64), ("Matern32", 16)])
def test_scan_emulation_matches_sequential_oracle(kind, L, two_filter):
    from kalman_jax_b200 import _lib as K
    lib = _lib()
    rng = np.random.default_rng(1)
    N = 3000
    t = np.cumsum(rng.uniform(0.05, 0.15, N))
    y = np.cos(0.04 * t + 0.33 * np.pi) * np.sin(0.2 * t) + np.sqrt(0.15) * rng.standard_normal(N)
    prior = getattr(oracle.priors, kind)(variance=1.0, lengthscale=5.0)
    model = oracle.SDEGP(prior, oracle.likelihoods.Gaussian(0.5), t, y, approx_inf=oracle.approx_inf.EP(power=0.5))
    mask = np.zeros([N, 1], dtype=bool)
    mask[rng.integers(0, N, 200)] = True
    site = (rng.standard_normal([N, 1, 1]), rng.uniform(0.2, 2.0, [N, 1, 1]))
    nlml, (fm, fP, _) = model.kalman_filter(model.y, model.dt, True, mask, site, model.r)
    _, sm, sP = model.rauch_tung_striebel_smoother(fm, fP, model.dt, True, True, None, None, model.r)
    d = prior.state_dim
    _, Pinf = prior.kernel_to_state_space()
    lam = (np.sqrt(5.0) if kind == "Matern52" else np.sqrt(3.0)) / 5.0
    o = [np.zeros(s) for s in ([N, d], [N, d, d], [N, d], [N, d, d])]
    m8 = np.ascontiguousarray(mask[:, 0].astype(np.uint8))
    rc = lib.emu_filter_smoother(K.BLOCK_MATERN52 if kind == "Matern52" else K.BLOCK_MATERN32, C.c_double(lam),
                                 _dp(np.ascontiguousarray(Pinf)), C.c_int64(N), L, _dp(model.dt),
                                 _dp(np.ascontiguousarray(site[0].reshape(-1))), _dp(np.ascontiguousarray(site[1].reshape(-1))),
                                 _dp(m8), *[_dp(a) for a in o], two_filter)
    assert rc == 0
    assert rel_err(o[0], fm[:, :, 0]) < 1e-11
    assert rel_err(o[1], fP) < 1e-11
    assert rel_err(o[2], sm[:, :, 0]) < 1e-11
    assert rel_err(o[3], sP) < 1e-11


LIKS = {
    "gauss": (lambda: oracle.likelihoods.Gaussian(0.37), 1, 0.37),
    "probit": (lambda: oracle.likelihoods.Probit(), 2, 0.0),
    "logit": (lambda: oracle.likelihoods.Logit(), 3, 0.0),
    "poisson": (lambda: oracle.likelihoods.Poisson(), 4, 0.0),
    "poisson_logistic": (lambda: oracle.likelihoods.Poisson("logistic"), 5, 0.0),
}
METHODS = {
    "ep1": (lambda: oracle.approx_inf.EP(power=1.0), 1),
    "pep_gh_damped": (lambda: oracle.approx_inf.EP(power=0.5, damping=0.6), 1),
    "pep_ut": (lambda: oracle.approx_inf.EP(power=0.9, intmethod="UT"), 1),
    "eep": (lambda: oracle.approx_inf.EEP(power=0.5, damping=0.7), 2),
    "eks": (lambda: oracle.approx_inf.EKS(), 2),
    "ghep": (lambda: oracle.approx_inf.GHEP(power=1), 3),
    "uks_damped": (lambda: oracle.approx_inf.UKS(damping=0.5), 3),
    "slep_ut3": (lambda: oracle.approx_inf.SLEP(power=0.3, intmethod="UT3"), 3),
    "vi": (lambda: oracle.approx_inf.VI(), 4),
    "vi_ut_damped": (lambda: oracle.approx_inf.VI(intmethod="UT", damping=0.5), 4),
}


@pytest.mark.parametrize("lik", sorted(LIKS))
@pytest.mark.parametrize("meth", sorted(METHODS))
@pytest.mark.parametrize("has_prev", [False, True, "negative_cavity"])
def test_site_update_math_matches_oracle(lik, meth, has_prev):
    from kalman_jax_b200 import _lib as K
    lib = _lib()
    likelihood, lik_code, lik_hyp = LIKS[lik][0](), LIKS[lik][1], LIKS[lik][2]
    rule, meth_code = METHODS[meth][0](), METHODS[meth][1]
    rng = np.random.default_rng(7)
    N = 400
    pm = rng.normal(0.0, 1.0, N)
    pv = rng.uniform(0.05, 1.5, N)
    if lik == "gauss":
        y = rng.normal(0, 1.5, N)
    elif lik in ("probit", "logit"):
        y = (rng.uniform(size=N) < 0.5).astype(float)
    else:
        y = rng.poisson(1.5, N).astype(float)
    # previous site variance well above the marginal variance keeps the cavity well conditioned;
    # "negative_cavity" mixes in marginals wider than the site (cavity variance < 0 -> reference NaNs)
    prev = (rng.normal(0, 1, N), rng.uniform(2.0, 6.0, N)) if has_prev else None
    if has_prev == "negative_cavity":
        prev = (prev[0], rng.uniform(0.2, 1.0, N))
    with np.errstate(all="ignore"):
        lml, sm, sv = rule.update(likelihood, y, pm, pv, prev)
    m = K.KjxModel()
    m.state_dim, m.func_dim, m.n_blocks = 2, 1, 1
    m.likelihood, m.method, m.lik_hyp, m.power, m.damping = lik_code, meth_code, lik_hyp, float(getattr(rule, "power", 1.0)), float(rule.damping)
    x, w = rule.cub
    m.n_cub = len(x)
    for i in range(len(x)):
        m.cub_x[i], m.cub_w[i] = x[i], w[i]
    out = [np.zeros(N) for _ in range(3)]
    rc = lib.emu_site_update(C.byref(m), C.c_int64(N), _dp(y), _dp(pm), _dp(pv),
                             _dp(prev[0]) if has_prev else None, _dp(prev[1]) if has_prev else None,
                             *[_dp(a) for a in out])
    assert rc == 0
    if has_prev == "negative_cavity":
        # Entries whose marginal is wider than the previous site have a negative cavity variance: the
        # reference's 1x1 Cholesky gives NaN there and the NaN pattern must match.  Entries with a
        # comfortably positive cavity must agree in value.  In between the cavity is near-singular and
        # the reference's own clamps (likelihoods.py:157-158,182) sit on rounding noise: not compared.
        prec = 1.0 / pv - float(getattr(rule, "power", 1.0)) / prev[1]      # cavity precision
        sel = (prec < -0.02 / pv) | (prec > 0.5 / pv)
    else:
        sel = np.ones(N, dtype=bool)
    for got, want in zip(out, (lml, sm, sv)):
        want = np.broadcast_to(np.asarray(want, dtype=float), (N,))
        assert np.array_equal(np.isnan(got[sel]), np.isnan(want[sel]))
        ok = ~np.isnan(want) & sel
        np.testing.assert_allclose(got[ok], want[ok], rtol=2e-9 if has_prev != "negative_cavity" else 1e-6, atol=1e-12)
