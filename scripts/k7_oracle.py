"""K7: experiments/coal/coal.py, method 0 (EEP, power 1), fold 0, 250 Adam steps -> NLPD
0.8147043424175697 (experiments/coal/output/0_0_nlpd.txt), here with the NumPy oracle and central-difference gradients
of its filter energy.  CPU only; about a minute."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from kalman_jax_b200.optimizers import adam
from kalman_jax_b200.utils import softplus, softplus_inv

D = np.loadtxt(os.path.join(ROOT, "tests", "data", "binned.csv")); cvind = np.loadtxt(os.path.join(ROOT, "tests", "data", "cvind.csv")).astype(int)
nt = cvind.shape[0] // 10; cvind = cvind[:10 * nt].reshape(10, nt)
x, y = D[:, 0:1], D[:, 1:]
ind_test = cvind[0]; ind_train = np.setdiff1d(cvind, ind_test)
xt, yt, xs, ys = x[ind_train], y[ind_train], x[ind_test], y[ind_test]
raw = np.array([softplus_inv(1.0), softplus_inv(1.0)], dtype=float)

def model_for(rawp, sites):
    v, l = softplus(rawp)
    m = oracle.SDEGP(oracle.priors.Matern52(float(v), float(l)), oracle.likelihoods.Poisson(), xt, yt, approx_inf=oracle.approx_inf.EEP(power=1))
    m.sites.site_params = sites
    return m

def energy(rawp, sites):
    m = model_for(rawp, sites)
    return float(m.kalman_filter(m.y, m.dt, False, None, sites, m.r))

init, update, get = adam(step_size=2.5e-1)
state = init([raw, np.array([])])
sites = None
t0 = time.time()
for j in range(250):
    p = get(state)[0]
    eps = 1e-5
    g = np.array([(energy(p + eps * np.eye(2)[k], sites) - energy(p - eps * np.eye(2)[k], sites)) / (2 * eps) for k in range(2)])
    m = model_for(p, sites)
    nl = m.run(); sites = m.sites.site_params
    state = update(j, [g, np.array([])], state)
    if j % 50 == 0: print(j, softplus(p), nl, flush=True)
# coal.py:102-104, 128: the model keeps the hyperparameters of the LAST run() (the 250th update is never applied to it)
m = model_for(p, sites)
print("final hyp", softplus(p), "NLPD", repr(m.negative_log_predictive_density(t=xs, y=ys)), "want 0.8147043424175697", time.time() - t0, "s")
