"""A small, representative slice of the GPU paths for compute-sanitizer (memcheck / racecheck / synccheck):
the fused d = 3 scan (run, posterior with mask), the linked single-shard iteration, the generic d = 59 path (fold, walks,
filter, gains, TMA-fed smoother), the one-chain first iteration of a Poisson model, the sampler and the batched kernels.

    compute-sanitizer --tool memcheck  python scripts/sanitize_subset.py
    compute-sanitizer --tool racecheck python scripts/sanitize_subset.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kalman_jax_b200 as kjx  # noqa: E402
from kalman_jax_b200.sharded import ShardedIteration, PeerExchange  # noqa: E402

rng = np.random.default_rng(0)
P, Lk, A = kjx.priors, kjx.likelihoods, kjx.approximate_inference

# fused scan, d = 3 and d = 2
N = 20_011
t = np.cumsum(rng.uniform(0.05, 0.15, N)); y = np.sin(0.2 * t) + 0.4 * rng.standard_normal(N)
for prior in (P.Matern52(1.0, 5.0), P.Matern32(1.0, 5.0)):
    m = kjx.SDEGP(prior, Lk.Gaussian(0.5), t, y, approx_inf=A.EP(power=0.5))
    for _ in range(2):
        m.run(compute_grad=False, store_posterior=True)
    m.predict(t=np.linspace(t[0] - 1, t[-1] + 1, 501))
# linked single-shard iteration (top scan with ticket, extra blocks)
dt = np.concatenate([[0.0], np.diff(t)])
it = ShardedIteration(P.Matern52(1.0, 5.0), Lk.Gaussian(0.5), A.EP(power=0.5), dt, y, 0.0, 0, 1, torch.device("cuda", 0),
                      site_mean=y.copy(), site_var=np.full(N, 0.5), peer=PeerExchange(1, 0, torch.device("cuda", 0)))
for _ in range(3):
    it.iteration()
# one-chain first iteration + later scan iterations of a non-Gaussian model
yp = rng.poisson(np.exp(0.5 * np.sin(0.2 * t[:3000]))).astype(float)
m = kjx.SDEGP(P.Matern52(1.0, 5.0), Lk.Poisson(), t[:3000], yp, approx_inf=A.EP(power=0.5))
for _ in range(2):
    m.run(compute_grad=False)
# generic path, d = 59 (shared-memory build) and a spatio-temporal d = 75 (global-scratch build)
Ng = 2000
tg = np.arange(Ng, dtype=float); yg = rng.poisson(0.05, Ng).astype(float)
prior = P.Sum([P.Matern52(variance=2., lengthscale=5.5e4),
               P.QuasiPeriodicMatern32(variance=1., lengthscale_periodic=2., period=365., lengthscale_matern=1.5e4),
               P.QuasiPeriodicMatern32(variance=1., lengthscale_periodic=2., period=7., lengthscale_matern=30 * 365.)])
m = kjx.SDEGP(prior, Lk.Poisson(), tg, yg, approx_inf=A.GHEP(power=1))
for _ in range(2):
    m.run(compute_grad=False, store_posterior=True)
m.predict(t=np.linspace(0, Ng, 50))
ts = np.sort(rng.uniform(0, 40, 300)); rs = rng.uniform(-3, 3, 300); ys = np.sign(rng.standard_normal(300))
m = kjx.SDEGP(P.SpatioTemporalMatern52(1.0, 1.0, 0.5, z=np.linspace(-3, 3, 25)), Lk.Probit(), ts, ys, r=rs, approx_inf=A.VI())
for _ in range(2):
    m.run(compute_grad=False)
# sampler (d > 32: several warps factor Q) and batched site update
m = kjx.SDEGP(prior, Lk.Gaussian(0.1), tg[:50], yg[:50])
m.prior_sample(64, seed=1)
A.EP(power=0.5).update(Lk.Poisson(), torch.as_tensor(yp, device="cuda"), torch.zeros(3000, device="cuda", dtype=torch.float64),
                       torch.ones(3000, device="cuda", dtype=torch.float64), None, None)
torch.cuda.synchronize()
print("sanitize subset done")
