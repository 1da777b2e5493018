"""Per-kernel device times of the generic-path C3 iteration (CUPTI via torch.profiler; not a bench number)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kalman_jax_b200 as kjx
from torch.profiler import profile, ProfilerActivity
rng = np.random.default_rng(1)
N = int(os.environ.get("C3_N", 36020))
t = np.arange(N, dtype=float); y = rng.poisson(0.035, N).astype(float)
P = kjx.priors
prior = P.Sum([P.Matern52(variance=2., lengthscale=5.5e4),
               P.QuasiPeriodicMatern32(variance=1., lengthscale_periodic=2., period=365., lengthscale_matern=1.5e4),
               P.QuasiPeriodicMatern32(variance=1., lengthscale_periodic=2., period=7., lengthscale_matern=30 * 365.)])
if os.environ.get("C3_LIK") == "gauss":      # cheap per-step scalar work: isolates the cost of the cubature sections
    model = kjx.SDEGP(prior, kjx.likelihoods.Gaussian(0.5), t, y, approx_inf=kjx.approximate_inference.EP(power=0.5))
else:
    model = kjx.SDEGP(prior, kjx.likelihoods.Poisson(), t, y, approx_inf=kjx.approximate_inference.GHEP(power=1))
model.run(compute_grad=False); model.run(compute_grad=False); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        model.run(compute_grad=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
