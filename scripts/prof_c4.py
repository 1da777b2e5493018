"""Per-kernel device times of a C4-shaped iteration (Ns = 64 -> d = 192, global-scratch path); not a bench number."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kalman_jax_b200 as kjx
from torch.profiler import profile, ProfilerActivity
rng = np.random.default_rng(99)
N = int(os.environ.get("C4_N", 1024))
t = np.sort(rng.uniform(0, 400, N)); r = rng.uniform(-3, 3, N); y = (rng.uniform(size=N) < 0.5).astype(float)
model = kjx.SDEGP(kjx.priors.SpatioTemporalMatern52(1.0, 1.0, 0.3, z=np.linspace(-3, 3, 64)), kjx.likelihoods.Probit(), t, y, r=r,
                  approx_inf=kjx.approximate_inference.VI())
model.run(compute_grad=False); model.run(compute_grad=False); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        model.run(compute_grad=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=60))
