#!/bin/bash
# round-2 evidence: the default bench line, the ncu launch list of the same command (eager launches: kernels inside CUDA
# graphs are not listed), and one --set full capture of the two largest kernels for their DRAM traffic per launch
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none -k regex:'fused_(smoother|filter)_kernel' -s 6 -c 2 --csv --page raw --log-file gpurun_out/ncu_full_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --no-e2e --no-parity --no-extra > gpurun_out/ncu_full.log 2>&1
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import kalman_jax_b200 as kjx
rng = np.random.default_rng(0)
N = 2000
t = np.sort(rng.uniform(0, 200, N)); y = np.sin(0.3 * t) + 0.3 * rng.standard_normal(N)
m = kjx.SDEGP(kjx.priors.Matern52(1.0, 2.0), kjx.likelihoods.Gaussian(0.1), t, y, approx_inf=kjx.approximate_inference.EP(power=1.0))
m.run(compute_grad=False); m.run(compute_grad=False)
tt = np.linspace(0, 200, 500)
for batched in (True, False):
    m.posterior_sample(4, t=tt, seed=1, batched=batched); torch.cuda.synchronize()
    t0 = time.time(); s = m.posterior_sample(256, t=tt, seed=2, batched=batched); torch.cuda.synchronize()
    print("posterior_sample 256 draws, N_all=%d: batched=%s %.1f ms" % (N + 500, batched, 1e3 * (time.time() - t0)))
PY
