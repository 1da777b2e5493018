#!/usr/bin/env python
"""bench_extra.py — secondary measurements of the same hot path (not the driver's contract; see bench.py for that).

One JSON object per line:
  * site_update: kjx_site_update throughput (steps/s, GB/s of its 6 arrays) for Poisson/GH20 EP, Probit VI, Gaussian EP
  * fp32:        the fused iteration in float (C5 recipe), steps/s and fraction of the 144 B/step fp32 roofline
  * c1:          BASELINE config 1 (N = 10k) iteration latency
  * c3:          aircraft-sized series (N = 36 020 daily bins, Sum[Matern52, QPM32, QPM32], d = 59, Poisson, GHEP)
  * c4:          spatio-temporal Probit, Ns = 64 -> d = 192, N = 4096 (only when asked for: `bench_extra.py c4`)
  * first_iter:  first iteration of a non-Gaussian model (one-chain filter that initialises the sites), N = 2^17
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import kalman_jax_b200 as kjx  # noqa: E402
from bench import synth, peaks  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def site_update():
    N = 1 << 24
    rng = np.random.default_rng(0)
    pm = torch.as_tensor(rng.normal(0, 1, N), device="cuda")
    pv = torch.as_tensor(rng.uniform(0.05, 1.5, N), device="cuda")
    sm, sv = torch.as_tensor(rng.normal(0, 1, N), device="cuda"), torch.as_tensor(rng.uniform(2, 6, N), device="cuda")
    hbm, _ = peaks()
    for name, lik, rule, y in [
        ("poisson_ep05_gh20", kjx.likelihoods.Poisson(), kjx.approximate_inference.EP(power=0.5), rng.poisson(1.5, N).astype(float)),
        ("probit_vi_gh20", kjx.likelihoods.Probit(), kjx.approximate_inference.VI(), (rng.uniform(size=N) < 0.5).astype(float)),
        ("poisson_eks", kjx.likelihoods.Poisson(), kjx.approximate_inference.EKS(), rng.poisson(1.5, N).astype(float)),
        ("gauss_ep", kjx.likelihoods.Gaussian(0.5), kjx.approximate_inference.EP(power=0.5), rng.normal(0, 1, N)),
    ]:
        y_d = torch.as_tensor(y, device="cuda")
        hyp = np.array(0.5) if name == "gauss_ep" else None
        t = timed(lambda: rule.update(lik, y_d, pm, pv, hyp, (sm, sv)))
        print(json.dumps({"bench": "site_update", "case": name, "N": N, "ms": t * 1e3, "steps_per_s": N / t,
                          "GBps": 8 * 8 * N / t / 1e9, "frac_hbm": 8 * 8 * N / t / 1e9 / hbm}))


def fused(dtype, log2n, name):
    N = 1 << log2n
    dt, y = synth(N)
    t = np.cumsum(dt)
    model = kjx.SDEGP(kjx.priors.Matern52(1.0, 5.0), kjx.likelihoods.Gaussian(0.5), t, y,
                      approx_inf=kjx.approximate_inference.EP(power=0.5), dtype=dtype)
    model.run(compute_grad=False)
    sec = timed(lambda: model.run(compute_grad=False), reps=10)
    hbm, _ = peaks()
    b = 288 if dtype == torch.float64 else 144
    print(json.dumps({"bench": name, "dtype": str(dtype), "N": N, "ms_per_iteration": sec * 1e3, "steps_per_s": N / sec,
                      "frac_hbm": b * N / sec / 1e9 / hbm, "note": "SDEGP.run(compute_grad=False): includes the host-side model setup and "
                      "one D2H read of nlml per iteration"}))


def c3():
    rng = np.random.default_rng(1)
    N = 36020
    t = np.arange(N, dtype=float)
    y = rng.poisson(0.035, N).astype(float)
    P, O = kjx.priors, __import__("oracle")
    mk = lambda pr: pr.Sum([pr.Matern52(variance=2., lengthscale=5.5e4),  # noqa: E731
                            pr.QuasiPeriodicMatern32(variance=1., lengthscale_periodic=2., period=365., lengthscale_matern=1.5e4),
                            pr.QuasiPeriodicMatern32(variance=1., lengthscale_periodic=2., period=7., lengthscale_matern=30 * 365.)])
    model = kjx.SDEGP(mk(P), kjx.likelihoods.Poisson(), t, y, approx_inf=kjx.approximate_inference.GHEP(power=1))
    t0 = time.perf_counter(); model.run(compute_grad=False); torch.cuda.synchronize(); first = time.perf_counter() - t0
    sec = timed(lambda: model.run(compute_grad=False), reps=2, warm=0)
    n_cpu = 300
    ref = O.SDEGP(mk(O.priors), O.likelihoods.Poisson(), t[:n_cpu], y[:n_cpu], approx_inf=O.approx_inf.GHEP(power=1))
    ref.run()
    t0 = time.perf_counter(); ref.run(); cpu = (time.perf_counter() - t0) / n_cpu
    print(json.dumps({"bench": "c3_aircraft_d59", "N": N, "d": 59, "first_iteration_s": first, "s_per_iteration": sec,
                      "steps_per_s": N / sec, "numpy_oracle_steps_per_s": 1 / cpu,
                      "note": "generic d <= 64 path: CTA-level chunked scan (fold -> boundary -> filter / gains / smoother per chunk); the first iteration initialises the Poisson sites with the one-chain filter"}))


def c4():
    """BASELINE config 4 shape: Ns = 64 inducing points -> d = 192 (per-step dense H), N = 4096 scattered points,
    Probit, VI.  The generic kernels run from per-CTA global scratch here (untuned: DESIGN.md)."""
    rng = np.random.default_rng(99)
    N = int(os.environ.get("C4_N", 4096))
    t = np.sort(rng.uniform(0, 400, N))
    r = rng.uniform(-3, 3, N)
    y = (rng.uniform(size=N) < 0.5).astype(float)
    z = np.linspace(-3, 3, 64)
    model = kjx.SDEGP(kjx.priors.SpatioTemporalMatern52(1.0, 1.0, 0.3, z=z), kjx.likelihoods.Probit(), t, y, r=r,
                      approx_inf=kjx.approximate_inference.VI())
    t0 = time.perf_counter(); model.run(compute_grad=False); torch.cuda.synchronize(); first = time.perf_counter() - t0
    sec = timed(lambda: model.run(compute_grad=False), reps=2, warm=0)
    d = 192
    print(json.dumps({"bench": "c4_st52_ns64_d192", "N": N, "d": d, "first_iteration_s": first, "s_per_iteration": sec,
                      "steps_per_s": N / sec, "algorithmic_tflops": 12.0 * d ** 3 * N / sec / 1e12,
                      "note": "generic path from per-CTA global scratch (d > 64), FP64 tensor-pipe products; untuned"}))


def first_iter():
    """First iteration of a non-Gaussian model (sites initialised from the predictive, sde_gp.py:287): the fixed-point
    sweeps against the one-chain kernel (KJX_FLAG_SEQUENTIAL), after a warm-up model has loaded every kernel."""
    import ctypes as C
    from kalman_jax_b200 import _lib
    N = 1 << 17
    dt, y = synth(N)
    t = np.cumsum(dt)
    yy = np.random.default_rng(0).poisson(np.exp(0.5 * y)).astype(float)
    mk = lambda n: kjx.SDEGP(kjx.priors.Matern52(1.0, 5.0), kjx.likelihoods.Poisson(), t[:n], yy[:n],  # noqa: E731
                             approx_inf=kjx.approximate_inference.EP(power=0.5))
    warm = mk(4096)
    warm.run(compute_grad=False); warm.run(compute_grad=False)
    params = warm._params(None)
    warm.kalman_filter(warm.y, warm.dt, params, True, None, None, warm.r, sequential=True)
    torch.cuda.synchronize()
    model = mk(N)
    t0 = time.perf_counter(); model.run(compute_grad=False); torch.cuda.synchronize(); first = time.perf_counter() - t0
    sec = timed(lambda: model.run(compute_grad=False), reps=5)
    model2 = mk(N)
    t0 = time.perf_counter()
    model2.kalman_filter(model2.y, model2.dt, params, True, None, None, model2.r, sequential=True)
    torch.cuda.synchronize(); chain = time.perf_counter() - t0
    print(json.dumps({"bench": "first_iter_poisson_pep_gh20", "N": N, "first_iteration_s": first, "first_iter_steps_per_s": N / first,
                      "one_chain_filter_s": chain, "one_chain_steps_per_s": N / chain,
                      "later_iteration_s": sec, "later_steps_per_s": N / sec,
                      "note": "first iteration = fixed-point sweeps (parallel in time) + one fused iteration; one_chain = the "
                              "single-warp sequential filter alone (KJX_FLAG_SEQUENTIAL)"}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["site_update", "fp32", "c1", "c3", "first_iter"]
    if "c4" in which:
        c4()
    if "site_update" in which:
        site_update()
    if "fp32" in which:
        fused(torch.float32, 24, "fused_fp32")
        fused(torch.float64, 24, "fused_fp64_via_SDEGP_run")
    if "c1" in which:
        fused(torch.float64, 0, "unused") if False else None
        N = 10000
        dt, y = synth(N)
        model = kjx.SDEGP(kjx.priors.Matern52(1.0, 5.0), kjx.likelihoods.Gaussian(0.5), np.cumsum(dt), y,
                          approx_inf=kjx.approximate_inference.EP(power=0.5))
        model.run(compute_grad=False)
        sec = timed(lambda: model.run(compute_grad=False), reps=50)
        print(json.dumps({"bench": "c1_n10k", "N": N, "us_per_iteration": sec * 1e6, "steps_per_s": N / sec}))
    if "c3" in which:
        c3()
    if "first_iter" in which:
        first_iter()
