#!/usr/bin/env python
"""bench.py — filter + smoother + site-update iteration throughput (time-steps/s) and % of HBM roofline.

Workload (BASELINE.json configs[4] == SURVEY.md C5): synthetic Matern-5/2 GP regression, Gaussian
likelihood, EP(power 0.5) sites, N = 2^24 time steps, state dimension d = 3, fp64; one "step" = one
steady-state SDEGP.run() iteration without the gradient: filter with the stored sites -> RTS smoother ->
per-step site update (sde_gp.py:179-187).  At --gpus G > 1 the SAME series is split into G contiguous
time shards (strong scaling), one process per GPU, two 28/19-double all-gathers per iteration.

  python bench.py [--gpus G] [--steps K] [--warmup W] [--n LOG2N] [--impl reference]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "filter+smoother time-steps/s (N=2^24, d=3/30) & % HBM roofline at 1/2/4/8 GPU"
BYTES_PER_STEP = 288          # 2(1+2f+f^2) + 2(d+d^2) + 2(f+f^2) scalars at d=3, f=1, fp64
BYTES_FILTER, BYTES_SMOOTHER = 128, 160


def synth(N, seed=12345):
    """SURVEY.md 8(d) synthetic series: spacing U(0.05, 0.15), the regression-notebook signal."""
    rng = np.random.default_rng(seed)
    dt = rng.uniform(0.05, 0.15, N)
    dt[0] = 0.0
    t = np.cumsum(dt)
    y = np.cos(0.04 * t + 0.33 * np.pi) * np.sin(0.2 * t) + np.sqrt(0.15) * rng.standard_normal(N)
    return dt, y


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def cpu_oracle_lib():
    import __graft_entry__ as g
    so = os.path.join(ROOT, "oracle", "libkjx_oracle.so")
    if not os.path.exists(so):
        g.build()
    lib = C.CDLL(so)
    lib.kjx_oracle_run.restype = C.c_double
    return lib


def cpu_chain(lib, dt, y, reps=1):
    """The C restatement of the reference recursion (oracle/seq_filter.c): one sequential chain."""
    n = len(dt)
    fm, fP = np.empty([n, 3]), np.empty([n, 3, 3])
    sm, sv, nsm, nsv = y.copy(), np.full(n, 0.5), np.empty(n), np.empty(n)
    dp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        lib.kjx_oracle_run(2, C.c_double(1.0), C.c_double(5.0), C.c_double(0.5), C.c_double(0.5), C.c_double(1.0), C.c_long(n),
                           dp(dt), dp(y), dp(sm), dp(sv), dp(fm), dp(fP), dp(nsm), dp(nsv), None, None)
        best = min(best, time.perf_counter() - t0)
    return n / best


def cpu_oracle_full(lib, dt, y, site_mean, site_var, power=0.5, damping=1.0, lik_var=0.5):
    """One whole run() iteration of the C oracle over the full series; returns nlml and every per-step output."""
    n = len(dt)
    out = dict(fm=np.empty([n, 3]), fP=np.empty([n, 3, 3]), nsm=np.empty(n), nsv=np.empty(n), pm=np.empty(n), pv=np.empty(n))
    dp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    out["nlml"] = lib.kjx_oracle_run(2, C.c_double(1.0), C.c_double(5.0), C.c_double(lik_var), C.c_double(power), C.c_double(damping),
                                     C.c_long(n), dp(dt), dp(y), dp(site_mean), dp(site_var), dp(out["fm"]), dp(out["fP"]),
                                     dp(out["nsm"]), dp(out["nsv"]), dp(out["pm"]), dp(out["pv"]))
    return out


def rel_err(a, b):
    """max |a - b| / max |b| — the definition the parity tests use (tests/golden_util.py)."""
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def fp64_peak():
    """Measured FP64 peak of this GPU (TFLOP/s): scripts/micro/dmma_rate (mma.sync.m8n8k4.f64 and plain DFMA streams,
    built in-tree by __graft_entry__.build()), run live; falls back to the committed measurement, then to 37.0."""
    exe = os.path.join(ROOT, "scripts", "micro", "dmma_rate")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        best = {"dmma": 0.0, "dfma": 0.0}
        for ln in out.splitlines():
            p = ln.split()
            if len(p) > 2 and p[0] in best and p[-1] == "TFLOP/s":
                best[p[0]] = max(best[p[0]], float(p[-2]))
        if best["dmma"] > 0:
            return {"tflops": max(best.values()), "dmma_tflops": best["dmma"], "dfma_tflops": best["dfma"],
                    "source": "measured live: scripts/micro/dmma_rate.cu (best over 4..32 warps per CTA, 2 CTAs per SM)"}
    except Exception:
        pass
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r02.json")))
        p["source"] = "profiles/fp64_peak_r02.json (committed measurement of the same microbenchmark)"
        return p
    except Exception:
        return {"tflops": 37.0, "source": "fallback: round-1 measurement of scripts/micro/dmma_rate.cu"}


def extra_workloads(peak):
    """The d ~ 30 / spatio-temporal legs of BASELINE's metric (C3 and C4), FP64-compute-bound: steady-state
    run() iterations per second on synthetic data of the named shapes, with the algorithmic 12 d^3 flops per step
    against the measured FP64 peak.  One GPU; a few iterations each (seconds)."""
    import torch
    import kalman_jax_b200 as kjx
    P = kjx.priors
    out = {}

    def timed(model, reps):
        t0 = time.perf_counter(); model.run(compute_grad=False); torch.cuda.synchronize(); first = time.perf_counter() - t0
        model.run(compute_grad=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            model.run(compute_grad=False)
        b.record()
        torch.cuda.synchronize()
        return first, a.elapsed_time(b) / reps * 1e-3

    def entry(name, workload, N, d, first, sec):
        tf = 12.0 * d ** 3 * N / sec / 1e12
        out[name] = {"workload": workload, "N": N, "state_dim": d, "s_per_iteration": sec, "first_iteration_s": first,
                     "value": N / sec, "unit": "time-steps/s",
                     "roofline": {"bound": "fp64", "achieved": tf, "peak": peak["tflops"], "unit": "TFLOP/s",
                                  "frac": tf / peak["tflops"], "algorithmic_flops_per_step": 12.0 * d ** 3}}

    rng = np.random.default_rng(1)
    N = 36020                                                   # aircraft.py:27-35: daily bins incl. padding
    t, y = np.arange(N, dtype=float), rng.poisson(0.035, N).astype(float)
    m52 = dict(variance=2., lengthscale=5.5e4)
    qp_year = dict(variance=1., lengthscale_periodic=2., period=365., lengthscale_matern=1.5e4)
    qp_week = dict(variance=1., lengthscale_periodic=2., period=7., lengthscale_matern=30 * 365.)
    for name, parts, rule, d in [
            ("c3_d59", [P.Matern52(**m52), P.QuasiPeriodicMatern32(**qp_year), P.QuasiPeriodicMatern32(**qp_week)],
             kjx.approximate_inference.GHEP(power=1), 59),
            ("c3_d31", [P.Matern52(**m52), P.QuasiPeriodicMatern32(**qp_year)], kjx.approximate_inference.VI(), 31)]:
        model = kjx.SDEGP(P.Sum(parts), kjx.likelihoods.Poisson(), t, y, approx_inf=rule)
        first, sec = timed(model, 5)
        entry(name, "C3: aircraft-shaped daily counts, Poisson, Sum[Matern52, QuasiPeriodicMatern32 x%d], %s; steady-state "
              "run() iteration" % (len(parts) - 1, rule.name), N, d, first, sec)
        del model
    rng = np.random.default_rng(99)
    N = 4096
    t, r = np.sort(rng.uniform(0, 400, N)), rng.uniform(-3, 3, N)
    y = (rng.uniform(size=N) < 0.5).astype(float)
    model = kjx.SDEGP(P.SpatioTemporalMatern52(1.0, 1.0, 0.3, z=np.linspace(-3, 3, 64)), kjx.likelihoods.Probit(), t, y, r=r,
                      approx_inf=kjx.approximate_inference.VI())
    first, sec = timed(model, 2)
    entry("c4_d192", "C4: spatio-temporal Matern-5/2, Ns=64 inducing points (d=192), Probit, VI, N=4096 scattered points; "
          "steady-state run() iteration (includes the host-side per-step measurement rows)", N, 192, first, sec)
    return out


def cpu_replicas(lib, dt, y, cores, reps=1):
    """`cores` independent copies of the chain, one per host thread (the reference algorithm cannot use more
    than one core on one series: step n needs step n-1, sde_gp.py:271)."""
    res = [0.0] * cores

    def work(i):
        res[i] = cpu_chain(lib, dt, y, reps)   # ctypes releases the GIL during the call
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    return cores * len(dt) * reps / wall


def run_reference(args):
    """--impl reference: the reference's CPU path, restated in C (JAX cannot be installed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib = cpu_oracle_lib()
    cores = os.cpu_count() or 1
    n = 1 << 21
    dt, y = synth(n)
    for _ in range(args.warmup):
        cpu_chain(lib, dt[: 1 << 18], y[: 1 << 18])
    t0 = time.perf_counter()
    vals = [cpu_replicas(lib, dt, y, cores) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    single = cpu_chain(lib, dt, y)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "time-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Matern-5/2 + Gaussian + EP(0.5), filter->smoother->site update, d=3 fp64 (C5 recipe)",
                   "N_sample": n, "note": "one sequential chain per series; value aggregates independent replicas"},
        "cpu_baseline": {"value": v, "unit": "time-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d independent replicas of a 2^21-step slice of the C5 series, one chain per host thread, "
                                   "C restatement oracle/seq_filter.c (single chain: %.3g steps/s)" % (cores, single),
                         "single_chain_value": single},
        "e2e": {"value": v, "unit": "time-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def make_collectives(world, rank, dev, exchange="peer"):
    """(gather, reduce_nlml, peer) for a world of `world` ranks: the NCCL all-gather of the shard summaries, the scalar
    all-reduce of the partial nlml values, and the symmetric-memory peer exchange (None when unavailable or not asked
    for; every rank agrees on the path)."""
    import torch
    import torch.distributed as dist
    gbuf = {}

    def gather(t):
        out = gbuf.get(t.numel())
        if out is None:
            out = gbuf[t.numel()] = torch.empty(world, t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out

    nl_buf = torch.zeros(1, dtype=torch.float64, device=dev)

    def reduce_nlml(t):
        nl_buf.copy_(t)
        dist.all_reduce(nl_buf)
        return nl_buf

    peer = None
    if exchange != "nccl":
        try:
            from kalman_jax_b200.sharded import PeerExchange
            peer = PeerExchange(world, rank, dev)
        except Exception as e:
            print("bench: symmetric memory unavailable (%s); using NCCL" % str(e).splitlines()[0], file=sys.stderr)
    ok = torch.tensor([1 if (peer is not None or exchange == "nccl") else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks must agree on the path
    if int(ok[0]) == 0:
        peer = None
    return gather, reduce_nlml, peer


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--log2n", type=int, default=24, help="log2 of the number of time steps")
    ap.add_argument("--impl", default="kjx")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size comparison with the C oracle")
    ap.add_argument("--no-extra", action="store_true", help="skip the d~30 / spatio-temporal legs (extra_workloads)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer-separate", "nccl"],
                    help="summary exchange: inside the kernels over peer memory (default), the separate exchange+fold "
                         "kernel over peer memory, or an NCCL all-gather")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import kalman_jax_b200 as kjx
    from kalman_jax_b200 import _lib
    from kalman_jax_b200.sharded import ShardedIteration, split_series

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gather = reduce_nlml = peer = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        gather, reduce_nlml, peer = make_collectives(world, rank, dev, args.exchange)
    elif args.exchange != "nccl":
        from kalman_jax_b200.sharded import PeerExchange
        peer = PeerExchange(1, 0, dev)                       # one shard: an ordinary device buffer, same kernels
    linked = peer is not None and args.exchange == "peer"
    if linked:
        exchange = ("inside the scan/filter/smoother kernels over symmetric memory (NVLink peer stores + epoch flags), "
                    "nlml partials included; no collective call" if world > 1 else "none (1 shard; same kernels, local buffer)")
        reduce_nlml = None
    elif peer is not None:
        exchange = "separate fused exchange+fold kernel over symmetric memory; nlml by NCCL all-reduce"
    else:
        exchange = "NCCL all_gather_into_tensor + fold kernel; nlml by NCCL all-reduce"

    N = 1 << args.log2n
    dt, y = synth(N)
    sh_dt, sh_y, dt_next, lo, hi = split_series(dt, y, world)[rank]
    n_local = hi - lo
    prior, lik, rule = kjx.priors.Matern52(1.0, 5.0), kjx.likelihoods.Gaussian(0.5), kjx.approximate_inference.EP(power=0.5)
    it = ShardedIteration(prior, lik, rule, sh_dt, sh_y, dt_next, rank, world, dev, gather,
                          site_mean=sh_y.copy(), site_var=np.full(n_local, 0.5), peer=peer, linked=linked)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_w = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        it.iteration(reduce_nlml=reduce_nlml)
    barrier()
    # Keep every rank under the same load for ~0.8 s more so nvidia-smi (200 ms period) delivers samples taken under
    # load before the timed region starts (not timed).  The count is agreed across ranks: the loop contains collectives.
    est = torch.tensor([(time.perf_counter() - t_w) / max(args.warmup, 3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    n_extra = int(min(2000, max(1, 0.8 / max(float(est[0]), 1e-5))))
    for _ in range(n_extra):
        it.iteration(update_sites=False, reduce_nlml=reduce_nlml)
    barrier()
    sampler.rows = sampler.rows[-2:]
    # per-phase CUDA-event times (untimed for `value`): eager launches with events between the SEPARATE library calls
    # (fold + scans | exchange + fold of the shard elements | filter | smoother) — the same fold / filter / smoother
    # kernels the linked iteration launches, so the split explains the timed number
    n_ev = max(3, min(args.steps, 20))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_ev)]
    if linked:
        gather_ph, reduce_ph, _ = make_collectives(world, rank, dev, "nccl") if world > 1 else (None, None, None)
        it_ph = ShardedIteration(prior, lik, rule, sh_dt, sh_y, dt_next, rank, world, dev, gather_ph,
                                 site_mean=sh_y.copy(), site_var=np.full(n_local, 0.5), peer=None, linked=False)
    else:
        it_ph, reduce_ph = it, reduce_nlml
    for k in range(3 + n_ev):
        it_ph.iteration(events=ev[max(k - 3, 0)], reduce_nlml=reduce_ph)
    barrier()
    phases = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in ev]).mean(axis=0)   # ms
    if linked:
        del it_ph
        torch.cuda.empty_cache()
    graphed = None
    # iterations per graph launch: the sites ping-pong between two buffers, so iterations are captured in pairs; the
    # largest of 5, 2, 1 pairs whose length divides --steps (an odd --steps falls back to one iteration per launch)
    pairs = next((p for p in (5, 2, 1) if args.steps % (2 * p) == 0), 0)
    if not args.no_graph:
        from kalman_jax_b200.sharded import GraphedIteration
        try:
            graphed = GraphedIteration(it, reduce_nlml, pairs=pairs)
        except Exception as e:      # keep going eagerly, but say so
            print("bench: CUDA-graph capture failed (%s); launching eagerly" % str(e).splitlines()[0], file=sys.stderr)
            graphed = None
    step_fn = graphed.replay if graphed else (lambda: it.iteration(reduce_nlml=reduce_nlml))
    for _ in range(4):
        step_fn()
    if graphed and pairs:
        graphed.replay_many()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    if graphed and pairs:                                   # K steps = K / (2 pairs) launches of the multi-iteration graph
        for k in range(args.steps // (2 * pairs)):
            nl_t = graphed.replay_many()
    else:
        for k in range(args.steps):
            nl_t = step_fn()
    stop.record()
    barrier()
    for _ in range(n_extra // 2):       # a few more samples under the same load right after the timed region
        it.iteration(update_sites=False, reduce_nlml=reduce_nlml)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = start.elapsed_time(stop)
    nlml = float(nl_t)

    # ---- parity at the benchmark size (untimed): one more iteration, its outputs gathered on rank 0 and compared
    # below with the plain-C sequential oracle run over the FULL series
    par_out = None
    if not args.no_parity:
        nl_par = it.iteration(reduce_nlml=reduce_nlml)
        outs = [it.post_mean, it.post_var, it.site_mean, it.site_var]      # site_* = the sites this iteration produced
        if world > 1:
            full = []
            for o in outs:
                g = torch.empty(world, o.numel(), dtype=o.dtype, device=dev)
                dist.all_gather_into_tensor(g, o.reshape(-1).contiguous())
                full.append(g.reshape(-1).cpu().numpy() if rank == 0 else None)
                del g
        else:
            full = [o.reshape(-1).cpu().numpy() for o in outs]
        par_out = (float(nl_par), full)
    tt = torch.tensor([ms_total, phases[0], phases[1], phases[2]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, ph = float(tt[0]), [float(x) for x in tt[1:]]
    ms_step = ms_total / args.steps
    value = N / (ms_step * 1e-3)

    # ---- end to end through the C ABI with HOST (pinned) buffers ----
    # e2e:          one training iteration as SDEGP.run() is used (sde_gp.py:179, once per optimiser step): the
    #               step's inputs dt, y go host -> device, the result nlml comes back; the site parameters are the
    #               state the iteration updates (self.sites.site_params) and stay on the device between calls.
    # e2e_all_host: the same call with the sites as host arrays too (4 arrays up, 2 arrays + nlml down per step).
    e2e = e2e_all = None
    if not args.no_e2e:
        pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
        h_dt, h_y, h_sm, h_sv = pin(sh_dt), pin(sh_y), pin(sh_y.copy()), pin(np.full(n_local, 0.5))
        o_sm, o_sv, o_nl = (torch.empty(n_local, dtype=torch.float64).pin_memory(),
                            torch.empty(n_local, dtype=torch.float64).pin_memory(), torch.empty(1, dtype=torch.float64).pin_memory())
        ke = max(3, min(args.steps, 10))

        def time_e2e(step):
            for _ in range(2):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                step()
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return N * ke / float(t[0])

        if world == 1:
            nbytes = C.c_size_t(0)
            _lib.check(_lib.lib().kjx_run_host_scratch_bytes(it.model.ref(), C.c_int64(N), 8, C.byref(nbytes)), "scratch")
            scratch = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

            def host_call(sm, sv, osm, osv, flags):
                _lib.check(_lib.lib().kjx_run_host_f64(it.model.ref(), C.c_int64(N), _lib.ptr(h_dt), _lib.ptr(h_y), sm, sv, osm, osv,
                                                       None, None, _lib.ptr(o_nl), _lib.ptr(scratch), C.c_size_t(nbytes.value),
                                                       flags, stream), "kjx_run_host")

            def all_host_step():
                host_call(_lib.ptr(h_sm), _lib.ptr(h_sv), _lib.ptr(o_sm), _lib.ptr(o_sv), 0)

            def resident_step():
                host_call(None, None, None, None, _lib.FLAG_RESIDENT_SITES)
            host_call(_lib.ptr(h_sm), _lib.ptr(h_sv), None, None, _lib.FLAG_RESIDENT_INIT)    # sites -> device, once
        else:
            def all_host_step():
                it.dt.copy_(h_dt, non_blocking=True); it.y.copy_(h_y, non_blocking=True)
                it.site_mean.copy_(h_sm, non_blocking=True); it.site_var.copy_(h_sv, non_blocking=True)
                nl = it.iteration(reduce_nlml=reduce_nlml)
                o_sm.copy_(it.site_mean, non_blocking=True); o_sv.copy_(it.site_var, non_blocking=True)
                o_nl.copy_(nl.reshape(1), non_blocking=True)
                torch.cuda.synchronize()

            def resident_step():
                it.dt.copy_(h_dt, non_blocking=True); it.y.copy_(h_y, non_blocking=True)
                nl = it.iteration(reduce_nlml=reduce_nlml)
                o_nl.copy_(nl.reshape(1), non_blocking=True)
                torch.cuda.synchronize()
        v_pipe = None
        if world == 1:
            # the same call without its final synchronisation (kjx_run_host_submit_f64): step k+1's upload, on a copy
            # stream into the other input slot, overlaps step k's kernels; every step still uploads its dt, y and has its
            # nlml read back, and step k's nlml is on the host before step k+2 is submitted
            copy_stream = torch.cuda.Stream(device=dev)
            o_nl2 = torch.empty(2, dtype=torch.float64).pin_memory()
            cs = C.c_void_p(copy_stream.cuda_stream)

            def submit(k):
                _lib.check(_lib.lib().kjx_run_host_submit_f64(it.model.ref(), C.c_int64(N), _lib.ptr(h_dt), _lib.ptr(h_y),
                                                              C.c_void_p(o_nl2.data_ptr() + 8 * (k & 1)), _lib.ptr(scratch),
                                                              C.c_size_t(nbytes.value), C.c_int32(k & 1), cs, stream), "kjx_run_host_submit")
                e = torch.cuda.Event()
                e.record()
                return e

            def pipelined(nsteps):
                evs = []
                for k in range(nsteps):
                    if k >= 2:
                        evs[k - 2].synchronize()
                    evs.append(submit(k))
                torch.cuda.synchronize()
            pipelined(4)
            barrier()
            t0 = time.perf_counter()
            pipelined(ke)
            v_pipe = N * ke / (time.perf_counter() - t0)
        v_all = time_e2e(all_host_step)
        e2e_all = {"value": v_all, "unit": "time-steps/s", "h2d_bytes_per_step": 4 * 8 * N, "d2h_bytes_per_step": 2 * 8 * N + 8,
                   "steps": ke, "nlml": float(o_nl[0]), "api": "as e2e, with the site parameters passed and returned as host arrays too"}
        v_res = time_e2e(resident_step)
        e2e_blocking = v_res
        if v_pipe is not None and v_pipe > v_res:
            v_res = v_pipe
        e2e = {"value": v_res, "unit": "time-steps/s", "h2d_bytes_per_step": 2 * 8 * N, "d2h_bytes_per_step": 8, "steps": ke,
               "blocking_call_value": e2e_blocking, "pipelined_value": v_pipe,
               "api": ("kjx_run_host_submit_f64 (step k+1's upload overlaps step k's kernels; kjx_run_host_f64, which returns "
                       "after the step has drained, is `blocking_call_value`)" if world == 1 and v_pipe is not None and v_pipe >= e2e_blocking
                       else "kjx_run_host_f64" if world == 1 else "pinned copies + sharded C-ABI calls per rank") +
                      ": dt, y from pinned host memory every step, nlml read back; the site parameters (the state the "
                      "iteration updates, self.sites.site_params) stay on the device between iterations",
               "nlml": float(o_nl[0])}

    def finish():
        # tear down without destroying the process group: the captured graphs hold NCCL work and
        # destroy_process_group() can block on them; every rank is past its last collective here
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        return finish()
    hbm, peak_src = peaks()
    dom = int(np.argmax(ph[1:])) + 1                        # filter-apply or smoother-apply phase
    traffic = None                                          # dram bytes per launch from the committed ncu --set full capture
    try:
        if world == 1 and args.log2n == 24:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")))["dram_bytes_per_launch"]
            traffic = tr[("fused_filter_kernel", "fused_smoother_kernel")[dom - 1]]
    except Exception:
        pass
    dom_bytes = (BYTES_FILTER, BYTES_SMOOTHER)[dom - 1] * n_local
    achieved = dom_bytes / (ph[dom] * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "time-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C5: synthetic Matern-5/2 (var 1, len 5) + Gaussian(0.5) + EP(power 0.5); one iteration = "
                               "filter(stored sites) -> RTS smoother -> site update; N=2^%d steps, d=3, f=1" % args.log2n,
                   "N": N, "state_dim": 3, "sharding": "time axis, %d contiguous shard(s)" % world,
                   "exchange": exchange,
                   "launch": ("CUDA graph replay, %d iterations per graph launch (sites ping/pong between two buffers)" % max(2 * pairs, 1))
                             if graphed else "eager (ctypes calls per iteration)",
                   "l2_policy": "inputs (%.1f GB per iteration per GPU) exceed the 126 MB L2; no flush needed" % (BYTES_PER_STEP * n_local / 1e9)},
        "roofline": {"bound": "hbm", "kernel": ("fused_filter_kernel (R3)", "fused_smoother_kernel (R4)")[dom - 1],
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": (BYTES_FILTER, BYTES_SMOOTHER)[dom - 1],
                     "phase_ms": {"reduce+scan+fold(R1,R2)": ph[0], "filter(R3)": ph[1], "smoother+sites(R4)": ph[2]},
                     "iteration": {"achieved": BYTES_PER_STEP * n_local / (ms_step * 1e-3) / 1e9,
                                   "frac": BYTES_PER_STEP * n_local / (ms_step * 1e-3) / 1e9 / hbm, "bytes_per_step": BYTES_PER_STEP}},
        "gpu_launches": it.launches_per_iteration * args.steps,
        "clocks": clocks,
        "nlml": nlml,
    }
    if e2e:
        line["e2e"] = e2e
        line["e2e_all_host"] = e2e_all
    if par_out is not None:
        lib = cpu_oracle_lib()
        t0 = time.perf_counter()
        want = cpu_oracle_full(lib, dt, y, y.copy(), np.full(N, 0.5))
        t_or = time.perf_counter() - t0
        got_nl, (g_pm, g_pv, g_sm, g_sv) = par_out
        par = {"oracle": "oracle/seq_filter.c, one sequential chain over all N = 2^%d steps (%.1f s on 1 host core)" % (args.log2n, t_or),
               "nlml_rel": abs(got_nl - want["nlml"]) / abs(want["nlml"]),
               "post_mean_rel": rel_err(g_pm, want["pm"]), "post_var_rel": rel_err(g_pv, want["pv"]),
               "site_rel": max(rel_err(g_sm, want["nsm"]), rel_err(g_sv, want["nsv"])),
               "tol": 1e-9, "norm": "max|a-b| / max|b| per array"}
        par["ok"] = bool(max(par["nlml_rel"], par["post_mean_rel"], par["post_var_rel"], par["site_rel"]) <= par["tol"])
        line["parity"] = par
        del want
    if world == 1 and not args.no_extra:
        # free the headline workload's buffers first
        del it, graphed
        torch.cuda.empty_cache()
        pk = fp64_peak()
        line["fp64_peak"] = pk
        try:
            line["extra_workloads"] = extra_workloads(pk)
        except Exception as e:
            line["extra_workloads"] = {"error": str(e).splitlines()[0]}
    if not args.no_cpu:
        lib = cpu_oracle_lib()
        n_s = 1 << 22
        v = cpu_chain(lib, dt[:n_s], y[:n_s], reps=3)
        line["cpu_baseline"] = {"value": v, "unit": "time-steps/s", "cores": 1, "kind": "port",
                                "sample": "first 2^22 steps of the same series, best of 3, C restatement of the reference's "
                                          "sequential loops (oracle/seq_filter.c); the recursion is one dependent chain"}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
