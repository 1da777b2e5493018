"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), G contiguous time shards of one series.

Runs three iterations of the time-sharded run() on real peers in each exchange mode — `linked`
(kjx_sharded_iteration_f64: exchange inside the scan / filter / smoother kernels over symmetric memory), `separate`
(kjx_exchange_fold_f64 between the phase calls) and `nccl` (all-gather + kjx_fold_shard_summaries_f64) — eagerly and as
replayed CUDA graphs, gathers every shard's outputs on rank 0 and compares them with the sequential oracle (1e-9).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kalman_jax_b200 as kjx  # noqa: E402
import oracle  # noqa: E402
from bench import make_collectives, synth  # noqa: E402
from golden_util import rel_err  # noqa: E402
from kalman_jax_b200.sharded import ShardedIteration, GraphedIteration, split_series  # noqa: E402

TOL = 1e-9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_003          # ragged: the shards differ in length
    dt, y = synth(N)
    t = np.cumsum(dt)
    rng = np.random.default_rng(5)
    sm0, sv0 = y + 0.1 * rng.standard_normal(N), rng.uniform(0.3, 0.8, N)
    iters, damping = 3, 0.7
    mk = lambda mod, inf: (getattr(mod.priors, "Matern52")(1.0, 5.0), mod.likelihoods.Gaussian(0.5), inf.EP(power=0.5, damping=damping))  # noqa: E731
    want = None
    if rank == 0:
        ref = oracle.SDEGP(*mk(oracle, oracle.approx_inf)[:2], t, y, approx_inf=mk(oracle, oracle.approx_inf)[2])
        ref.sites.site_params = (sm0.reshape(N, 1, 1).copy(), sv0.reshape(N, 1, 1).copy())
        want = []
        for _ in range(iters):
            nl, (fm, fP, s_f) = ref.kalman_filter(ref.y, ref.dt, True, None, ref.sites.site_params, ref.r)
            s_new, pm, pv = ref.rauch_tung_striebel_smoother(fm, fP, ref.dt, True, False, ref.y, s_f, ref.r)
            ref.sites.site_params = s_new
            want.append((nl, pm.reshape(-1), pv.reshape(-1), s_new[0].reshape(-1), s_new[1].reshape(-1)))
    sh_dt, sh_y, dt_next, lo, hi = split_series(dt, y, world)[rank]
    counts = [b[4] - b[3] for b in split_series(dt, y, world)]
    failures = []
    for mode in ("linked", "separate", "nccl"):
        for graphed in (False, True):
            gather, reduce_nlml, peer = make_collectives(world, rank, dev, "nccl" if mode == "nccl" else "peer")
            assert mode == "nccl" or peer is not None, "symmetric memory unavailable"
            linked = mode == "linked"
            prior, lik, rule = mk(kjx, kjx.approximate_inference)
            it = ShardedIteration(prior, lik, rule, sh_dt, sh_y, dt_next, rank, world, dev, gather,
                                  site_mean=sm0[lo:hi].copy(), site_var=sv0[lo:hi].copy(), peer=peer, linked=linked)
            red = None if linked else reduce_nlml
            if graphed:
                # GraphedIteration warms up with 2 iterations: restore the starting sites afterwards, then replay
                g = GraphedIteration(it, red)
                it.site_mean.copy_(torch.as_tensor(sm0[lo:hi], device=dev)); it.site_var.copy_(torch.as_tensor(sv0[lo:hi], device=dev))
                g.parity = 0
                step = g.replay
            else:
                step = lambda: it.iteration(reduce_nlml=red)  # noqa: E731
            for k in range(iters):
                nl = float(step())
                # after a step the newest sites are it.site_* when eager (python swap) and alternate ping/pong when graphed
                if graphed:
                    nsm, nsv = (it.new_site_mean, it.new_site_var) if k % 2 == 0 else (it.site_mean, it.site_var)
                    # captured graph 0 wrote into the tensors that were `new_site_*` at ITS capture; resolve by value below
                outs = []
                src = [it.post_mean, it.post_var] + ([it.site_mean, it.site_var] if not graphed else [nsm, nsv])
                for o in src:
                    pad = torch.zeros(max(counts), dtype=torch.float64, device=dev)
                    pad[: o.numel()] = o.reshape(-1)
                    full = torch.empty(world, max(counts), dtype=torch.float64, device=dev)
                    dist.all_gather_into_tensor(full, pad)
                    outs.append(np.concatenate([full[r, : counts[r]].cpu().numpy() for r in range(world)]))
                if rank == 0:
                    w = want[k]
                    errs = dict(nlml=abs(nl - w[0]) / abs(w[0]), pm=rel_err(outs[0], w[1]), pv=rel_err(outs[1], w[2]),
                                sm=rel_err(outs[2], w[3]), sv=rel_err(outs[3], w[4]))
                    bad = {k2: v for k2, v in errs.items() if not (v <= TOL)}
                    if bad:
                        failures.append((mode, graphed, k, bad))
            del it
            torch.cuda.synchronize()
            dist.barrier()
    if rank == 0:
        print("MGPU_FAILURES", failures)
        print("MGPU_OK" if not failures else "MGPU_FAIL")
    sys.stdout.flush()
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
