#!/usr/bin/env python
"""Kernel timeline (CUPTI through torch.profiler) of graphed shard iterations: start offset, duration and the gap to
the previous kernel, for every kernel of a few replays.  Diagnostic only (numbers taken under a profiler are never
bench values).

    python scripts/timeline.py [--log2n 21] [--iters 3]                  one GPU, the whole series as one shard
    torchrun --nproc-per-node G scripts/timeline.py --log2n 24 ...       G shards, rank 0 prints
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kalman_jax_b200 as kjx  # noqa: E402
from bench import synth  # noqa: E402
from kalman_jax_b200.sharded import ShardedIteration, GraphedIteration, split_series  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=21)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--exchange", default="peer")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gather = reduce_nlml = peer = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from bench import make_collectives
        gather, reduce_nlml, peer = make_collectives(world, rank, dev, args.exchange)
    elif args.exchange != "nccl":
        from kalman_jax_b200.sharded import PeerExchange
        peer = PeerExchange(1, 0, dev)
    linked = peer is not None and args.exchange == "peer"
    if linked:
        reduce_nlml = None
    N = 1 << args.log2n
    dt, y = synth(N)
    sh_dt, sh_y, dt_next, lo, hi = split_series(dt, y, world)[rank]
    it = ShardedIteration(kjx.priors.Matern52(1.0, 5.0), kjx.likelihoods.Gaussian(0.5), kjx.approximate_inference.EP(power=0.5),
                          sh_dt, sh_y, dt_next, rank, world, dev, gather, site_mean=sh_y.copy(),
                          site_var=np.full(hi - lo, 0.5), peer=peer, linked=linked)
    g = GraphedIteration(it, reduce_nlml)
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.iters):
            g.replay()
        torch.cuda.synchronize()
    if rank == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        ev.sort(key=lambda e: e.time_range.start)
        t0, prev_end = ev[0].time_range.start, None
        print("%-70s %10s %9s %8s" % ("kernel", "start_us", "dur_us", "gap_us"))
        for e in ev:
            s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
            gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
            print("%-70s %10.1f %9.1f %8.1f" % (e.name[:70], s, d, gap))
            prev_end = max(prev_end or 0, e.time_range.end)
    if world > 1:
        torch.cuda.synchronize()
        import torch.distributed as dist
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
