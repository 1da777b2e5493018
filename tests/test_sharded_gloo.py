"""The N > 1 path on CPU: two `gloo` ranks, each owning a contiguous shard of the time axis.  Every rank reduces its
shard to one filter scan element (product math compiled for the host, tests/host/kjx_hostemu.cu), the elements are
all-gathered with torch.distributed, folded with the library's own host entry point
(kjx_fold_shard_summaries_host, include/kjx.h), and each rank then filters / smooths its shard from the boundary
state / information it received.  The stitched result must equal the single-chain oracle."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind_name, N, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_host_emulation import _lib as hostemu_lib, _dp
    from kalman_jax_b200 import _lib as K, priors, likelihoods, approximate_inference as ai
    from kalman_jax_b200.sde_gp import _Model
    from kalman_jax_b200.sharded import split_series
    emu = hostemu_lib()
    rng = np.random.default_rng(21)
    dt = rng.uniform(0.05, 0.15, N); dt[0] = 0.0
    t = np.cumsum(dt)
    y = np.cos(0.04 * t + 0.33 * np.pi) * np.sin(0.2 * t) + np.sqrt(0.15) * rng.standard_normal(N)
    site_var = rng.uniform(0.3, 0.9, N)
    prior = getattr(priors, kind_name)(1.0, 5.0)
    model = _Model(prior, likelihoods.Gaussian(0.5), ai.EP(power=0.5))
    d = model.state_dim
    kind, lam = prior.blocks()[0][0], prior.blocks()[0][2]
    sh_dt, sh_y, _, lo, hi = split_series(dt, y, world)[rank]
    n = hi - lo
    R = np.ascontiguousarray(site_var[lo:hi])
    ns = int(K.lib().kjx_filter_summary_doubles(d))
    mine = np.zeros(ns)
    assert emu.emu_shard_summary(kind, C.c_double(lam), _dp(model.Pinf), C.c_int64(n), _dp(np.ascontiguousarray(sh_dt)),
                                 _dp(np.ascontiguousarray(sh_y)), _dp(R), _dp(mine)) == 0
    gathered = [torch.zeros(ns, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))                     # the one collective of an iteration
    allsum = np.ascontiguousarray(torch.stack(gathered).numpy())
    state, info = np.zeros(d + d * d), np.zeros(d + d * d)
    assert K.lib().kjx_fold_shard_summaries_host(model.ref(), _dp(allsum), C.c_int64(ns), world, rank, _dp(state), _dp(info)) == 0
    fm, fP, sm, sP = np.zeros([n, d]), np.zeros([n, d, d]), np.zeros([n, d]), np.zeros([n, d, d])
    assert emu.emu_shard_apply(kind, C.c_double(lam), _dp(model.Pinf), C.c_int64(n), _dp(np.ascontiguousarray(sh_dt)),
                               _dp(np.ascontiguousarray(sh_y)), _dp(R), _dp(state), _dp(info), _dp(fm), _dp(fP), _dp(sm), _dp(sP)) == 0
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), fm=fm, fP=fP, sm=sm, sP=sP, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind_name,world", [("Matern52", 2), ("Matern32", 3)])
def test_two_rank_gloo_sharded_iteration_matches_oracle(tmp_path, kind_name, world):
    import oracle
    from test_host_emulation import _lib as hostemu_lib
    hostemu_lib()                                  # build once, before the ranks race for it
    N = 4001
    mp.spawn(_worker, args=(world, _free_port(), kind_name, N, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(21)
    dt = rng.uniform(0.05, 0.15, N); dt[0] = 0.0
    t = np.cumsum(dt)
    y = np.cos(0.04 * t + 0.33 * np.pi) * np.sin(0.2 * t) + np.sqrt(0.15) * rng.standard_normal(N)
    site_var = rng.uniform(0.3, 0.9, N)
    ref = oracle.SDEGP(getattr(oracle.priors, kind_name)(1.0, 5.0), oracle.likelihoods.Gaussian(0.5), t, y,
                       approx_inf=oracle.approx_inf.EP(power=0.5))
    _, (wfm, wfP, _) = ref.kalman_filter(ref.y, ref.dt, True, None, (y[:, None, None], site_var[:, None, None]), ref.r)
    _, wsm, wsP = ref.rauch_tung_striebel_smoother(wfm, wfP, ref.dt, True, True, None, None, ref.r)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [(r * N) // world for r in range(world)]
    for key, want in (("fm", wfm[:, :, 0]), ("fP", wfP), ("sm", wsm[:, :, 0]), ("sP", wsP)):
        got = np.concatenate([p[key] for p in parts])
        assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-10, key
