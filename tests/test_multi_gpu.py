"""The multi-GPU data plane on real peers (skipped with fewer than 2 GPUs): time shards over torchrun ranks, the
in-kernel exchange (kjx_sharded_iteration_f64), the separate exchange+fold kernel (kjx_exchange_fold_f64) and the NCCL
all-gather form, eager and as replayed CUDA graphs, against the sequential oracle at 1e-9.  Also the one-shard form of
the linked iteration, which runs on a single GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from golden_util import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_iteration_on_real_peers(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert "MGPU_OK" in out.stdout, (out.stdout[-3000:], out.stderr[-3000:])


@pytest.mark.parametrize("prior_name,N", [("Matern52", 100_003), ("Matern32", 4097), ("Matern52", 5)])
def test_linked_iteration_single_shard_vs_oracle(prior_name, N):
    """kjx_sharded_iteration_f64 with world = 1 (ordinary device buffer): same kernels, self-exchange."""
    import kalman_jax_b200 as kjx
    from bench import synth
    from kalman_jax_b200.sharded import ShardedIteration, GraphedIteration, PeerExchange
    dt, y = synth(N)
    t = np.cumsum(dt)
    rng = np.random.default_rng(5)
    sm0, sv0 = y + 0.1 * rng.standard_normal(N), rng.uniform(0.3, 0.8, N)
    ref = oracle.SDEGP(getattr(oracle.priors, prior_name)(1.0, 5.0), oracle.likelihoods.Gaussian(0.5), t, y,
                       approx_inf=oracle.approx_inf.EP(power=0.5, damping=0.7))
    ref.sites.site_params = (sm0.reshape(N, 1, 1).copy(), sv0.reshape(N, 1, 1).copy())
    dev = torch.device("cuda", 0)
    it = ShardedIteration(getattr(kjx.priors, prior_name)(1.0, 5.0), kjx.likelihoods.Gaussian(0.5),
                          kjx.approximate_inference.EP(power=0.5, damping=0.7), dt, y, 0.0, 0, 1, dev,
                          site_mean=sm0.copy(), site_var=sv0.copy(), peer=PeerExchange(1, 0, dev))
    assert it.linked and it.launches_per_iteration == 5
    for k in range(4):
        if k == 2:                       # from here on as replayed graphs (capture warms up with two more iterations)
            keep = (it.site_mean.clone(), it.site_var.clone())
            g = GraphedIteration(it)
            it.site_mean.copy_(keep[0]); it.site_var.copy_(keep[1])
            g.parity = 0
        nl = float(it.iteration() if k < 2 else g.replay())
        newest = (it.site_mean, it.site_var) if (k < 2 or k == 3) else (it.new_site_mean, it.new_site_var)
        w_nl, (fm, fP, s_f) = ref.kalman_filter(ref.y, ref.dt, True, None, ref.sites.site_params, ref.r)
        s_new, pm, pv = ref.rauch_tung_striebel_smoother(fm, fP, ref.dt, True, False, ref.y, s_f, ref.r)
        ref.sites.site_params = s_new
        assert abs(nl - w_nl) <= 1e-9 * abs(w_nl)
        assert rel_err(it.post_mean.cpu().numpy(), pm.reshape(-1)) < 1e-9
        assert rel_err(it.post_var.cpu().numpy(), pv.reshape(-1)) < 1e-9
        assert rel_err(newest[0].cpu().numpy(), s_new[0].reshape(-1)) < 1e-9
        assert rel_err(newest[1].cpu().numpy(), s_new[1].reshape(-1)) < 1e-9
