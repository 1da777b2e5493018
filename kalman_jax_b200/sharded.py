"""Time-sharded filter / smoother iteration: one contiguous block of time steps per GPU (one process per
GPU) and ONE fixed-size all-gather per iteration (tightened by the two-filter identity):

    local fold + scan -> the shard's filter element (A,b,C,eta,J), 27 doubles at d = 3
            --all-gather-->  fold ranks < me -> filtered state entering my shard
                             fold ranks > me -> information (eta, J) about my last state
    local filter recursion (log Z, filtered moments parked in the workspace)
    local RTS recursion backwards + site update

No bulk data ever crosses NVLink.  The shards' partial nlml values are summed with one scalar all-reduce
off the critical path.  With world == 1 the same calls run without any collective.  The reference has no
counterpart (it is a single sequential loop, sde_gp.py:271,349).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .sde_gp import _Model


class PeerExchange(object):
    """Receive buffers of the in-kernel exchange (kjx_sharded_iteration / kjx_exchange_fold): symmetric memory that
    every rank's kernels can write over NVLink.  world == 1 needs no peer: an ordinary device buffer."""
    STRIDE = 32

    def __init__(self, world, rank, device, group=None):
        n = int(_lib.lib().kjx_shard_link_doubles(C.c_int32(world), C.c_int64(self.STRIDE)))
        if world > 1:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = group or dist.group.WORLD
            self.buf = symm_mem.empty(n, dtype=torch.float64, device=device)
            self.buf.zero_()
            self.handle = symm_mem.rendezvous(self.buf, group)
            torch.cuda.synchronize(device)
            dist.barrier(group)                              # every buffer is zeroed before anyone writes
            ptrs = list(self.handle.buffer_ptrs)
        else:
            self.buf = torch.zeros(n, dtype=torch.float64, device=device)
            ptrs = [self.buf.data_ptr()]
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.my_ptr = int(ptrs[rank])                                     # this rank's own buffer as every kernel addresses it
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)     # kjx_exchange_fold (separate-kernel form)
        self.ctrl = torch.zeros(4, dtype=torch.int64, device=device)      # kjx_sharded_iteration: epoch + three tickets
        self.world, self.rank = world, rank


class ShardedIteration(object):
    """Holds one rank's shard of (dt, y, sites) resident in HBM and runs run()-style iterations on it.

    gather(t) must return the all-gathered [world, t.numel()] tensor on the same device
    (torch.distributed.all_gather_into_tensor); it is None for world == 1.
    """
    def __init__(self, prior, likelihood, sites, dt_shard, y_shard, dt_next, rank, world, device, gather=None,
                 site_mean=None, site_var=None, peer=None, linked=None):
        """peer: PeerExchange or None (NCCL all-gather through `gather`).  linked (default: whenever there is a peer):
        the whole iteration is ONE library call, kjx_sharded_iteration — exchange, fold and nlml sum inside the scan /
        filter / smoother kernels; otherwise the four separate calls with kjx_exchange_fold or gather + fold between."""
        self.model = _Model(prior, likelihood, sites)
        self.peer = peer
        self.linked = (peer is not None) if linked is None else bool(linked)
        assert not self.linked or peer is not None
        self.rank, self.world, self.device, self.gather = rank, world, torch.device(device), gather
        self.dt = torch.as_tensor(dt_shard, dtype=torch.float64, device=self.device).contiguous()
        self.y = torch.as_tensor(y_shard, dtype=torch.float64, device=self.device).contiguous().reshape(-1)
        self.N = self.dt.numel()
        self.dt_next = float(dt_next)
        d = self.model.state_dim
        self.d = d
        z = lambda *s: torch.empty(*s, dtype=torch.float64, device=self.device)  # noqa: E731
        self.site_mean = None if site_mean is None else torch.as_tensor(site_mean, dtype=torch.float64, device=self.device).contiguous()
        self.site_var = None if site_var is None else torch.as_tensor(site_var, dtype=torch.float64, device=self.device).contiguous()
        self.new_site_mean, self.new_site_var = z(self.N), z(self.N)
        self.post_mean, self.post_var = z(self.N), z(self.N)
        L = _lib.lib()
        self.fns = int(L.kjx_filter_summary_doubles(d))
        self.fmsg = torch.zeros(self.fns, dtype=torch.float64, device=self.device)
        self.f_state, self.s_info = z(d + d * d), z(d + d * d)
        nbytes = C.c_size_t(0)
        _lib.check(L.kjx_workspace_bytes(self.model.ref(), C.c_int64(self.N), 8, C.byref(nbytes)), "kjx_workspace_bytes")
        self.ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=self.device)   # owned: carries the
        # per-chunk prefixes and the packed filtered moments between the phases of an iteration
        self.nlml_part = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.nlml = self.nlml_part
        self.nlml_all = torch.zeros(1, dtype=torch.float64, device=self.device)
        # kernels per iteration.  linked: fold, level-0 scan, top scan (+ publish) | filter | smoother = 5;
        # separate calls: fold, 3 segment scans | exchange-fold | filter, nlml-reduce | smoother = 8
        self.launches_per_iteration = 5 if self.linked else 8

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def iteration(self, update_sites=True, events=None, reduce_nlml=None):
        """One filter -> smoother -> site-update pass over this shard.  `events`: optional list of 4
        torch.cuda.Event recorded at the phase boundaries (reduce+scan | filter | smoother).
        reduce_nlml(t): sums the shards' partial nlml (all-reduce) — None for world == 1."""
        L, m, st = _lib.lib(), self.model.ref(), self._stream()
        N, p, wsb = C.c_int64(self.N), _lib.ptr, C.c_size_t(self.ws.numel())
        if self.linked:
            flags = 0 if update_sites else _lib.FLAG_NO_SITE_UPDATE
            _lib.check(L.kjx_sharded_iteration_f64(m, N, p(self.dt), p(self.y), p(self.site_mean), p(self.site_var),
                                                   p(self.peer.ptrs), C.c_void_p(self.peer.my_ptr), C.c_int64(self.peer.STRIDE), self.world, self.rank,
                                                   p(self.peer.ctrl), p(self.post_mean), p(self.post_var),
                                                   p(self.new_site_mean), p(self.new_site_var), p(self.nlml_all), p(self.ws),
                                                   wsb, C.c_uint32(flags), st), "kjx_sharded_iteration")
            self.nlml = self.nlml_all
            if update_sites:
                self._swap_sites()
            return self.nlml
        if events:
            events[0].record()
        _lib.check(L.kjx_filter_summary_f64(m, N, p(self.dt), p(self.y), p(self.site_mean), p(self.site_var),
                                            p(self.fmsg), p(self.ws), wsb, st), "kjx_filter_summary")
        if self.peer is not None:
            _lib.check(L.kjx_exchange_fold_f64(m, p(self.fmsg), p(self.peer.ptrs), C.c_int64(self.peer.STRIDE), self.world,
                                               self.rank, p(self.peer.epoch), p(self.f_state), p(self.s_info), st),
                       "kjx_exchange_fold")
        else:
            allf = self.gather(self.fmsg) if self.gather else self.fmsg.reshape(1, -1)
            _lib.check(L.kjx_fold_shard_summaries_f64(m, p(allf), C.c_int64(allf.shape[1]), self.world, self.rank,
                                                      p(self.f_state), p(self.s_info), st), "kjx_fold_shard_summaries")
        if events:
            events[1].record()
        _lib.check(L.kjx_filter_apply_f64(m, N, p(self.dt), p(self.y), p(self.site_mean), p(self.site_var),
                                          p(self.f_state), p(self.nlml_part), p(self.ws), wsb, st), "kjx_filter_apply")
        if events:
            events[2].record()
        flags = 0 if update_sites else _lib.FLAG_NO_SITE_UPDATE
        _lib.check(L.kjx_smoother_apply_f64(m, N, p(self.dt), p(self.y), p(self.site_mean), p(self.site_var),
                                            p(self.s_info), p(self.post_mean), p(self.post_var), p(self.new_site_mean),
                                            p(self.new_site_var), p(self.ws), wsb, C.c_uint32(flags), st), "kjx_smoother_apply")
        if events:
            events[3].record()
        self.nlml = reduce_nlml(self.nlml_part) if reduce_nlml else self.nlml_part
        if update_sites:
            self._swap_sites()
        return self.nlml

    def _swap_sites(self):
        if self.site_mean is None:
            self.site_mean, self.site_var = torch.empty_like(self.new_site_mean), torch.empty_like(self.new_site_var)
        self.site_mean, self.new_site_mean = self.new_site_mean, self.site_mean
        self.site_var, self.new_site_var = self.new_site_var, self.site_var


class GraphedIteration(object):
    """CUDA graphs replaying ShardedIteration.iteration() without any host work per step: at 8 GPUs a shard's iteration
    is ~200 us of device time, the same order as the Python / ctypes launch path, so the launch sequence is captured
    once and replayed.  The sites ping-pong between two buffers, so the natural unit is a PAIR of iterations; `pairs`
    of them go into one graph (replay_many), plus two single-iteration graphs (ping -> pong, pong -> ping) for replay()."""
    def __init__(self, it, reduce_nlml=None, pairs=0):
        self.it, self.graphs, self.parity = it, [], 0
        assert it.site_mean is not None
        stream = torch.cuda.Stream(device=it.device)
        stream.wait_stream(torch.cuda.current_stream(it.device))
        with torch.cuda.stream(stream):
            for _ in range(2):                                   # warm up on the side stream (NCCL, allocator)
                it.iteration(reduce_nlml=reduce_nlml)
        torch.cuda.current_stream(it.device).wait_stream(stream)
        torch.cuda.synchronize(it.device)
        self.nlml = []
        for _ in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                self.nlml.append(it.iteration(reduce_nlml=reduce_nlml))
            self.graphs.append(g)
        self.pairs, self.many = int(pairs), None
        if self.pairs > 0:
            self.many = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.many, stream=stream):
                for _ in range(2 * self.pairs):
                    self.many_nlml = it.iteration(reduce_nlml=reduce_nlml)

    def replay(self):
        self.graphs[self.parity].replay()
        out = self.nlml[self.parity]
        self.parity ^= 1
        return out

    def replay_many(self):
        """2 * pairs iterations in one graph launch (only from parity 0: an even number of iterations so far)."""
        assert self.many is not None and self.parity == 0
        self.many.replay()
        return self.many_nlml


def split_series(dt, y, world):
    """Contiguous time blocks: rank g owns steps [g N / G, (g+1) N / G); also dt of the next shard's first step."""
    N = len(dt)
    bounds = [(g * N) // world for g in range(world + 1)]
    shards = []
    for g in range(world):
        lo, hi = bounds[g], bounds[g + 1]
        shards.append((np.asarray(dt[lo:hi]), np.asarray(y[lo:hi]), float(dt[hi]) if hi < N else 0.0, lo, hi))
    return shards
