"""SDEGP — drop-in mirror of /root/reference/kalmanjax/sde_gp.py:12-384 whose Kalman filter, RTS smoother
and site updates run in hand-written sm_100a CUDA kernels (libkjx.so, include/kjx.h).

Same constructor and method names / return tuples as the reference:
    SDEGP(prior, likelihood, t, y, r=None, approx_inf=None)
    .run(params=None)                       -> (neg_log_marg_lik, dlZ)           sde_gp.py:163-188
    .run_two_stage(params=None)             -> (neg_log_marg_lik, dlZ)           sde_gp.py:190-219
    .neg_log_marg_lik(params=None)          -> (neg_log_marg_lik, dlZ)           sde_gp.py:145-161
    .kalman_filter(y, dt, params, store, mask, site_params, r)                   sde_gp.py:230-309
    .rauch_tung_striebel_smoother(params, m_filtered, P_filtered, dt, store, return_full, y, site_params, r)
                                                                                 sde_gp.py:311-384
    .predict(t, r) / .predict_everywhere(...) / .negative_log_predictive_density(t, y, r)
                                                                                 sde_gp.py:52-143
Arrays returned by the filter / smoother are torch CUDA tensors in the reference's shapes
([N,d,1], [N,d,d], [N,f,1], [N,f,f]); the training data, step sizes and sites stay resident in HBM
between iterations.  PyTorch is used for device memory and streams only.

Gradients: the reference differentiates the filter energy with jax autodiff (sde_gp.py:158,179,216).
Here `dlZ` is a central finite difference over the (fast) energy-only filter — SURVEY.md 8(f) row 1;
pass compute_grad=False to skip the 2 x (#hyperparameters) extra filter passes.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .approximate_inference import EP
from .utils import softplus, softplus_list, input_admin, test_input_admin

pi = 3.141592653589793


# ---------------------------------------------------------------------------------------------------
def _np_dtype(dtype):
    return np.float64 if dtype == torch.float64 else np.float32


def _as_dev(x, dtype, device):
    """numpy / torch -> contiguous device tensor of the compute dtype (no copy if already there)."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device=device)


def _theta_prior(prior, params_prior):
    return softplus_list(params_prior) if isinstance(params_prior, list) else softplus(params_prior)


class _Model(object):
    """kjx_model_t plus the host arrays it points to (kept alive for the duration of the calls)."""
    def __init__(self, prior, likelihood, sites, theta_prior=None, theta_lik=None, H_steps=None, r0=None,
                 cubature=None, power=None):
        m = _lib.KjxModel()
        _, _, _, H, Pinf = prior.kernel_to_state_space(theta_prior)
        blocks = prior.blocks(theta_prior)
        if len(blocks) > _lib.KJX_MAX_BLOCKS:
            raise _lib.KjxError("prior has more than %d diagonal blocks" % _lib.KJX_MAX_BLOCKS)
        self.Pinf = np.ascontiguousarray(Pinf, dtype=np.float64)
        m.state_dim = self.Pinf.shape[0]
        m.n_blocks = len(blocks)
        for i, (kind, count, lam, omega) in enumerate(blocks):
            m.blocks[i].kind, m.blocks[i].count, m.blocks[i].lam, m.blocks[i].omega = kind, count, lam, omega
        m.Pinf = self.Pinf.ctypes.data
        if H_steps is not None:
            self.H = None
            self.H_steps = H_steps
            m.H = None
            m.H_steps = H_steps.data_ptr()
            m.func_dim = H_steps.shape[1]
        else:
            if H is None:
                H = prior.measurement_model(r0, theta_prior)
            self.H = np.ascontiguousarray(H, dtype=np.float64)
            m.H = self.H.ctypes.data
            m.H_steps = None
            m.func_dim = self.H.shape[0]
        m.likelihood = likelihood.kjx_code
        m.method = sites.kjx_code
        m.lik_hyp = likelihood.lik_hyp(theta_lik)
        m.power = float(sites.power if power is None else power)
        m.damping = float(sites.damping)
        x, w = (sites.cubature_func(1) if cubature is None else cubature)
        x, w = np.asarray(x, dtype=np.float64).reshape(-1), np.asarray(w, dtype=np.float64).reshape(-1)
        if x.size > _lib.KJX_MAX_CUBATURE:
            raise _lib.KjxError("more than %d cubature points" % _lib.KJX_MAX_CUBATURE)
        m.n_cub = x.size
        # which two-dimensional rule stands behind the table (f = 2 with the SLR family; kjx.h: cub_rule)
        m.cub_rule = 0 if cubature is not None else {'UT3': 3, 'UT5': 5, 'UT': 5}.get(getattr(sites, 'intmethod', 'GH'), 0)
        for i in range(x.size):
            m.cub_x[i], m.cub_w[i] = x[i], w[i]
        self.c = m
        self.state_dim, self.func_dim = m.state_dim, m.func_dim

    def ref(self):
        return C.byref(self.c)


_WS_CACHE = {}


def _workspace(model, N, dtype, device):
    """caller-owned scratch for one call; cached per (device, size) so steady-state iterations do not allocate."""
    nbytes = C.c_size_t(0)
    _lib.check(_lib.lib().kjx_workspace_bytes(model.ref(), C.c_int64(N), 8 if dtype == torch.float64 else 4,
                                              C.byref(nbytes)), "kjx_workspace_bytes")
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes.value:
        ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws, nbytes.value


def _fn(name, dtype):
    return getattr(_lib.lib(), "%s_%s" % (name, "f64" if dtype == torch.float64 else "f32"))


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def site_update(likelihood, sites, y, post_mean, post_cov, hyp=None, site_params=None, force_power=None,
                dtype=torch.float64, device="cuda"):
    """Batched ApproxInf.update on the GPU (kjx_site_update): one entry per time step; scalar sites, or 2 x 2 sites for a
    two-latent likelihood (post_mean [N,2,1], post_cov [N,2,2])."""
    f = int(getattr(likelihood, "func_dim", 1))

    class _NoPrior(object):   # kjx_site_update ignores the prior; give the model a trivial one with f function values
        def kernel_to_state_space(self, h=None):
            return None, None, None, np.kron(np.eye(f), np.array([[1.0, 0.0]])), np.eye(2 * f)

        def blocks(self, h=None):
            return [(_lib.BLOCK_MATERN32, f, 1.0, 0.0)]
    for a in (post_mean, post_cov, y):
        if isinstance(a, torch.Tensor):
            device = a.device if a.is_cuda else device
            dtype = a.dtype if a.dtype in (torch.float32, torch.float64) else dtype
            break
    model = _Model(_NoPrior(), likelihood, sites, None, hyp, power=force_power)
    y_d = _as_dev(y, dtype, device).reshape(-1)
    m_d = _as_dev(post_mean, dtype, device).reshape(-1)
    v_d = _as_dev(post_cov, dtype, device).reshape(-1)
    N = y_d.numel()
    sm_prev = sv_prev = None
    if site_params is not None:
        sm_prev = _as_dev(site_params[0], dtype, device).reshape(-1)
        sv_prev = _as_dev(site_params[1], dtype, device).reshape(-1)
    lml = torch.empty(N, dtype=dtype, device=device)
    sm = torch.empty(N * f, dtype=dtype, device=device)
    sv = torch.empty(N * f * f, dtype=dtype, device=device)
    _lib.check(_fn("kjx_site_update", dtype)(model.ref(), C.c_int64(N), _lib.ptr(y_d), _lib.ptr(m_d), _lib.ptr(v_d),
                                             _lib.ptr(sm_prev), _lib.ptr(sv_prev), _lib.ptr(lml), _lib.ptr(sm),
                                             _lib.ptr(sv), _stream(y_d.device)), "kjx_site_update")
    if f > 1:
        return lml, sm.reshape(N, f, 1), sv.reshape(N, f, f)
    return lml, sm, sv


# ---------------------------------------------------------------------------------------------------
class SDEGP(object):
    """The SDE form of a GP model with Kalman filtering / smoothing on the GPU (sde_gp.py:12-50)."""
    def __init__(self, prior, likelihood, t, y, r=None, approx_inf=None, dtype=torch.float64, device="cuda",
                 verbose=False):
        (self.t, self.y, self.r, self.dt) = input_admin(t, y, r)
        self.prior = prior
        self.likelihood = likelihood
        self.dtype, self.device = dtype, torch.device(device)
        if verbose:
            print('building SDE-GP with', self.prior.name, 'prior and', self.likelihood.name, 'likelihood ...')
        self.F, self.L, self.Qc, self.H, self.Pinf = self.prior.kernel_to_state_space()
        self.state_dim = self.F.shape[0]
        H = self.prior.measurement_model(self.r[0])
        self.func_dim = H.shape[0]
        self.minf = np.zeros([self.state_dim, 1])
        self.sites = EP() if approx_inf is None else approx_inf
        if verbose:
            print('inference method is', self.sites.name)
        self._resident = {}   # id(host array) -> device tensor (training data stays in HBM)
        self._model_cache = {}

    # ------------------------------------------------------------------ plumbing
    def _dev(self, x):
        """Device copy of a host array.  Only the TRAINING arrays (self.y, self.dt) stay resident between calls; anything
        else (merged train/test grids of predict(), zeros of posterior_sample(), ...) is uploaded for the call only, so
        repeated predictions do not accumulate device memory."""
        if x is None or isinstance(x, torch.Tensor):
            return _as_dev(x, self.dtype, self.device)
        if not (x is self.y or x is self.dt):
            return _as_dev(x, self.dtype, self.device)
        key = (id(x), self.dtype)
        hit = self._resident.get(key)
        if hit is None or hit[0] is not x:
            hit = (x, _as_dev(x, self.dtype, self.device))
            self._resident[key] = hit
        return hit[1]

    def _params(self, params):
        if params is None:
            params = [self.prior.hyp.copy() if not isinstance(self.prior.hyp, list) else list(self.prior.hyp),
                      self.likelihood.hyp.copy()]
        return params

    def _model(self, params, r=None, N=None):
        """kjx_model_t for these hyperparameters (and, for a spatio-temporal prior, these spatial inputs).  Cached: a
        training loop calls run() with the same hyperparameters for the energy, the finite differences and the update,
        and the per-step measurement rows of a spatio-temporal prior cost a Cholesky solve with N right-hand sides."""
        uses_r = hasattr(self.prior, "spatial_weights") and not self.prior.fixed_grid      # only then H depends on r
        key = self._model_key(params, r if uses_r else None)
        hit = self._model_cache.get(key)
        if hit is not None and (not uses_r or np.array_equal(hit[1], r, equal_nan=True)):
            return hit[0]
        model = self._build_model(params, r)
        if len(self._model_cache) >= 32:
            self._model_cache.pop(next(iter(self._model_cache)))
        # (a private copy of r: the caller may overwrite its array in place between calls)
        self._model_cache[key] = (model, np.array(r, dtype=np.float64, copy=True) if uses_r else None)
        return model

    def _model_key(self, params, r):
        leaves = params[0] if isinstance(params[0], list) else [params[0]]
        raw = b"".join(np.ascontiguousarray(np.asarray(p, dtype=np.float64)).tobytes() for p in list(leaves) + [params[1]])
        sig = (float(self.sites.power), float(self.sites.damping), self.sites.kjx_code, self.likelihood.kjx_code,
               id(self.sites.cubature_func), id(self.prior), id(self.likelihood))
        rk = None if r is None else (np.shape(r), float(np.nansum(np.asarray(r, dtype=np.float64))))
        return (raw, sig, rk)

    def _build_model(self, params, r=None):
        theta_prior = _theta_prior(self.prior, params[0])
        theta_lik = softplus(params[1]) if np.size(params[1]) else None
        H_steps = None
        if hasattr(self.prior, "spatial_weights") and not self.prior.fixed_grid:
            # spatio-temporal prior with scattered data: H depends on r_n (priors.py:1148-1161).  All steps at once:
            # one factorisation of Kzz and one solve with N right-hand sides instead of N of each
            r_all = np.asarray(r, dtype=np.float64).reshape(len(r), -1)
            # rows without a spatial location (test steps of a different spatial layout, sde_gp.py:77: predict-only, their
            # measurement row is never used) get a placeholder so the batched solve stays finite
            r_all = np.where(np.isnan(r_all), 0.0, r_all)
            Kx = self.prior.spatial_weights(r_all[:, 0], theta_prior)                 # [N, M]
            bs = self.prior.state_dim // self.prior.M                                 # temporal states per inducing point
            Hs = np.zeros((Kx.shape[0], 1, self.prior.state_dim))
            Hs[:, 0, ::bs] = Kx                                                       # kron(Kx[n], [1, 0, ..])
            H_steps = torch.as_tensor(Hs, dtype=torch.float64, device=self.device).contiguous()
        return _Model(self.prior, self.likelihood, self.sites, theta_prior, theta_lik, H_steps=H_steps,
                      r0=None if r is None else np.asarray(r)[0])

    def update_model(self, theta_prior=None):
        self.F, self.L, self.Qc, self.H, self.Pinf = self.prior.kernel_to_state_space(hyperparams=theta_prior)

    # ------------------------------------------------------------------ sde_gp.py:230-309
    def kalman_filter(self, y, dt, params, store=False, mask=None, site_params=None, r=None, sequential=False):
        model = self._model(params, r)
        y_d, dt_d = self._dev(y), self._dev(dt)
        N, d, f = dt_d.shape[0], model.state_dim, model.func_dim
        mask_d = None
        if mask is not None:
            mask_np = np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask).reshape(N, -1)
            mask_d = torch.as_tensor(np.any(mask_np, axis=1).astype(np.uint8), device=self.device)   # sde_gp.py:297
        sm_in = sv_in = None
        if site_params is not None:
            sm_in, sv_in = self._dev(site_params[0]), self._dev(site_params[1])
        fm = fP = sm_out = sv_out = None
        if store:
            fm = torch.empty(N, d, 1, dtype=self.dtype, device=self.device)
            fP = torch.empty(N, d, d, dtype=self.dtype, device=self.device)
            if site_params is None:
                sm_out = torch.empty(N, f, 1, dtype=self.dtype, device=self.device)
                sv_out = torch.empty(N, f, f, dtype=self.dtype, device=self.device)
        nlml = torch.empty(1, dtype=self.dtype, device=self.device)
        ws, ws_bytes = _workspace(model, N, self.dtype, self.device)
        flags = _lib.FLAG_SEQUENTIAL if sequential else 0
        _lib.check(_fn("kjx_filter", self.dtype)(
            model.ref(), C.c_int64(N), _lib.ptr(dt_d), _lib.ptr(y_d), _lib.ptr(mask_d), _lib.ptr(sm_in), _lib.ptr(sv_in),
            _lib.ptr(fm), _lib.ptr(fP), _lib.ptr(sm_out), _lib.ptr(sv_out), _lib.ptr(nlml), _lib.ptr(ws),
            C.c_size_t(ws.numel()), C.c_uint32(flags), _stream(self.device)), "kjx_filter")
        if store:
            sites = (sm_in.reshape(N, f, 1), sv_in.reshape(N, f, f)) if site_params is not None else (sm_out, sv_out)
            return nlml[0], (fm, fP, sites)
        return nlml[0]

    # ------------------------------------------------------------------ sde_gp.py:311-384
    def rauch_tung_striebel_smoother(self, params, m_filtered, P_filtered, dt, store=False, return_full=False,
                                     y=None, site_params=None, r=None, sequential=False):
        model = self._model(params, r)
        dt_d = self._dev(dt)
        N, d, f = dt_d.shape[0], model.state_dim, model.func_dim
        fm, fP = self._dev(m_filtered), self._dev(P_filtered)
        y_d = self._dev(y)
        sm_in = sv_in = new_sm = new_sv = None
        if site_params is not None:
            if y is None:
                raise ValueError("the site update needs the observations: pass y together with site_params")
            sm_in, sv_in = self._dev(site_params[0]), self._dev(site_params[1])
            new_sm = torch.empty(N, f, 1, dtype=self.dtype, device=self.device)
            new_sv = torch.empty(N, f, f, dtype=self.dtype, device=self.device)
        s_mean = s_cov = None
        if store:
            k = d if return_full else f
            s_mean = torch.empty(N, k, 1, dtype=self.dtype, device=self.device)
            s_cov = torch.empty(N, k, k, dtype=self.dtype, device=self.device)
        ws, _ = _workspace(model, N, self.dtype, self.device)
        flags = (_lib.FLAG_RETURN_FULL if return_full else 0) | (_lib.FLAG_SEQUENTIAL if sequential else 0)
        _lib.check(_fn("kjx_smoother", self.dtype)(
            model.ref(), C.c_int64(N), _lib.ptr(dt_d), _lib.ptr(fm), _lib.ptr(fP), _lib.ptr(y_d), _lib.ptr(sm_in),
            _lib.ptr(sv_in), _lib.ptr(s_mean), _lib.ptr(s_cov), _lib.ptr(new_sm), _lib.ptr(new_sv), _lib.ptr(ws),
            C.c_size_t(ws.numel()), C.c_uint32(flags), _stream(self.device)), "kjx_smoother")
        if site_params is not None:
            site_params = (new_sm, new_sv)
        if store:
            return site_params, s_mean, s_cov
        return site_params

    # ------------------------------------------------------------------ gradients (finite differences)
    def _energy(self, params, site_params):
        """Filter energy for the finite differences: ALWAYS through the fp64 entry points — in float the energy carries
        ~1e-4 of rounding noise, which a difference over eps = 1e-5 would turn into a meaningless gradient."""
        keep = self.dtype
        self.dtype = torch.float64
        try:
            return float(self.kalman_filter(self.y, self.dt, params, False, None, site_params, self.r))
        finally:
            self.dtype = keep

    def _fd_grad(self, params, site_params, eps=1e-5):
        prior_p, lik_p = params
        is_list = isinstance(prior_p, list)
        leaves = [np.array(p, dtype=np.float64) for p in (prior_p if is_list else [prior_p])] + [np.array(lik_p, dtype=np.float64)]
        grads = [np.zeros_like(l) for l in leaves]
        for li, leaf in enumerate(leaves):
            for j in range(leaf.size):
                vals = []
                for sgn in (1.0, -1.0):
                    pert = [l.copy() for l in leaves]
                    pert[li].reshape(-1)[j] += sgn * eps
                    pp = pert[:-1] if is_list else pert[0]
                    vals.append(self._energy([pp, pert[-1]], site_params))
                grads[li].reshape(-1)[j] = (vals[0] - vals[1]) / (2 * eps)
        return [grads[:-1] if is_list else grads[0], grads[-1]]

    # ------------------------------------------------------------------ sde_gp.py:163-188
    def run(self, params=None, compute_grad=True, store_posterior=False):
        """One training iteration: filter (energy) then smoother + site update, fused in one kjx_run call."""
        params = self._params(params)
        sites_in = self.sites.site_params
        dlZ = self._fd_grad(params, sites_in) if compute_grad else None
        model = self._model(params, self.r)
        y_d, dt_d = self._dev(self.y), self._dev(self.dt)
        N, d, f = dt_d.shape[0], model.state_dim, model.func_dim
        sm_in = sv_in = None
        if sites_in is not None:
            sm_in, sv_in = self._dev(sites_in[0]), self._dev(sites_in[1])
        # run() only uses the filtered moments internally (sde_gp.py:183-187): the fused scan keeps them packed
        # in its workspace.  Dense buffers are needed only by the one-chain filter that initialises the sites of
        # a non-Gaussian model on the first iteration.
        fm = fP = None
        if f > 1 or (sites_in is None and not (self.likelihood.kjx_code == _lib.LIK_GAUSSIAN and self.sites.kjx_code == _lib.INF_EP)):
            fm = torch.empty(N, d, 1, dtype=self.dtype, device=self.device)
            fP = torch.empty(N, d, d, dtype=self.dtype, device=self.device)
        new_sm = torch.empty(N, f, 1, dtype=self.dtype, device=self.device)
        new_sv = torch.empty(N, f, f, dtype=self.dtype, device=self.device)
        pm = pv = None
        if store_posterior:
            pm = torch.empty(N, f, 1, dtype=self.dtype, device=self.device)
            pv = torch.empty(N, f, f, dtype=self.dtype, device=self.device)
        nlml = torch.empty(1, dtype=self.dtype, device=self.device)
        ws, _ = _workspace(model, N, self.dtype, self.device)
        _lib.check(_fn("kjx_run", self.dtype)(
            model.ref(), C.c_int64(N), _lib.ptr(dt_d), _lib.ptr(y_d), _lib.ptr(sm_in), _lib.ptr(sv_in), _lib.ptr(fm),
            _lib.ptr(fP), _lib.ptr(new_sm), _lib.ptr(new_sv), _lib.ptr(pm), _lib.ptr(pv), _lib.ptr(nlml), _lib.ptr(ws),
            C.c_size_t(ws.numel()), C.c_uint32(0), _stream(self.device)), "kjx_run")
        self.sites.site_params = (new_sm, new_sv)
        if store_posterior:
            self.posterior = (pm, pv)
        return float(nlml[0]), dlZ

    # ------------------------------------------------------------------ sde_gp.py:190-219
    def run_two_stage(self, params=None, compute_grad=True):
        params = self._params(params)
        _, (fm, fP, self.sites.site_params) = self.kalman_filter(self.y, self.dt, params, True, None,
                                                                 self.sites.site_params, self.r)
        self.sites.site_params = self.rauch_tung_striebel_smoother(params, fm, fP, self.dt, False, False, self.y,
                                                                   self.sites.site_params, self.r)
        nlml = self._energy(params, self.sites.site_params)
        dlZ = self._fd_grad(params, self.sites.site_params) if compute_grad else None
        return nlml, dlZ

    # ------------------------------------------------------------------ sde_gp.py:145-161
    def neg_log_marg_lik(self, params=None, compute_grad=True):
        params = self._params(params)
        nlml = self._energy(params, self.sites.site_params)
        dlZ = self._fd_grad(params, self.sites.site_params) if compute_grad else None
        return nlml, dlZ

    # ------------------------------------------------------------------ sde_gp.py:85-103
    def predict_everywhere(self, y, r, dt, train_id, mask, site_params=None, sampling=False, return_full=False):
        params = self._params(None)
        site_params = self.sites.site_params if site_params is None else site_params
        if site_params is not None and not sampling:
            Nall = dt.shape[0]
            site_mean = torch.zeros(Nall, self.func_dim, 1, dtype=self.dtype, device=self.device)
            site_cov = 1e5 * torch.eye(self.func_dim, dtype=self.dtype, device=self.device).repeat(Nall, 1, 1)   # :95
            idx = torch.as_tensor(train_id, device=self.device)
            site_mean[idx] += self._dev(site_params[0]).reshape(-1, self.func_dim, 1)
            site_cov[idx] = self._dev(site_params[1]).reshape(-1, self.func_dim, self.func_dim)
            site_params = (site_mean, site_cov)
        if site_params is not None and not return_full and self.dtype == torch.float64:
            # filter (masked steps predict-only) + smoother marginals as ONE fused scan (kjx_posterior)
            model = self._model(params, r)
            y_d, dt_d = self._dev(y), self._dev(dt)
            N = dt_d.shape[0]
            mask_d = None
            if mask is not None:
                mask_d = torch.as_tensor(np.any(np.asarray(mask).reshape(N, -1), axis=1).astype(np.uint8), device=self.device)
            sm_in, sv_in = self._dev(site_params[0]), self._dev(site_params[1])
            pm = torch.empty(N, self.func_dim, 1, dtype=self.dtype, device=self.device)
            pc = torch.empty(N, self.func_dim, self.func_dim, dtype=self.dtype, device=self.device)
            ws, _ = _workspace(model, N, self.dtype, self.device)
            rc = _lib.lib().kjx_posterior_f64(model.ref(), C.c_int64(N), _lib.ptr(dt_d), _lib.ptr(y_d), _lib.ptr(mask_d),
                                              _lib.ptr(sm_in), _lib.ptr(sv_in), _lib.ptr(pm), _lib.ptr(pc), None, _lib.ptr(ws),
                                              C.c_size_t(ws.numel()), _stream(self.device))
            if rc == 0:
                return pm, pc, (sm_in.reshape(N, self.func_dim, 1), sv_in.reshape(N, self.func_dim, self.func_dim))
            if rc != -2:                       # anything but "unsupported": a real error
                _lib.check(rc, "kjx_posterior")
        _, (fm, fP, site_params) = self.kalman_filter(y, dt, params, True, mask, site_params, r)
        _, pm, pc = self.rauch_tung_striebel_smoother(params, fm, fP, dt, True, return_full, None, None, r)
        return pm, pc, site_params

    # ------------------------------------------------------------------ sde_gp.py:52-83
    def predict(self, t=None, r=None, site_params=None, sampling=False):
        if t is None:
            t, y, r = self.t, self.y, self.r
        (t, y, r, r_test, dt, train_id, test_id, mask) = test_input_admin(self.t, self.y, self.r, t, None, r)
        return_full = r_test.shape[1] != r.shape[1]   # spatial test locations of a different size than the training ones
        y = np.where(np.isnan(y), 0.0, y)   # masked entries are never read for their value (sde_gp.py:286)
        pm, pc, site_params = self.predict_everywhere(y, r, dt, train_id, mask, site_params, sampling, return_full)
        idx = torch.as_tensor(test_id, device=self.device)
        pm, pc = pm[idx], pc[idx]
        if return_full:
            pm, pc = self.compute_measurement(r_test, pm, pc)                        # sde_gp.py:77-82
        return torch.squeeze(pm), torch.squeeze(pc)

    # ------------------------------------------------------------------ sde_gp.py:105-108 (vmapped over the test points)
    def compute_measurement(self, r, mean, cov, hyp_prior=None):
        """H(r_n) m_n and H(r_n) P_n H(r_n)^T for every test point: the full-state posterior [M,d,1], [M,d,d] measured at
        that point's own spatial locations r[M, K] (K function values per test time) -> [M,K,1], [M,K,K]."""
        theta = _theta_prior(self.prior, self.prior.hyp) if hyp_prior is None else hyp_prior
        r = np.asarray(r, dtype=np.float64).reshape(len(r), -1)
        H = np.stack([self.prior.measurement_model(r[n], theta) for n in range(r.shape[0])])     # [M, K, d]
        H = torch.as_tensor(H, dtype=mean.dtype, device=mean.device)
        return H @ mean, H @ cov @ H.transpose(1, 2)

    # ------------------------------------------------------------------ sde_gp.py:386-417
    def prior_sample(self, num_samps, t=None, seed=0):
        """num_samps draws of f ~ N(0, K) at the (sorted) inputs t, [N, func_dim, num_samps] like the reference.
        kjx_prior_sample: the per-step noise factorisations run in parallel over time, one warp walks each sample."""
        params = self._params(None)
        if t is None:
            t = self.t
        else:
            t = np.asarray(t, dtype=np.float64)
            t = t.reshape(t.shape[0], -1)
            t = t[np.argsort(t[:, 0])]                                              # :398-399
        dt = np.concatenate([np.array([0.0]), np.diff(t[:, 0])])                    # :400
        N = dt.shape[0]
        model = self._model(params, None)
        dt_d = self._dev(dt)
        out = torch.empty(N, self.func_dim, int(num_samps), dtype=torch.float64, device=self.device)
        nbytes = C.c_size_t(0)
        _lib.check(_lib.lib().kjx_prior_sample_workspace_bytes(model.ref(), C.c_int64(N), C.c_int32(int(num_samps)), C.byref(nbytes)),
                   "kjx_prior_sample_workspace_bytes")
        ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().kjx_prior_sample_f64(model.ref(), C.c_int64(N), _lib.ptr(dt_d), C.c_int32(int(num_samps)),
                                                   C.c_uint64(int(seed)), _lib.ptr(out), _lib.ptr(ws), C.c_size_t(ws.numel()),
                                                   _stream(self.device)), "kjx_prior_sample")
        return out

    # ------------------------------------------------------------------ sde_gp.py:419-450
    def posterior_sample(self, num_samps, t=None, r=None, seed=0, batched=True):
        """Posterior draws at the test inputs by smoothing prior draws through the Gaussian sites (Doucet's conditional
        simulation, as the reference does it): prior_samp - smooth(prior_samp + site noise) + post_mean."""
        if t is None:
            t, r = self.t, self.r
        (t_all, y_all, r_all, r_test, dt_all, train_id, test_id, mask) = test_input_admin(self.t, self.y, self.r, t, None, r)
        y_all = np.where(np.isnan(y_all), 0.0, y_all)
        post_mean, _, (site_mean, site_cov) = self.predict_everywhere(y_all, r_all, dt_all, train_id, mask)      # :438
        prior_samp = self.prior_sample(num_samps, t=t_all, seed=seed)                                            # :439
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(seed) + 123)
        noise = torch.randn(prior_samp.shape, generator=gen, dtype=torch.float64, device=self.device)
        prior_samp_y = prior_samp + torch.sqrt(site_cov.reshape(-1, 1, 1)) * noise                               # utils.py:247-250
        smoothed = self._smooth_draws(prior_samp_y, site_cov, y_all, r_all, dt_all, train_id, mask) if batched else None
        if smoothed is None:
            smoothed = torch.empty_like(prior_samp_y)
            zeros = np.zeros_like(y_all)
            for i in range(int(num_samps)):                                                                     # :443-449
                sm_i, _, _ = self.predict_everywhere(zeros, r_all, dt_all, train_id, mask,
                                                     (prior_samp_y[..., i].reshape(-1, 1, 1).contiguous(), site_cov), sampling=True)
                smoothed[..., i] = sm_i[..., 0]
        idx = torch.as_tensor(test_id, device=self.device)
        return torch.squeeze(prior_samp - smoothed + post_mean)[idx]                                             # :450

    def _smooth_draws(self, draws, site_cov, y_all, r_all, dt_all, train_id, mask):
        """The reference's loop over samples (sde_gp.py:443-449) as ONE filter + smoother call: the S pseudo-observation
        series are laid end to end with a gap before each that is long enough for exp(-gap * lam) to underflow to zero for
        every block of the prior, so the transition into a series is exactly A = 0, Q = Pinf — the prior, i.e. what the
        first step of a separate call sees — and the smoother gain out of the previous series is exactly zero.  S * N steps
        go through the parallel-in-time scan at once instead of S launch sequences of N steps.
        Returns None (caller loops) when a block does not decay (Cosine / Periodic) or the model is spatial / multi-latent."""
        S, N = int(draws.shape[-1]), int(draws.shape[0])
        if S < 2 or self.func_dim != 1 or self.dtype != torch.float64 or hasattr(self.prior, "spatial_weights"):
            return None
        lam = [b[2] for b in self.prior.blocks(_theta_prior(self.prior, self._params(None)[0]))]
        if min(lam) <= 0.0 or not np.isfinite(np.max(dt_all)):
            return None
        gap = max(800.0 / min(lam), 1e3 * float(np.max(dt_all)))           # exp(-800) == 0 in fp64
        dt_cat = np.tile(np.asarray(dt_all, dtype=np.float64).reshape(-1), S)
        dt_cat[N::N] = gap
        mask_cat = None if mask is None else np.tile(np.asarray(mask).reshape(N, -1), (S, 1))
        sm_cat = draws.reshape(N, S).t().contiguous().reshape(S * N, 1, 1)
        sv_cat = site_cov.reshape(1, N).expand(S, N).contiguous().reshape(S * N, 1, 1)
        pm, _, _ = self.predict_everywhere(np.zeros((S * N, 1)), r_all, dt_cat, train_id, mask_cat, (sm_cat, sv_cat), sampling=True)
        return pm.reshape(S, N).t().reshape(N, 1, S)

    # ------------------------------------------------------------------ sde_gp.py:110-143
    def negative_log_predictive_density(self, t=None, y=None, r=None):
        if t is None:
            t, y, r = self.t, self.y, self.r
        (t, y_all, r, r_test, dt, train_id, test_id, mask) = test_input_admin(self.t, self.y, self.r, t, y, r)
        return_full = r_test.shape[1] != r.shape[1]
        y_filter = np.where(np.isnan(y_all), 0.0, y_all)
        pm, pc, _ = self.predict_everywhere(y_filter, r, dt, train_id, mask, sampling=False, return_full=return_full)
        idx = torch.as_tensor(test_id, device=self.device)
        pm, pc = pm[idx], pc[idx]
        if return_full:
            pm, pc = self.compute_measurement(r_test, pm, pc)                        # sde_gp.py:133-137
            if pm.shape[1] != 1:
                raise NotImplementedError("NLPD of several function values per test time needs a multi-dimensional "
                                          "moment match (func_dim > 1)")
        hyp_lik = softplus(self.likelihood.hyp) if np.size(self.likelihood.hyp) else None
        # vmapped moment match with power 1 and the default 20-point Gauss-Hermite rule (sde_gp.py:139-142)
        lpd, _, _ = self.likelihood.moment_match(y_all[test_id], pm, pc, hyp_lik, 1, None)
        return float(-torch.mean(lpd))
