"""Where does the d = 192 (C4 shape) path lose digits?  Filtered and smoothed moments of two iterations against the
oracle, run with KJX_GEN_CHUNKS=1 (plain sequential walk) and with the default chunking.  Diagnostic, not a test."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import kalman_jax_b200 as kjx  # noqa: E402


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float).reshape(-1), np.asarray(b, dtype=float).reshape(-1)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def main():
    ell = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
    rng = np.random.default_rng(99)
    N = int(os.environ.get("DIAG_N", "72"))
    t = np.sort(rng.uniform(0, 6, N)); r = rng.uniform(-3, 3, N); y = (rng.uniform(size=N) < 0.5).astype(float)
    z = np.linspace(-3, 3, 64)
    mk = lambda mod, inf: mod.SDEGP(mod.priors.SpatioTemporalMatern52(1.0, 1.0, ell, z=z), mod.likelihoods.Probit(), t, y, r=r,  # noqa: E731
                                    approx_inf=inf.VI())
    model, ref = mk(kjx, kjx.approximate_inference), mk(oracle, oracle.approx_inf)
    cache = "/tmp/diag_d192_%g_%d.npz" % (ell, N)
    have = os.path.exists(cache)
    store = dict(np.load(cache)) if have else {}
    for it in range(2):
        params = model._params(None)
        sites_before = model.sites.site_params
        if sites_before is not None:
            _, (fm, fP, _) = model.kalman_filter(model.y, model.dt, params, True, None, sites_before, model.r)
        nlml, _ = model.run(compute_grad=False, store_posterior=True)
        if not have:
            want_nlml, (wfm, wfP, wsites) = ref.kalman_filter(ref.y, ref.dt, True, None, ref.sites.site_params, ref.r)
            w_new, wsm, wsc = ref.rauch_tung_striebel_smoother(wfm, wfP, ref.dt, True, False, ref.y, wsites, ref.r)
            ref.sites.site_params = w_new
            store.update({"nlml%d" % it: want_nlml, "fm%d" % it: wfm, "fP%d" % it: wfP, "sm%d" % it: wsm, "sc%d" % it: wsc,
                          "nm%d" % it: w_new[0], "nv%d" % it: w_new[1]})
        line = {"ell": ell, "chunks": os.environ.get("KJX_GEN_CHUNKS", "default"), "it": it,
                "nlml": abs(nlml - float(store["nlml%d" % it])) / abs(float(store["nlml%d" % it])),
                "smooth_mean": rel_err(_np(model.posterior[0]), store["sm%d" % it]),
                "smooth_var": rel_err(_np(model.posterior[1]), store["sc%d" % it]),
                "site_mean": rel_err(_np(model.sites.site_params[0]), store["nm%d" % it]),
                "site_var": rel_err(_np(model.sites.site_params[1]), store["nv%d" % it])}
        if sites_before is not None:
            line["filt_mean"] = rel_err(_np(fm), store["fm%d" % it]); line["filt_cov"] = rel_err(_np(fP), store["fP%d" % it])
        print({k: (float("%.2e" % v) if isinstance(v, float) else v) for k, v in line.items()}, flush=True)
    if not have:
        np.savez(cache, **store)


if __name__ == "__main__":
    main()
