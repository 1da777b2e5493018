"""The C-ABI shared library loads and exports every symbol include/kjx.h declares (no compute, no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kjx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kjx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "kalman_jax_b200", "libkjx.so"))
    names = _declared()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.kjx_version() >= 100
    lib.kjx_status_string.restype = ctypes.c_char_p
    assert lib.kjx_status_string(0) == b"ok"
    assert b"workspace" in lib.kjx_status_string(-3)


def test_python_binding_lists_the_same_symbols():
    from kalman_jax_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared()


def test_api_misuse_returns_status_not_crash():
    from kalman_jax_b200 import _lib
    lib = _lib.lib()
    nbytes = ctypes.c_size_t(0)
    assert lib.kjx_workspace_bytes(None, ctypes.c_int64(10), 8, ctypes.byref(nbytes)) == -1
    m = _lib.KjxModel()            # zeroed model: invalid
    assert lib.kjx_workspace_bytes(ctypes.byref(m), ctypes.c_int64(10), 8, ctypes.byref(nbytes)) == -1


def test_workspace_size_is_monotone_in_N():
    import numpy as np
    from kalman_jax_b200 import _lib, priors, likelihoods, approximate_inference as ai
    from kalman_jax_b200.sde_gp import _Model
    model = _Model(priors.Matern52(1.0, 5.0), likelihoods.Gaussian(0.5), ai.EP(power=0.5))
    last = 0
    for N in (0, 1, 100, 10_000, 1 << 20, 1 << 24):
        nbytes = ctypes.c_size_t(0)
        assert _lib.lib().kjx_workspace_bytes(model.ref(), ctypes.c_int64(N), 8, ctypes.byref(nbytes)) == 0
        assert nbytes.value >= last
        last = nbytes.value
    assert last < 90 * (1 << 24)    # scan summaries (a few bytes/step) + packed filtered moments (72 B/step at d=3)


def test_every_prior_gives_a_valid_model_and_workspace():
    """Host logic only: each prior's block description adds up to its state dimension (kjx_workspace_bytes validates
    the model), the generic path's workspace grows with N, and the sampler refuses what it cannot do."""
    import numpy as np
    from kalman_jax_b200 import _lib, priors as P, likelihoods, approximate_inference as ai
    from kalman_jax_b200.sde_gp import _Model
    zoo = [P.Matern12(1.0, 2.0), P.Matern32(1.0, 2.0), P.Matern52(1.0, 2.0), P.Matern72(1.0, 2.0), P.Cosine([0.5]),
           P.Periodic(1.0, 1.0, 3.0), P.QuasiPeriodicMatern12(1.0, 1.0, 3.0, 9.0), P.QuasiPeriodicMatern32(1.0, 1.0, 3.0, 9.0),
           P.SubbandMatern12(1.0, 2.0, 0.7), P.SubbandMatern32(1.0, 2.0, 0.7), P.SubbandMatern52(1.0, 2.0, 0.7),
           P.Sum([P.SubbandMatern52(1.0, 2.0, 0.7), P.Matern32(1.0, 2.0)]),
           P.Sum([P.Matern52(1.0, 2.0), P.QuasiPeriodicMatern32(1.0, 1.0, 3.0, 9.0), P.Matern12(0.5, 1.0)]),
           P.SpatioTemporalMatern52(1.0, 1.0, 0.5), P.SpatialMatern52(1.0, 0.7), P.SpatialMatern32(1.0, 0.7)]
    lib = _lib.lib()
    for prior in zoo:
        model = _Model(prior, likelihoods.Poisson(), ai.EP(power=0.5), r0=np.zeros(1))
        assert model.state_dim == sum(_lib_block_size(k) * c for k, c, _, _ in prior.blocks())
        last = 0
        for N in (1, 100, 5000):
            nbytes = ctypes.c_size_t(0)
            assert lib.kjx_workspace_bytes(model.ref(), ctypes.c_int64(N), 8, ctypes.byref(nbytes)) == 0, prior.name
            assert nbytes.value >= last
            last = nbytes.value
        nb = ctypes.c_size_t(0)
        assert lib.kjx_prior_sample_workspace_bytes(model.ref(), ctypes.c_int64(100), ctypes.c_int32(8), ctypes.byref(nb)) == 0
        assert nb.value >= 101 * model.state_dim * 8 * 8
    # sampling: spatio-temporal priors (no constant measurement row) are unsupported, like in the reference (sde_gp.py:432)
    st = _Model(P.SpatioTemporalMatern52(1.0, 1.0, 0.5), likelihoods.Probit(), ai.EP(),
                H_steps=None, r0=np.zeros(1))
    st.c.H = None
    rc = lib.kjx_prior_sample_f64(st.ref(), ctypes.c_int64(10), None, ctypes.c_int32(0), ctypes.c_uint64(0), None, None,
                                  ctypes.c_size_t(0), None)
    assert rc == _lib.ERR_UNSUPPORTED
    # bad block description: dimension mismatch
    bad = _Model(P.Matern52(1.0, 2.0), likelihoods.Gaussian(0.5), ai.EP())
    bad.c.state_dim = 4
    assert lib.kjx_workspace_bytes(bad.ref(), ctypes.c_int64(10), 8, ctypes.byref(ctypes.c_size_t(0))) == _lib.ERR_INVALID_ARG


def _lib_block_size(kind):
    return {1: 2, 2: 3, 3: 4, 4: 1, 5: 2, 6: 4, 7: 6}[kind]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kalman_jax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
