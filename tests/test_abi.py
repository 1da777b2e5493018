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


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kalman_jax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
