"""ctypes binding of libkjx.so (the C ABI declared in include/kjx.h).

There is no CPU fallback: if the CUDA library is missing every compute entry point raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkjx.so")

KJX_MAX_BLOCKS = 16
KJX_MAX_CUBATURE = 32

# enums (include/kjx.h)
BLOCK_MATERN32, BLOCK_MATERN52, BLOCK_QP_MATERN32, BLOCK_MATERN12, BLOCK_ROTATION, BLOCK_MATERN72 = 1, 2, 3, 4, 5, 6
BLOCK_SUBBAND_MATERN52 = 7
LIK_GAUSSIAN, LIK_BERNOULLI_PROBIT, LIK_BERNOULLI_LOGIT, LIK_POISSON_EXP, LIK_POISSON_LOGISTIC = 1, 2, 3, 4, 5
LIK_HETERO_SOFTPLUS, LIK_HETERO_EXP = 6, 7
INF_EP, INF_EEP, INF_SLEP, INF_VI, INF_MOMENT_MATCH = 1, 2, 3, 4, 5
FLAG_SEQUENTIAL, FLAG_RETURN_FULL, FLAG_NO_SITE_UPDATE, FLAG_RESIDENT_SITES, FLAG_RESIDENT_INIT = 1, 2, 4, 8, 16
OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4


class KjxBlock(C.Structure):
    _fields_ = [("kind", C.c_int32), ("count", C.c_int32), ("lam", C.c_double), ("omega", C.c_double)]


class KjxModel(C.Structure):
    _fields_ = [
        ("state_dim", C.c_int32), ("func_dim", C.c_int32), ("n_blocks", C.c_int32), ("reserved0", C.c_int32),
        ("blocks", KjxBlock * KJX_MAX_BLOCKS),
        ("Pinf", C.c_void_p), ("H", C.c_void_p), ("H_steps", C.c_void_p),
        ("likelihood", C.c_int32), ("method", C.c_int32),
        ("lik_hyp", C.c_double), ("power", C.c_double), ("damping", C.c_double),
        ("n_cub", C.c_int32), ("cub_rule", C.c_int32),
        ("cub_x", C.c_double * KJX_MAX_CUBATURE), ("cub_w", C.c_double * KJX_MAX_CUBATURE),
    ]


class KjxError(RuntimeError):
    pass


_lib = None

# every symbol include/kjx.h declares (tests check the shared object exports each of them)
SYMBOLS = [
    "kjx_version", "kjx_status_string", "kjx_workspace_bytes",
    "kjx_filter_f64", "kjx_smoother_f64", "kjx_run_f64", "kjx_posterior_f64", "kjx_site_update_f64", "kjx_discretise_f64",
    "kjx_filter_f32", "kjx_smoother_f32", "kjx_run_f32", "kjx_site_update_f32", "kjx_discretise_f32",
    "kjx_filter_summary_doubles",
    "kjx_filter_summary_f64", "kjx_filter_apply_f64", "kjx_smoother_apply_f64",
    "kjx_fold_shard_summaries_host", "kjx_fold_shard_summaries_f64", "kjx_exchange_fold_f64", "kjx_run_host_f64", "kjx_run_host_scratch_bytes",
    "kjx_prior_sample_workspace_bytes", "kjx_prior_sample_f64",
    "kjx_shard_link_doubles", "kjx_sharded_iteration_f64", "kjx_run_host_submit_f64",
]


def lib():
    """Load libkjx.so (built in-tree by __graft_entry__.build()).  Raises if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KjxError("libkjx.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.kjx_status_string.restype = C.c_char_p
        L.kjx_filter_summary_doubles.restype = C.c_size_t
        L.kjx_shard_link_doubles.restype = C.c_size_t
        for name in SYMBOLS:
            getattr(L, name)  # AttributeError if the symbol is missing
        _lib = L
    return _lib


def check(status, what):
    if status != 0:
        raise KjxError("%s failed: %s (%d)" % (what, lib().kjx_status_string(status).decode(), status))


def ptr(t):
    """device (or host) address of a torch tensor / numpy array, or NULL."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())
