"""NumPy-backed stand-in for the jax==0.1.67 surface used by /root/reference (see ../README.md)."""
import sys
import functools
import numpy as _np

from . import config as _config_mod  # noqa: F401

# `import jax.numpy as np`  ->  plain numpy (fp64 by default, like jax_enable_x64=True)
numpy = _np
sys.modules[__name__ + ".numpy"] = _np

partial = functools.partial


def jit(fun=None, static_argnums=None, **_):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    """Python-loop vmap: slice each mapped arg along its axis, call, stack outputs along out_axes."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(a)[ax]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else _np.take(_np.asarray(a), i, axis=ax) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if isinstance(outs[0], tuple):
            k = len(outs[0])
            oax = out_axes if isinstance(out_axes, (tuple, list)) else (out_axes,) * k
            return tuple(_np.stack([_np.asarray(o[j]) for o in outs], axis=oax[j]) for j in range(k))
        return _np.stack([_np.asarray(o) for o in outs], axis=out_axes)
    return mapped


def _tree_flatten(p):
    if isinstance(p, (list, tuple)):
        leaves, defs = [], []
        for q in p:
            l, d = _tree_flatten(q)
            leaves += l
            defs.append(d)
        return leaves, (type(p), defs)
    return [_np.array(p, dtype=float)], None


def _tree_unflatten(leaves, d):
    if d is None:
        return leaves.pop(0)
    typ, defs = d
    return typ([_tree_unflatten(leaves, q) for q in defs]) if typ is list else tuple(
        _tree_unflatten(leaves, q) for q in defs)


def value_and_grad(fun, argnums=0, has_aux=False, _eps=1e-6):
    """Central finite differences (the shim has no autodiff); gradients are informational only."""
    def wrapped(*args):
        out = fun(*args)
        val = out[0] if has_aux else out
        leaves, treedef = _tree_flatten(args[argnums])
        grads = [_np.zeros_like(l) for l in leaves]
        for li, leaf in enumerate(leaves):
            flat = leaf.reshape(-1)
            for j in range(flat.size):
                vals = []
                for sgn in (+1.0, -1.0):
                    pert = [l.copy() for l in leaves]
                    pert[li].reshape(-1)[j] += sgn * _eps
                    a2 = list(args)
                    a2[argnums] = _tree_unflatten(pert, treedef)
                    o2 = fun(*a2)
                    vals.append(float(o2[0] if has_aux else o2))
                grads[li].reshape(-1)[j] = (vals[0] - vals[1]) / (2 * _eps)
        g = _tree_unflatten(grads, treedef)
        return out, g
    return wrapped


def grad(fun, argnums=0):
    def wrapped(*args):
        x = _np.asarray(args[argnums], dtype=float)
        h = 1e-6 * max(1.0, float(_np.max(_np.abs(x))))
        a_p, a_m = list(args), list(args)
        a_p[argnums], a_m[argnums] = x + h, x - h
        return (fun(*a_p) - fun(*a_m)) / (2 * h)
    return wrapped


def jacrev(fun, argnums=0):
    raise NotImplementedError("jax_shim: jacrev is not emulated (base-class analytical_linearisation is out of scope)")


from . import ops, random, nn  # noqa: E402,F401
from . import scipy  # noqa: E402,F401
from . import experimental  # noqa: E402,F401
