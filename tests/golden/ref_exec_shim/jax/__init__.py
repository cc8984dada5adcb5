"""TEST INFRASTRUCTURE: a minimal stand-in for the parts of the `jax` API the reference's solver files use,
backed by torch (FP64, CPU) - so that the UNMODIFIED reference sources under /root/reference/code can be imported
and executed in a container where JAX cannot be installed (tests/golden/make_ref_exec_golden.py).  Semantics kept:
`jnp.abs` differentiates as select(x >= 0, g, -g) (JAX's rule: +1 at 0), everything is float64."""
import torch

from . import numpy  # noqa: F401
from . import random  # noqa: F401
from .config import config  # noqa: F401
from .numpy import _t

torch.set_default_dtype(torch.float64)


def jit(fun=None, static_argnums=None, **_):
    return fun


def _tree_map(f, t):
    if isinstance(t, dict):
        return {k: _tree_map(f, v) for k, v in t.items()}
    return f(t)


def vmap(fun, in_axes=0, out_axes=0):
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        conv = [_tree_map(_t, a) for a in args]
        return torch.func.vmap(fun, in_dims=tuple(axes), out_dims=out_axes)(*conv)
    return wrapped


def grad(fun, argnums=0):
    def wrapped(*args):
        conv = [_tree_map(_t, a) for a in args]
        return torch.func.grad(fun, argnums=argnums)(*conv)
    return wrapped


def value_and_grad(fun, argnums=0):
    """Reverse mode over a params pytree (dict of dicts of arrays), by torch.autograd."""
    def wrapped(*args):
        args = list(args)
        leaves = []

        def req(x):
            x = _t(x).detach().clone().requires_grad_(True)
            leaves.append(x)
            return x
        args[argnums] = _tree_map(req, args[argnums])
        val = fun(*args)
        grads = torch.autograd.grad(val, leaves, allow_unused=True)
        it = iter(torch.zeros_like(l) if g is None else g for l, g in zip(leaves, grads))
        return val.detach(), _tree_map(lambda _: next(it), args[argnums])
    return wrapped
