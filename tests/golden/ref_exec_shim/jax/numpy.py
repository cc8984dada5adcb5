"""jax.numpy names used by the reference (kernel_matrix.py, model_GP_solver_*.py), on torch float64."""
import math

import numpy as _np
import torch

pi = math.pi


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(_np.asarray(x, dtype=_np.float64))


def abs(x):                                  # noqa: A001  - JAX: jvp = select(x >= 0, g, -g)
    x = _t(x)
    return torch.where(x >= 0, x, -x)


def exp(x): return torch.exp(_t(x))
def cos(x): return torch.cos(_t(x))
def sin(x): return torch.sin(_t(x))
def sqrt(x): return torch.sqrt(_t(x))
def log(x): return torch.log(_t(x))
def square(x): return torch.square(_t(x))
def sum(x, axis=None): return torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis)   # noqa: A001
def matmul(a, b): return torch.matmul(_t(a), _t(b))
def eye(n): return torch.eye(int(n), dtype=torch.float64)
def hstack(xs): return torch.cat([_t(x).reshape(-1) for x in xs])
def array(x): return _t(x)
def zeros(shape): return torch.zeros(shape, dtype=torch.float64)
def ones(shape): return torch.ones(shape, dtype=torch.float64)
def linspace(a, b, num): return torch.linspace(a, b, num, dtype=torch.float64)


class linalg(object):
    @staticmethod
    def solve(A, B): return torch.linalg.solve(_t(A), _t(B))

    @staticmethod
    def slogdet(A):
        s, l = torch.linalg.slogdet(_t(A))
        return s, l

    @staticmethod
    def norm(x): return torch.linalg.norm(_t(x))


def tile(x, reps):
    # grids stay numpy (the reference reads `.size` of the tiled pair grids as an int attribute, as on jax arrays)
    if not isinstance(x, torch.Tensor):
        return _np.tile(_np.asarray(x, dtype=_np.float64), reps)
    return x.repeat(*reps) if isinstance(reps, (tuple, list)) else x.repeat(reps)


def transpose(x):
    return x.T if isinstance(x, torch.Tensor) else _np.transpose(x)
