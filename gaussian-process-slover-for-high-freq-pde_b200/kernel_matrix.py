"""Kernel builders - same class names and call signatures as the reference's kernel_matrix.py,
evaluated by libgphm's fused Gram kernels on the GPU (FP64, as the reference forces with
jax_enable_x64, kernel_matrix.py:6-7).

  Kernel_matrix(jitter, K_u).get_kernel_matrix(X1, X2, paras)   kernel_matrix.py:12-30
  Kernel_1d.kappa / D_x1_kappa / DD_x1_kappa                     kernel_matrix.py:45-57
  SE_Cos_1d, Matern52_Cos_1d, Matern52_1d, SE_1d                 kernel_matrix.py:107-193

The reference's `kappa` is a scalar function that is only ever called under `vmap` over flattened
pair lists; here `kappa(x1, y1, paras)` accepts scalars or equally shaped arrays and evaluates
every pair in one launch (the vmapped call shape).  `paras` is the reference dict
{'log-w','log-ls','freq'} of length-Q vectors.  Results are torch CUDA tensors (float64).
"""
import math

import torch

from . import _lib

DT = torch.float64


def _dev():
    if not torch.cuda.is_available():
        raise _lib.GphmError("gphm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_dev(x):
    """torch CUDA float64 contiguous view/copy of numpy / list / torch input."""
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=_dev(), dtype=DT).contiguous()
    import numpy as np
    return torch.as_tensor(np.asarray(x, dtype=np.float64)).to(_dev()).contiguous()


def pack_theta(paras):
    """{'log-w','log-ls','freq'} -> contiguous [log-w | log-ls | freq] device vector."""
    lw, ls, f = as_dev(paras["log-w"]).reshape(-1), as_dev(paras["log-ls"]).reshape(-1), as_dev(paras["freq"]).reshape(-1)
    if not (lw.numel() == ls.numel() == f.numel()):
        raise ValueError("kernel parameter vectors must have equal length")
    return torch.cat((lw, ls, f)), lw.numel()


class Kernel_1d(object):
    """Base kernel class (kernel_matrix.py:36-104).  Subclasses set `kernel_id`."""
    kernel_id = None

    def __init__(self, fix_dict=None, fix_paras=None):
        self.fix_dict = fix_dict
        self.fix_paras = fix_paras

    def _pairs(self, x1, y1, paras, order):
        if self.kernel_id is None:
            raise NotImplementedError                          # kernel_matrix.py:45-47
        lib = _lib.load()
        a, b = as_dev(x1), as_dev(y1)
        a, b = torch.broadcast_tensors(a, b)
        a, b = a.contiguous(), b.contiguous()
        theta, Q = pack_theta(paras)
        out = torch.empty_like(a)
        _lib.check(lib.gphm_kappa_pairs(self.kernel_id, order, _lib.ptr(a), _lib.ptr(b), a.numel(), _lib.ptr(theta), Q,
                                        _lib.ptr(out), _lib.stream_ptr()), "gphm_kappa_pairs")
        return out

    def kappa(self, x1, y1, paras):
        return self._pairs(x1, y1, paras, 0)

    def D_x1_kappa(self, x1, y1, paras):      # cov(f'(x1), f(y1))   kernel_matrix.py:49-52
        return self._pairs(x1, y1, paras, 1)

    def DD_x1_kappa(self, x1, y1, paras):     # cov(f''(x1), f(y1))  kernel_matrix.py:54-57
        return self._pairs(x1, y1, paras, 2)

    def gram(self, x1, x2, paras, deriv_order=0, jitter=0.0):
        """(len(x1), len(x2)) Gram of the deriv_order-th x1-derivative from the 1-D coordinate
        vectors (what the reference builds by vmapping over np.meshgrid pair grids)."""
        if self.kernel_id is None:
            raise NotImplementedError
        lib = _lib.load()
        a, b = as_dev(x1).reshape(-1), as_dev(x2).reshape(-1)
        theta, Q = pack_theta(paras)
        out = torch.empty((a.numel(), b.numel()), dtype=DT, device=a.device)
        _lib.check(lib.gphm_gram(self.kernel_id, deriv_order, _lib.ptr(a), a.numel(), _lib.ptr(b), b.numel(),
                                 _lib.ptr(theta), Q, float(jitter), _lib.ptr(out), _lib.stream_ptr()), "gphm_gram")
        return out

    def update_key(self, key):
        self.key = key


class SE_Cos_1d(Kernel_1d):
    """weight x SE x cosine (kernel_matrix.py:107-128)."""
    kernel_id = _lib.KERNEL_IDS["SE_Cos_1d"]


class Matern52_Cos_1d(Kernel_1d):
    """weight x Matern-5/2 x cosine (kernel_matrix.py:131-155)."""
    kernel_id = _lib.KERNEL_IDS["Matern52_Cos_1d"]


class Matern52_1d(Kernel_1d):
    """weight x Matern-5/2 (kernel_matrix.py:158-176)."""
    kernel_id = _lib.KERNEL_IDS["Matern52_1d"]


class SE_1d(Kernel_1d):
    """weight x SE (kernel_matrix.py:179-193)."""
    kernel_id = _lib.KERNEL_IDS["SE_1d"]


KERNELS = {"SE_Cos_1d": SE_Cos_1d, "Matern52_Cos_1d": Matern52_Cos_1d, "Matern52_1d": Matern52_1d, "SE_1d": SE_1d}


class Kernel_matrix(object):
    """kernel_matrix.py:12-30."""

    def __init__(self, jitter, K_u):
        self.jitter = jitter
        self.K_u = K_u

    def get_kernel_matrix(self, X1, X2, paras):
        """X1, X2: the flattened N^2 pair grids the reference passes (X1[i*N+j] = x_i,
        X2[i*N+j] = x_j).  Returns the (N, N) Gram + jitter * I."""
        X1, X2 = as_dev(X1).reshape(-1), as_dev(X2).reshape(-1)
        N = int(round(math.sqrt(X1.numel())))
        if N * N != X1.numel() or X2.numel() != X1.numel():
            raise ValueError("get_kernel_matrix expects two flattened N*N pair grids")
        K = self.K_u.kappa(X1, X2, paras).reshape(N, N)
        K.diagonal().add_(self.jitter)
        return K
