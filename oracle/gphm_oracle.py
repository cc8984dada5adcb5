"""CPU oracle for the GP-HM log-joint + gradient + Adam path.   *** TEST INFRASTRUCTURE ***

This file is a torch-FP64 **CPU restatement** of the reference algorithm.  It is the checker,
never the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The shipped package never imports anything in `oracle/`.

Parity status: PINNED.  (1) 1-D Poisson and 2-D Poisson with Matern52_Cos_1d by the reference's own two
golden runs (tests/golden/*.npz, converted by tests/golden/make_golden.py from code/result_log/**.pkl;
replayed by tests/test_oracle_golden.py).  (2) Every kernel class (SE_Cos_1d, Matern52_Cos_1d,
Matern52_1d, SE_1d) x {Poisson, Allen-Cahn} in 1-D and 2-D and advection by vectors obtained from
EXECUTING the reference's unmodified sources with a torch-backed stand-in for the jax / optax API
(tests/golden/make_ref_exec_golden.py -> tests/golden/ref_exec.npz; checked by
tests/test_oracle_ref_exec.py): loss, every gradient leaf, two Adam steps, predictions.  Not pinned:
XLA's own evaluation order, and optax beyond what the two result logs exercise (restated from its
published algorithm).

Reference lines followed (all under /root/reference/code):
  kernel_matrix.py:21-30     Kernel_matrix.get_kernel_matrix  (vmap(kappa) + jitter*I)
  kernel_matrix.py:49-57     D_x1_kappa / DD_x1_kappa         (grad / grad-grad of kappa wrt x1)
  kernel_matrix.py:114-128   SE_Cos_1d.kappa
  kernel_matrix.py:138-155   Matern52_Cos_1d.kappa
  kernel_matrix.py:163-176   Matern52_1d.kappa
  kernel_matrix.py:184-193   SE_1d.kappa
  model_GP_solver_2d.py:87-183         value_and_grad_kernel / boundary_and_eq_gap / loss / step
  model_GP_solver_2d.py:185-220        preds
  model_GP_solver_1d.py:80-158,160-180 same for 1-D
  model_GP_solver_advection.py:87-179  same with D_x1_kappa and beta*U_x + U_y
  optax 0.1.4 `adam(lr)` (b1=.9, b2=.999, eps=1e-8, eps_root=0, bias corrected) - third-party,
  not in the tree; restated in `adam_update`.

Two formulations of the same mathematics:
  * literal  - forms the (N,N,Q) tensors, LU `solve` + `slogdet`, gradients by autograd.  This
               mirrors the reference dataflow line by line and is the parity oracle.
  * efficient- closed-form Toeplitz Gram tables, Cholesky, hand-derived backward.  Same 28 N^3
               dataflow as the GPU path; used as the timed CPU baseline (it is the *stronger*
               baseline) and cross-checked against `literal` in tests.

Autodiff-of-abs convention (parity critical): kernels are functions of d=|x1-y1| and the
reference differentiates through jnp.abs, whose JVP takes the +1 branch at 0.  Hence the
second-derivative Gram has the analytic k''(0) on its diagonal and the first-derivative Gram
is k'(d)*sgn(x1-y1) with sgn(0)=+1 (k'(0)=0 anyway).  Closed forms below are in d >= 0.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

DT = torch.float64
KERNEL_NAMES = ("SE_Cos_1d", "Matern52_Cos_1d", "Matern52_1d", "SE_1d")
SQRT5 = math.sqrt(5.0)
TWO_PI = 2.0 * math.pi


# ----------------------------------------------------------------------------------------------
# closed-form kernels: k, k', k'' in d >= 0 and their partials wrt (log-w, log-ls, freq)
# ----------------------------------------------------------------------------------------------
def kernel_terms(name: str, d: torch.Tensor, lw: torch.Tensor, ls: torch.Tensor, f: torch.Tensor,
                 order: int, partials: bool = False):
    """Per-component terms of the order-th d-derivative of kappa.

    d: (...,) distances >= 0;  lw, ls, f: (Q,).  Returns `term` of shape (..., Q) whose sum over
    the last axis is k (order 0), k' (1) or k'' (2); with partials=True also returns
    (d_lw, d_ls, d_f), each (..., Q): the partial derivative of that sum wrt the q-th parameter.
    kernel_matrix.py:114-128 (SE*cos), :138-155 (Matern52*cos), :163-176, :184-193 (no cosine).
    """
    assert name in KERNEL_NAMES, name
    dd = d.unsqueeze(-1)
    w = torch.exp(lw)
    has_cos = name in ("SE_Cos_1d", "Matern52_Cos_1d")
    matern = name in ("Matern52_Cos_1d", "Matern52_1d")

    if matern:
        a = SQRT5 * torch.exp(ls)
        t = a * dd
        e = torch.exp(-t)
        m0 = (1.0 + t + t * t / 3.0) * e
        m1 = -(t / 3.0) * (1.0 + t) * e
        m2 = -(1.0 / 3.0) * (1.0 + t - t * t) * e
        m3 = (t / 3.0) * (3.0 - t) * e
        b0, b1, b2 = m0, a * m1, a * a * m2
        bl0, bl1, bl2 = t * m1, a * (m1 + t * m2), a * a * (2.0 * m2 + t * m3)
    else:
        l = torch.exp(ls)
        d2 = dd * dd
        s = torch.exp(-l * d2)
        b0 = s
        b1 = -2.0 * l * dd * s
        b2 = (4.0 * l * l * d2 - 2.0 * l) * s
        bl0 = -l * d2 * s
        bl1 = (-2.0 * l * dd + 2.0 * l * l * d2 * dd) * s
        bl2 = (-2.0 * l + 10.0 * l * l * d2 - 4.0 * l * l * l * d2 * d2) * s

    if has_cos:
        om = TWO_PI * f
        c = torch.cos(om * dd)
        sn = torch.sin(om * dd)
        c0, c1, c2 = c, -om * sn, -om * om * c
        cf0 = -TWO_PI * dd * sn
        cf1 = -TWO_PI * sn - TWO_PI * om * dd * c
        cf2 = -2.0 * TWO_PI * om * c + TWO_PI * om * om * dd * sn
    else:
        one = torch.ones_like(b0)
        zero = torch.zeros_like(b0)
        c0, c1, c2 = one, zero, zero
        cf0 = cf1 = cf2 = zero

    if order == 0:
        term = w * b0 * c0
        if partials:
            return term, (term, w * bl0 * c0, w * b0 * cf0)
    elif order == 1:
        term = w * (b1 * c0 + b0 * c1)
        if partials:
            return term, (term, w * (bl1 * c0 + bl0 * c1), w * (b1 * cf0 + b0 * cf1))
    elif order == 2:
        term = w * (b2 * c0 + 2.0 * b1 * c1 + b0 * c2)
        if partials:
            return term, (term, w * (bl2 * c0 + 2.0 * bl1 * c1 + bl0 * c2),
                          w * (b2 * cf0 + 2.0 * b1 * cf1 + b0 * cf2))
    else:
        raise ValueError(order)
    return term


def kappa_scalar(name: str, x1: torch.Tensor, y1: torch.Tensor, theta: Dict[str, torch.Tensor]):
    """Literal scalar kappa written exactly like the reference (used to verify closed forms by
    autograd away from x1 == y1).  kernel_matrix.py:114-128,138-155,163-176,184-193."""
    lw, ls, f = theta["log-w"], theta["log-ls"], theta["freq"]
    d = torch.abs(x1 - y1)
    if name == "SE_Cos_1d":
        return (torch.exp(lw) * torch.exp(-d ** 2 * torch.exp(ls)) * torch.cos(TWO_PI * d * f)).sum()
    if name == "Matern52_Cos_1d":
        mat = (1 + SQRT5 * d * torch.exp(ls) + 5 / 3 * d ** 2 * torch.exp(ls) ** 2) * torch.exp(-SQRT5 * d * torch.exp(ls))
        return (torch.exp(lw) * mat * torch.cos(TWO_PI * d * f)).sum()
    if name == "Matern52_1d":
        mat = (1 + SQRT5 * d * torch.exp(ls) + 5 / 3 * d ** 2 * torch.exp(ls) ** 2) * torch.exp(-SQRT5 * d * torch.exp(ls))
        return (torch.exp(lw) * mat).sum()
    if name == "SE_1d":
        return (torch.exp(lw) * torch.exp(-d ** 2 * torch.exp(ls))).sum()
    raise ValueError(name)


def gram(name: str, x1: torch.Tensor, x2: torch.Tensor, theta: Dict[str, torch.Tensor], order: int = 0,
         jitter: float = 0.0) -> torch.Tensor:
    """(len(x1), len(x2)) Gram of the order-th x1-derivative of kappa (literal (N,M,Q) formulation).

    order 0 + jitter: Kernel_matrix.get_kernel_matrix (kernel_matrix.py:21-30; square only there);
    order 0, rectangular, no jitter: the cross-Grams of preds (model_GP_solver_2d.py:198-202);
    order 1: vmap(D_x1_kappa) = k'(d) * sgn(x1-x2), sgn(0)=+1 (model_GP_solver_advection.py:107-117);
    order 2: vmap(DD_x1_kappa) = k''(d), diagonal = k''(0)    (model_GP_solver_2d.py:107-117).
    """
    diff = x1.reshape(-1, 1) - x2.reshape(1, -1)
    d = diff.abs()
    term = kernel_terms(name, d, theta["log-w"], theta["log-ls"], theta["freq"], order)
    out = term.sum(-1)
    if order == 1:
        sgn = torch.where(diff >= 0, torch.ones_like(diff), -torch.ones_like(diff))
        out = out * sgn
    if jitter:
        assert out.shape[0] == out.shape[1]
        out = out + jitter * torch.eye(out.shape[0], dtype=out.dtype)
    return out


# ----------------------------------------------------------------------------------------------
# problems
# ----------------------------------------------------------------------------------------------
@dataclass
class Problem2D:
    """Constants the reference bakes into the jitted executable via static `self`
    (model_GP_solver_2d.py:40-85; advection: model_GP_solver_advection.py:40-85)."""
    kernel: str
    eq_type: str                      # 'poisson_2d' | 'allencahn_2d' | 'advection'
    x: torch.Tensor                   # (N1,)
    y: torch.Tensor                   # (N2,)
    src: torch.Tensor                 # (N1,N2)
    bvals: torch.Tensor               # (2*N2+2*N1,) edges U[0,:],U[-1,:],U[:,0],U[:,-1]
    llk_weight: float = 200.0
    logdet: float = 1.0
    beta: float = 1.0
    jitter: float = 1e-6

    @property
    def deriv_order(self):
        return 1 if self.eq_type == "advection" else 2

    @property
    def c1(self):
        return self.beta if self.eq_type == "advection" else 1.0


@dataclass
class Problem1D:
    """model_GP_solver_1d.py:38-78."""
    kernel: str
    eq_type: str                      # 'poisson_1d' | 'allencahn_1d'
    x: torch.Tensor                   # (N,)
    src: torch.Tensor                 # (N,)
    xind: torch.Tensor                # (Nb,) long
    yb: torch.Tensor                  # (Nb,)
    llk_weight: float = 200.0
    logdet: float = 1.0
    jitter: float = 1e-6


def init_params_2d(N1: int, N2: int, Q: int, freq_scale: float) -> Dict:
    """model_GP_solver_2d.py:245-261."""
    def kp():
        return {"log-w": torch.full((Q,), math.log(1.0 / Q), dtype=DT),
                "log-ls": torch.zeros(Q, dtype=DT),
                "freq": torch.linspace(0, 1, Q, dtype=DT) * freq_scale}
    return {"log_tau": torch.zeros((), dtype=DT), "log_v": torch.zeros((), dtype=DT),
            "kernel_paras_1": kp(), "kernel_paras_2": kp(), "U": torch.zeros(N1, N2, dtype=DT)}


def init_params_1d(N: int, Q: int, freq_scale: float) -> Dict:
    """model_GP_solver_1d.py:203-213."""
    return {"log_tau": torch.zeros((), dtype=DT), "log_v": torch.zeros((), dtype=DT),
            "kernel_paras": {"log-w": torch.full((Q,), math.log(1.0 / Q), dtype=DT),
                             "log-ls": torch.zeros(Q, dtype=DT),
                             "freq": torch.linspace(0, 1, Q, dtype=DT) * freq_scale},
            "u": torch.zeros(N, 1, dtype=DT)}


def _nonlin(eq_type: str, U: torch.Tensor):
    if eq_type.startswith("allencahn"):
        return U * (U * U - 1.0)
    return torch.zeros_like(U)


def boundary_vector_2d(U: torch.Tensor) -> torch.Tensor:
    """hstack(U[0,:], U[-1,:], U[:,0], U[:,-1]) - corners appear twice (model_GP_solver_2d.py:127)."""
    return torch.cat((U[0, :], U[-1, :], U[:, 0], U[:, -1]))


# ----------------------------------------------------------------------------------------------
# literal formulation (parity oracle): LU solve + slogdet + autograd
# ----------------------------------------------------------------------------------------------
def forward_terms_2d(p: Problem2D, params: Dict) -> Dict[str, torch.Tensor]:
    """model_GP_solver_2d.py:87-174 (advection: model_GP_solver_advection.py:87-170)."""
    U = params["U"]
    th1, th2 = params["kernel_paras_1"], params["kernel_paras_2"]
    K1 = gram(p.kernel, p.x, p.x, th1, 0, p.jitter)
    K2 = gram(p.kernel, p.y, p.y, th2, 0, p.jitter)
    A = torch.linalg.solve(K1, U)                    # K1inv_U   (N1,N2)
    B = torch.linalg.solve(K2, U.T)                  # K2inv_Ut  (N2,N1)
    D1 = gram(p.kernel, p.x, p.x, th1, p.deriv_order)
    D2 = gram(p.kernel, p.y, p.y, th2, p.deriv_order)
    Ux = D1 @ A
    Uy = (D2 @ B).T
    bgap = ((boundary_vector_2d(U) - p.bvals.reshape(-1)) ** 2).sum()
    R = p.c1 * Ux + Uy + _nonlin(p.eq_type, U) - p.src
    eqgap = (R ** 2).sum()
    N1, N2 = U.shape
    logdet1 = torch.linalg.slogdet(K1)[1]
    logdet2 = torch.linalg.slogdet(K2)[1]
    quad = (A * B.T).sum()
    log_tau, log_v = params["log_tau"], params["log_v"]
    Nb, Nc = p.bvals.numel(), N1 * N2
    log_prior = -0.5 * N2 * logdet1 * p.logdet - 0.5 * N1 * logdet2 * p.logdet - 0.5 * quad
    log_b = 0.5 * Nb * log_tau - 0.5 * torch.exp(log_tau) * bgap
    eq_ll = 0.5 * Nc * log_v - 0.5 * torch.exp(log_v) * eqgap
    loss = -(log_prior + log_b * p.llk_weight + eq_ll)
    return {"loss": loss, "logdet1": logdet1, "logdet2": logdet2, "quad": quad, "bgap": bgap,
            "eqgap": eqgap, "Ux": Ux, "Uy": Uy, "A": A, "Bt": B.T, "K1": K1, "K2": K2}


def forward_terms_1d(p: Problem1D, params: Dict) -> Dict[str, torch.Tensor]:
    """model_GP_solver_1d.py:80-149."""
    u = params["u"].reshape(-1, 1)
    th = params["kernel_paras"]
    K = gram(p.kernel, p.x, p.x, th, 0, p.jitter)
    a = torch.linalg.solve(K, u)
    D = gram(p.kernel, p.x, p.x, th, 2)
    uxx = D @ a
    bgap = ((u[p.xind].reshape(-1) - p.yb.reshape(-1)) ** 2).sum()
    r = uxx.reshape(-1) + _nonlin(p.eq_type, u).reshape(-1) - p.src.reshape(-1)
    eqgap = (r ** 2).sum()
    logdet = torch.linalg.slogdet(K)[1]
    quad = (u * a).sum()
    log_tau, log_v = params["log_tau"], params["log_v"]
    Nb, Nc = p.xind.numel(), u.shape[0]
    log_prior = -0.5 * logdet * p.logdet - 0.5 * quad
    log_b = 0.5 * Nb * log_tau - 0.5 * torch.exp(log_tau) * bgap
    eq_ll = 0.5 * Nc * log_v - 0.5 * torch.exp(log_v) * eqgap
    loss = -(log_prior + log_b * p.llk_weight + eq_ll)
    return {"loss": loss, "logdet1": logdet, "quad": quad, "bgap": bgap, "eqgap": eqgap,
            "Ux": uxx, "A": a, "K1": K}


def loss_extra_literal(p: Problem1D, kernel_extra: str, params: Dict, params_extra: Dict) -> torch.Tensor:
    """model_GP_solver_1d_extra.py:107-141: second-stage loss with the first GP (params) frozen.
    params_extra: {'u' (N,1), 'kernel_paras': {'log-w','log-ls'}, 'log_tau', 'log_v'}."""
    fw = forward_terms_1d(p, params)
    u, uxx = params["u"].reshape(-1, 1), fw["Ux"]
    ue = params_extra["u"].reshape(u.shape[0], -1).sum(1, keepdim=True)
    kp = dict(params_extra["kernel_paras"])
    kp.setdefault("freq", torch.zeros_like(kp["log-w"]))
    K = gram(kernel_extra, p.x, p.x, kp, 0, p.jitter)
    a = torch.linalg.solve(K, ue)
    uxxe = gram(kernel_extra, p.x, p.x, kp, 2) @ a
    bgap = ((u[p.xind].reshape(-1) + ue[p.xind].reshape(-1) - p.yb.reshape(-1)) ** 2).sum()
    r = uxx.reshape(-1) + uxxe.reshape(-1) - p.src.reshape(-1)
    if p.eq_type.startswith("allencahn"):
        t = (u + ue).reshape(-1)
        r = r + t * (t * t - 1.0)
    eqgap = (r ** 2).sum()
    log_prior = -0.5 * torch.linalg.slogdet(K)[1] * p.logdet - 0.5 * (ue * a).sum()
    log_b = 0.5 * p.xind.numel() * params_extra["log_tau"] - 0.5 * torch.exp(params_extra["log_tau"]) * bgap
    eq_ll = 0.5 * u.shape[0] * params_extra["log_v"] - 0.5 * torch.exp(params_extra["log_v"]) * eqgap
    return -(log_prior + log_b * p.llk_weight + eq_ll)


def _clone_leaves(params: Dict, requires_grad: bool) -> Dict:
    out = {}
    for k, v in params.items():
        if isinstance(v, dict):
            out[k] = _clone_leaves(v, requires_grad)
        else:
            out[k] = torch.as_tensor(v, dtype=DT).clone().detach().requires_grad_(requires_grad)
    return out


def flatten(params: Dict):
    """Deterministic (path, tensor) list of the params pytree."""
    out = []
    for k in sorted(params.keys()):
        v = params[k]
        if isinstance(v, dict):
            out += [(k + "/" + kk, vv) for kk, vv in flatten(v)]
        else:
            out.append((k, v))
    return out


def loss_and_grad_literal(p, params: Dict) -> Tuple[Dict[str, float], Dict]:
    """jax.value_and_grad(self.loss)(params, key)  (model_GP_solver_2d.py:179, _1d.py:154)."""
    q = _clone_leaves(params, True)
    fw = forward_terms_2d(p, q) if isinstance(p, Problem2D) else forward_terms_1d(p, q)
    leaves = [t for _, t in flatten(q)]
    grads = torch.autograd.grad(fw["loss"], leaves, allow_unused=True)
    gd = {}
    for (path, leaf), g in zip(flatten(q), grads):
        g = torch.zeros_like(leaf) if g is None else g
        node = gd
        parts = path.split("/")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = g.detach()
    terms = {k: float(fw[k].detach()) for k in ("loss", "logdet1", "logdet2", "quad", "bgap", "eqgap") if k in fw}
    return terms, gd


# ----------------------------------------------------------------------------------------------
# efficient formulation (timed CPU baseline): Toeplitz tables, Cholesky, analytic backward
# ----------------------------------------------------------------------------------------------
def is_uniform(x: torch.Tensor, rtol: float = 1e-9) -> bool:
    if x.numel() < 3:
        return True
    h = x[1:] - x[:-1]
    return bool((h - h[0]).abs().max() <= rtol * h.abs().max())


def _toeplitz_from_table(tbl: torch.Tensor, antisym: bool = False) -> torch.Tensor:
    N = tbl.numel()
    idx = torch.arange(N)
    lag = idx.reshape(-1, 1) - idx.reshape(1, -1)
    M = tbl[lag.abs()]
    if antisym:                       # k'(d) * sgn(x_i - x_j): + below the diagonal (x ascending)
        M = torch.where(lag >= 0, M, -M)
    return M


def _diag_sums(M: torch.Tensor, antisym: bool = False) -> torch.Tensor:
    """s[m] = sum over |i-j| = m of M[i,j]  (antisym: lower minus upper)."""
    N = M.shape[0]
    Z = torch.zeros(N, 2 * N, dtype=M.dtype)
    Z[:, :N] = M.flip(1)
    anti = Z.reshape(-1)[: N * (2 * N - 1)].reshape(N, 2 * N - 1).sum(0)   # index c = i + (N-1-j)
    lower = anti[N - 1:]            # i - j = 0..N-1
    upper = anti[:N].flip(0)        # j - i = 0..N-1
    s = lower - upper if antisym else lower + upper
    s = s.clone()
    s[0] = 0.0 if antisym else lower[0]
    return s


def _theta_grad_axis(name, x, theta, order, Kbar, Dbar):
    """sum_ij Kbar*dK/dtheta + Dbar*dD/dtheta  ->  dict of three (Q,) vectors."""
    lw, ls, f = theta["log-w"], theta["log-ls"], theta["freq"]
    if is_uniform(x):
        d = (x - x[0]).abs()
        sK = _diag_sums(Kbar)
        sD = _diag_sums(Dbar, antisym=(order == 1))
        if order == 1 and x[-1] < x[0]:
            sD = -sD
        _, pK = kernel_terms(name, d, lw, ls, f, 0, True)
        _, pD = kernel_terms(name, d, lw, ls, f, order, True)
        g = [sK @ pk + sD @ pd for pk, pd in zip(pK, pD)]
    else:
        diff = x.reshape(-1, 1) - x.reshape(1, -1)
        d = diff.abs()
        sgn = torch.where(diff >= 0, torch.ones_like(diff), -torch.ones_like(diff)) if order == 1 else 1.0
        _, pK = kernel_terms(name, d, lw, ls, f, 0, True)
        _, pD = kernel_terms(name, d, lw, ls, f, order, True)
        g = [(Kbar.unsqueeze(-1) * pk).sum((0, 1)) + ((Dbar * sgn).unsqueeze(-1) * pd).sum((0, 1))
             for pk, pd in zip(pK, pD)]
    return {"log-w": g[0], "log-ls": g[1], "freq": g[2]}


def _gram_pair(name, x, theta, order, jitter):
    lw, ls, f = theta["log-w"], theta["log-ls"], theta["freq"]
    if is_uniform(x):
        d = (x - x[0]).abs()
        K = _toeplitz_from_table(kernel_terms(name, d, lw, ls, f, 0).sum(-1))
        D = _toeplitz_from_table(kernel_terms(name, d, lw, ls, f, order).sum(-1), antisym=(order == 1))
        if order == 1 and x[-1] < x[0]:
            D = -D
    else:
        th = {"log-w": lw, "log-ls": ls, "freq": f}
        K = gram(name, x, x, th, 0)
        D = gram(name, x, x, th, order)
    K = K + jitter * torch.eye(K.shape[0], dtype=DT)
    return K, D


def loss_and_grad_efficient(p, params: Dict) -> Tuple[Dict[str, float], Dict]:
    """Same value/gradient as `loss_and_grad_literal`, computed the way the GPU path does
    (SURVEY App. C): Cholesky, K^-1 applications, hand-derived backward, diagonal-sum theta
    gradients on uniform grids.  No autograd, no (N,N,Q) tensors on uniform grids."""
    with torch.no_grad():
        if isinstance(p, Problem1D):
            return _efficient_1d(p, params)
        return _efficient_2d(p, params)


def _efficient_2d(p: Problem2D, params):
    U = torch.as_tensor(params["U"], dtype=DT)
    th1, th2 = params["kernel_paras_1"], params["kernel_paras_2"]
    tau, v = float(params["log_tau"]), float(params["log_v"])
    N1, N2 = U.shape
    order, c1, lam, ld = p.deriv_order, p.c1, p.llk_weight, float(p.logdet)
    K1, D1 = _gram_pair(p.kernel, p.x, th1, order, p.jitter)
    K2, D2 = _gram_pair(p.kernel, p.y, th2, order, p.jitter)
    L1 = torch.linalg.cholesky(K1)
    L2 = torch.linalg.cholesky(K2)
    logdet1 = 2.0 * torch.log(torch.diagonal(L1)).sum()
    logdet2 = 2.0 * torch.log(torch.diagonal(L2)).sum()
    A = torch.cholesky_solve(U, L1)                        # K1^-1 U
    Bt = torch.cholesky_solve(U.T.contiguous(), L2).T      # U K2^-1
    Ux = D1 @ A
    Uy = Bt @ D2.T
    R = c1 * Ux + Uy + _nonlin(p.eq_type, U) - p.src
    eqgap = (R * R).sum()
    eb = boundary_vector_2d(U) - p.bvals.reshape(-1)
    bgap = (eb * eb).sum()
    quad = (A * Bt).sum()
    Nb, Nc = p.bvals.numel(), N1 * N2
    loss = (0.5 * ld * (N2 * logdet1 + N1 * logdet2) + 0.5 * quad
            - lam * (0.5 * Nb * tau - 0.5 * math.exp(tau) * bgap)
            - (0.5 * Nc * v - 0.5 * math.exp(v) * eqgap))
    # backward
    G = math.exp(v) * R
    W = torch.cholesky_solve(Bt, L1)
    S1 = torch.cholesky_solve(c1 * (D1.T @ G), L1)
    S2 = torch.cholesky_solve((G @ D2).T.contiguous(), L2).T
    gU = W + S1 + S2
    if p.eq_type.startswith("allencahn"):
        gU = gU + G * (3.0 * U * U - 1.0)
    s = lam * math.exp(tau)
    gU[0, :] += s * eb[:N2]
    gU[-1, :] += s * eb[N2:2 * N2]
    gU[:, 0] += s * eb[2 * N2:2 * N2 + N1]
    gU[:, -1] += s * eb[2 * N2 + N1:]
    K1inv = torch.cholesky_inverse(L1)
    K2inv = torch.cholesky_inverse(L2)
    K1bar = 0.5 * ld * N2 * K1inv - (S1 + 0.5 * W) @ A.T
    D1bar = c1 * (G @ A.T)
    K2bar = 0.5 * ld * N1 * K2inv - (S2 + 0.5 * W).T @ Bt
    D2bar = G.T @ Bt
    g1 = _theta_grad_axis(p.kernel, p.x, th1, order, K1bar, D1bar)
    g2 = _theta_grad_axis(p.kernel, p.y, th2, order, K2bar, D2bar)
    grads = {"U": gU, "kernel_paras_1": g1, "kernel_paras_2": g2,
             "log_tau": torch.tensor(-lam * (0.5 * Nb - 0.5 * math.exp(tau) * float(bgap)), dtype=DT),
             "log_v": torch.tensor(-(0.5 * Nc - 0.5 * math.exp(v) * float(eqgap)), dtype=DT)}
    terms = {"loss": float(loss), "logdet1": float(logdet1), "logdet2": float(logdet2), "quad": float(quad),
             "bgap": float(bgap), "eqgap": float(eqgap)}
    return terms, grads


def _efficient_1d(p: Problem1D, params):
    u = torch.as_tensor(params["u"], dtype=DT).reshape(-1, 1)
    th = params["kernel_paras"]
    tau, v = float(params["log_tau"]), float(params["log_v"])
    N = u.shape[0]
    lam, ld = p.llk_weight, float(p.logdet)
    K, D = _gram_pair(p.kernel, p.x, th, 2, p.jitter)
    L = torch.linalg.cholesky(K)
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum()
    a = torch.cholesky_solve(u, L)
    uxx = D @ a
    r = uxx + _nonlin(p.eq_type, u) - p.src.reshape(-1, 1)
    eqgap = (r * r).sum()
    eb = u[p.xind].reshape(-1) - p.yb.reshape(-1)
    bgap = (eb * eb).sum()
    quad = (u * a).sum()
    Nb = p.xind.numel()
    loss = (0.5 * ld * logdet + 0.5 * quad - lam * (0.5 * Nb * tau - 0.5 * math.exp(tau) * bgap)
            - (0.5 * N * v - 0.5 * math.exp(v) * eqgap))
    g = math.exp(v) * r
    s = torch.cholesky_solve(D.T @ g, L)
    gu = a + s
    if p.eq_type.startswith("allencahn"):
        gu = gu + g * (3.0 * u * u - 1.0)
    gu.reshape(-1).index_add_(0, p.xind, lam * math.exp(tau) * eb)
    Kinv = torch.cholesky_inverse(L)
    Kbar = 0.5 * ld * Kinv - (s + 0.5 * a) @ a.T
    Dbar = g @ a.T
    gth = _theta_grad_axis(p.kernel, p.x, th, 2, Kbar, Dbar)
    grads = {"u": gu, "kernel_paras": gth,
             "log_tau": torch.tensor(-lam * (0.5 * Nb - 0.5 * math.exp(tau) * float(bgap)), dtype=DT),
             "log_v": torch.tensor(-(0.5 * N - 0.5 * math.exp(v) * float(eqgap)), dtype=DT)}
    terms = {"loss": float(loss), "logdet1": float(logdet), "quad": float(quad), "bgap": float(bgap),
             "eqgap": float(eqgap)}
    return terms, grads


# ----------------------------------------------------------------------------------------------
# Adam (optax 0.1.4 adam(lr): scale_by_adam(b1=.9,b2=.999,eps=1e-8,eps_root=0) then scale(-lr))
# ----------------------------------------------------------------------------------------------
def adam_init(params: Dict) -> Dict:
    def z(t):
        return {k: z(v) for k, v in t.items()} if isinstance(t, dict) else torch.zeros_like(torch.as_tensor(t, dtype=DT))
    return {"count": 0, "mu": z(params), "nu": z(params)}


def adam_update(params: Dict, grads: Dict, state: Dict, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """Functional: returns (new_params, new_state).  model_GP_solver_2d.py:180-182."""
    t = state["count"] + 1
    c1, c2 = 1.0 - b1 ** t, 1.0 - b2 ** t

    def rec(p, g, m, n):
        if isinstance(p, dict):
            out = {k: rec(p[k], g[k], m[k], n[k]) for k in p}
            return ({k: o[0] for k, o in out.items()}, {k: o[1] for k, o in out.items()},
                    {k: o[2] for k, o in out.items()})
        p = torch.as_tensor(p, dtype=DT)
        m2 = b1 * m + (1.0 - b1) * g
        n2 = b2 * n + (1.0 - b2) * g * g
        upd = (m2 / c1) / (torch.sqrt(n2 / c2) + eps)
        return p - lr * upd, m2, n2

    newp, mu, nu = rec(params, grads, state["mu"], state["nu"])
    return newp, {"count": t, "mu": mu, "nu": nu}


def step(p, params: Dict, state: Dict, lr: float, formulation: str = "literal"):
    """params, opt_state, loss = self.step(params, opt_state, key)  (model_GP_solver_2d.py:176-183)."""
    fn = loss_and_grad_literal if formulation == "literal" else loss_and_grad_efficient
    terms, grads = fn(p, params)
    newp, news = adam_update(params, grads, state, lr)
    return newp, news, terms


# ----------------------------------------------------------------------------------------------
# prediction
# ----------------------------------------------------------------------------------------------
def preds_2d(p: Problem2D, params: Dict, xte: torch.Tensor, yte: torch.Tensor) -> torch.Tensor:
    """model_GP_solver_2d.py:185-220: Kmn1 K1^-1 U K2^-1 Kmn2^T (no jitter on cross-Grams)."""
    with torch.no_grad():
        th1, th2 = params["kernel_paras_1"], params["kernel_paras_2"]
        U = torch.as_tensor(params["U"], dtype=DT)
        K1 = gram(p.kernel, p.x, p.x, th1, 0, p.jitter)
        K2 = gram(p.kernel, p.y, p.y, th2, 0, p.jitter)
        M1 = gram(p.kernel, xte, p.x, th1, 0) @ torch.linalg.solve(K1, U)
        M2 = torch.linalg.solve(K2, M1.T)
        return (gram(p.kernel, yte, p.y, th2, 0) @ M2).T


def preds_1d(p: Problem1D, params: Dict, xte: torch.Tensor) -> torch.Tensor:
    """model_GP_solver_1d.py:160-180."""
    with torch.no_grad():
        th = params["kernel_paras"]
        u = torch.as_tensor(params["u"], dtype=DT).reshape(-1, 1)
        K = gram(p.kernel, p.x, p.x, th, 0, p.jitter)
        return gram(p.kernel, xte.reshape(-1), p.x, th, 0) @ torch.linalg.solve(K, u)


def rel_l2(pred: torch.Tensor, truth: torch.Tensor) -> float:
    """model_GP_solver_2d.py:297-300."""
    return float(torch.linalg.norm(pred.reshape(-1) - truth.reshape(-1)) / torch.linalg.norm(truth.reshape(-1)))


# ----------------------------------------------------------------------------------------------
# manufactured problems (model_GP_solver_2d.py:355-416, _1d.py:299-354, _advection.py:354-410)
# ----------------------------------------------------------------------------------------------
EQUATIONS_1D = {
    "poisson_1d-mix_sin": lambda x: torch.sin(x) + 0.1 * torch.sin(20 * x) + 0.05 * torch.sin(100 * x),
    "poisson_1d-single_sin": lambda x: torch.sin(100 * x),
    "poisson_1d-sin_cos": lambda x: torch.sin(6 * x) * torch.cos(100 * x),
    "poisson_1d-x_time_sinx": lambda x: x * torch.sin(200 * x),
    "poisson_1d-x2_add_sinx": lambda x: torch.sin(500 * x) - 2 * (x - 0.5) ** 2,
    "allencahn_1d-sin_cos": lambda x: torch.sin(6 * x) * torch.cos(100 * x),
    "allencahn_1d-single_sin": lambda x: torch.sin(100 * x),
}
EQUATIONS_2D = {
    "poisson_2d-sin_sin": lambda x, y: torch.sin(100 * x) * torch.sin(100 * y),
    "poisson_2d-sin_cos": lambda x, y: torch.sin(100 * x) * torch.cos(100 * y),
    "poisson_2d-sin_add_cos": lambda x, y: torch.sin(6 * x) * torch.cos(20 * x) + torch.sin(6 * y) * torch.cos(20 * y),
    "allencahn_2d-mix-sincos": lambda x, y: (torch.sin(x) + 0.1 * torch.sin(20 * x) + torch.cos(100 * x)) *
                                            (torch.sin(y) + 0.1 * torch.sin(20 * y) + torch.cos(100 * y)),
}


def _derivs(fn, args, wrt: int, order: int):
    args = [a.clone().detach().requires_grad_(True) for a in args]
    out = fn(*args)
    for _ in range(order):
        (out,) = torch.autograd.grad(out.sum(), args[wrt], create_graph=True)
    return out.detach()


def make_problem_1d(equation: str, kernel: str, N: int, scale: float, llk_weight=200.0, logdet=1.0, M=300):
    u = EQUATIONS_1D[equation]
    eq_type = equation.split("-")[0]
    x = torch.linspace(0, 1, N, dtype=DT) * scale
    xte = torch.linspace(0, 1, M, dtype=DT) * scale
    src = _derivs(u, [x], 0, 2)
    if eq_type == "allencahn_1d":
        src = src + u(x) * (u(x) ** 2 - 1)
    xind = torch.tensor([0, N - 1])
    p = Problem1D(kernel, eq_type, x, src, xind, u(x[xind]), float(llk_weight), float(logdet))
    return p, xte, u(xte)


def make_problem_2d(equation: str, kernel: str, N: int, scale: float, llk_weight=200.0, logdet=1.0, beta=1.0,
                    M=300, N2: Optional[int] = None):
    N2 = N if N2 is None else N2
    if equation.startswith("advection"):
        u = lambda x, y: torch.sin(x - beta * y)
    else:
        u = EQUATIONS_2D[equation]
    eq_type = equation.split("-")[0]
    x = torch.linspace(0, 1, N, dtype=DT) * scale
    y = torch.linspace(0, 1, N2, dtype=DT) * scale
    X, Y = torch.meshgrid(x, y, indexing="ij")
    if eq_type == "advection":
        src = beta * _derivs(u, [X, Y], 0, 1) + _derivs(u, [X, Y], 1, 1)
    else:
        src = _derivs(u, [X, Y], 0, 2) + _derivs(u, [X, Y], 1, 2)
        if eq_type == "allencahn_2d":
            src = src + u(X, Y) * (u(X, Y) ** 2 - 1)
    bvals = boundary_vector_2d(u(X, Y))
    p = Problem2D(kernel, eq_type, x, y, src, bvals, float(llk_weight), float(logdet), float(beta))
    xt = torch.linspace(0, 1, M, dtype=DT) * scale
    XT, YT = torch.meshgrid(xt, xt, indexing="ij")
    return p, (xt, xt), u(XT, YT)


def state_S1(p: Problem2D, Q: int = 30, freq_scale: float = 20.0) -> Dict:
    """Deterministic non-degenerate parity state S1 of SURVEY 8(d)/App. G.4."""
    X, Y = torch.meshgrid(p.x, p.y, indexing="ij")
    ustar = torch.sin(6 * X) * torch.cos(20 * X) + torch.sin(6 * Y) * torch.cos(20 * Y)
    q = torch.arange(Q, dtype=DT)

    def kp(s):
        return {"log-w": math.log(1.0 / Q) - 0.05 * torch.cos(q + s),
                "log-ls": 0.1 * torch.sin(q + s),
                "freq": freq_scale * q / (Q - 1) + 0.05 * torch.sin(2 * (q + s))}
    return {"U": 0.7 * ustar + 0.3 * torch.sin(3 * X) * torch.cos(5 * Y),
            "kernel_paras_1": kp(0), "kernel_paras_2": kp(1),
            "log_tau": torch.tensor(0.3, dtype=DT), "log_v": torch.tensor(-0.2, dtype=DT)}
