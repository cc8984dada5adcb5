"""Shared host-side machinery of the three solver classes: marshals the reference's params /
opt_state pytrees to the packed device layout of include/gphm.h and calls libgphm.

Packed layout:  U (n1*n2)  |  small = [log-w1|log-ls1|freq1|log-w2|log-ls2|freq2|log_tau|log_v]
The reference's functional `step(params, opt_state, key) -> (params, opt_state, loss)` is kept;
`train` uses the same kernels in place on one persistent packed state (no per-step allocation).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .kernel_matrix import DT, as_dev, _dev

TERM_NAMES = ("loss", "logdet1", "logdet2", "quad", "boundary_gap", "eq_gap", "d_log_tau", "d_log_v")


class PackedState(object):
    """Device-resident params + Adam state in the C-ABI layout."""

    def __init__(self, nf, Q, device):
        ns = 6 * Q + 2
        self.U = torch.zeros(nf, dtype=DT, device=device)
        self.small = torch.zeros(ns, dtype=DT, device=device)
        self.mU, self.vU = torch.zeros_like(self.U), torch.zeros_like(self.U)
        self.msmall, self.vsmall = torch.zeros_like(self.small), torch.zeros_like(self.small)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)
        self.terms = torch.zeros(8, dtype=DT, device=device)


class SolverCore(object):
    """One gphm_plan plus the pytree <-> packed conversions.  `dim` is 1 or 2."""

    def __init__(self, dim, kernel_name, eq_name, x, y, src, bvals, xind, llk_weight, logdet, beta, jitter, Q,
                 force_general=False):
        self.lib = _lib.load()
        self.device = _dev()
        self.dim = dim
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        self.n1 = x.size
        if dim == 2:
            y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
            self.n2 = y.size
        else:
            y, self.n2 = None, 1
        src = np.ascontiguousarray(np.asarray(src, dtype=np.float64).reshape(-1))
        bvals = np.ascontiguousarray(np.asarray(bvals, dtype=np.float64).reshape(-1))
        if src.size != self.n1 * self.n2:
            raise ValueError("source term must have N1*N2 entries")
        self.Q = int(Q)
        d = _lib.ProblemDesc()
        d.dim, d.kernel_id, d.eq_type = dim, _lib.KERNEL_IDS[kernel_name], _lib.EQ_IDS[eq_name]
        d.n1, d.n2, d.Q, d.nb = self.n1, self.n2, self.Q, bvals.size
        d.force_general = int(force_general)       # bit 0: general Gram path, bit 1: no FFT diagonal sums
        d.llk_weight, d.logdet, d.beta, d.jitter = float(llk_weight), float(logdet), float(beta), float(jitter)
        self.desc = d
        if dim == 2 and bvals.size != 2 * self.n1 + 2 * self.n2:
            raise ValueError("bvals must hold the four edges (2*N1 + 2*N2 values)")
        xi = None
        if dim == 1:
            xi = np.ascontiguousarray(np.asarray(xind, dtype=np.int32).reshape(-1))
            if xi.size != bvals.size:
                raise ValueError("Xind and y must have the same length")
        nbytes = self.lib.gphm_workspace_bytes(ctypes.byref(d))
        if nbytes == 0:
            _lib.check(-1, "gphm_workspace_bytes")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        plan = ctypes.c_void_p()
        hp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(self.lib.gphm_plan_create(ctypes.byref(d), hp(x), hp(y), hp(src), hp(bvals), hp(xi),
                                             _lib.ptr(self.workspace), nbytes, ctypes.byref(plan)), "gphm_plan_create")
        self.plan = plan
        self._pred_work = {}
        self._l2_work = torch.empty(self.lib.gphm_rel_l2_work_bytes(), dtype=torch.uint8, device=self.device)

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.gphm_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # ---- pytree <-> packed -------------------------------------------------------------------
    @property
    def nf(self):
        return self.n1 * self.n2

    def _kp_keys(self):
        return ("kernel_paras_1", "kernel_paras_2") if self.dim == 2 else ("kernel_paras",)

    def _u_key(self):
        return "U" if self.dim == 2 else "u"

    def pack_tree(self, tree, U_out, small_out):
        """Write a params-shaped pytree (params, grads, mu, nu) into packed buffers."""
        Q = self.Q
        U_out.copy_(as_dev(tree[self._u_key()]).reshape(-1))
        small_out.zero_()
        for a, key in enumerate(self._kp_keys()):
            for j, leaf in enumerate(("log-w", "log-ls", "freq")):
                v = as_dev(tree[key][leaf]).reshape(-1)
                if v.numel() != Q:
                    raise ValueError("%s/%s must have Q=%d entries" % (key, leaf, Q))
                small_out[(3 * a + j) * Q:(3 * a + j + 1) * Q].copy_(v)
        small_out[6 * Q] = as_dev(tree["log_tau"]).reshape(())
        small_out[6 * Q + 1] = as_dev(tree["log_v"]).reshape(())

    def unpack_tree(self, U, small):
        Q = self.Q
        shape = (self.n1, self.n2) if self.dim == 2 else (self.n1, 1)
        out = {"log_tau": small[6 * Q].clone(), "log_v": small[6 * Q + 1].clone(), self._u_key(): U.reshape(shape).clone()}
        for a, key in enumerate(self._kp_keys()):
            out[key] = {leaf: small[(3 * a + j) * Q:(3 * a + j + 1) * Q].clone()
                        for j, leaf in enumerate(("log-w", "log-ls", "freq"))}
        return out

    def new_state(self, params=None, opt_state=None):
        st = PackedState(self.nf, self.Q, self.device)
        if params is not None:
            self.pack_tree(params, st.U, st.small)
        if opt_state is not None:
            self.pack_tree(opt_state["mu"], st.mU, st.msmall)
            self.pack_tree(opt_state["nu"], st.vU, st.vsmall)
            st.count.fill_(int(opt_state["count"]))
        return st

    def init_opt_state(self, params):
        """optimizer.init(params): zero first/second moments, count 0 (optax ScaleByAdamState)."""
        zeros = lambda t: {k: zeros(v) for k, v in t.items()} if isinstance(t, dict) else torch.zeros_like(as_dev(t))
        return {"count": torch.zeros((), dtype=torch.int64, device=self.device), "mu": zeros(params), "nu": zeros(params)}

    # ---- calls -------------------------------------------------------------------------------
    def value_and_grad(self, st, forward_only=False):
        """Returns (terms[8] device tensor, gU, gsmall) at the packed state."""
        gU = None if forward_only else torch.empty_like(st.U)
        gs = None if forward_only else torch.empty_like(st.small)
        terms = torch.empty(8, dtype=DT, device=self.device)
        _lib.check(self.lib.gphm_logjoint_grad(self.plan, _lib.ptr(st.U), _lib.ptr(st.small), _lib.ptr(gU), _lib.ptr(gs),
                                               _lib.ptr(terms), _lib.FORWARD_ONLY if forward_only else 0,
                                               _lib.stream_ptr()), "gphm_logjoint_grad")
        return terms, gU, gs

    def step_inplace(self, st, lr):
        """One fused step on a PackedState; st.terms gets the pre-update loss terms.  No host sync."""
        _lib.check(self.lib.gphm_step(self.plan, _lib.ptr(st.U), _lib.ptr(st.small), _lib.ptr(st.mU), _lib.ptr(st.vU),
                                      _lib.ptr(st.msmall), _lib.ptr(st.vsmall), _lib.ptr(st.count), float(lr),
                                      _lib.ptr(st.terms), _lib.stream_ptr()), "gphm_step")

    def step_host(self, hU, hsmall, hmU, hvU, hmsmall, hvsmall, hcount, hterms, lr):
        """Functional-style step on HOST (ideally pinned) buffers: H2D, step, D2H, synchronise."""
        _lib.check(self.lib.gphm_step_host(self.plan, _lib.ptr(hU), _lib.ptr(hsmall), _lib.ptr(hmU), _lib.ptr(hvU),
                                           _lib.ptr(hmsmall), _lib.ptr(hvsmall), _lib.ptr(hcount), float(lr),
                                           _lib.ptr(hterms), _lib.stream_ptr()), "gphm_step_host")

    def step_host_params(self, hU, hsmall, hterms, lr, reset_opt=False, hcount=None):
        """Step on HOST params with the Adam state resident in the plan (gphm_step_host_params)."""
        _lib.check(self.lib.gphm_step_host_params(self.plan, _lib.ptr(hU), _lib.ptr(hsmall), int(bool(reset_opt)),
                                                  _lib.ptr(hcount), float(lr), _lib.ptr(hterms), _lib.stream_ptr()),
                   "gphm_step_host_params")

    def status(self):
        piv = ctypes.c_int(0)
        rc = self.lib.gphm_plan_status(self.plan, ctypes.byref(piv), _lib.stream_ptr())
        if rc < 0:
            _lib.check(rc, "gphm_plan_status")
        return rc, piv.value

    def raise_on_bad_status(self):
        """Raises on a non-SPD Gram matrix / non-finite loss.  GPHM_ILL_CONDITIONED (the conditioning guard of the
        Toeplitz inverse-generator route fired) is not an error: the plan moves to the Cholesky route for every
        following call, with a warning."""
        rc, piv = self.status()
        if rc == _lib.ILL_CONDITIONED:
            import warnings
            _lib.check(self.lib.gphm_plan_use_cholesky(self.plan), "gphm_plan_use_cholesky")
            self.fell_back_to_cholesky = True
            warnings.warn("libgphm: Gram matrix too ill-conditioned for the Toeplitz inverse-generator route (axis mask %d): "
                          "continuing on the blocked Cholesky route" % piv, RuntimeWarning)
            return
        if rc == 4:
            raise _lib.GphmError("gphm_plan_status: " + _lib.last_error())
        if rc == _lib.NOT_SPD:
            raise FloatingPointError("Gram matrix is not positive definite (pivot %d)" % piv)
        if rc == _lib.NONFINITE:
            raise FloatingPointError("loss is not finite")

    def check_conditioning(self, st):
        """One forward-only evaluation + a status read (the only host sync): lets the conditioning guard of the Toeplitz
        inverse-generator route move the plan to the Cholesky route BEFORE a training loop starts."""
        self.value_and_grad(st, forward_only=True)
        self.raise_on_bad_status()

    def predict(self, st, xt, yt=None):
        xt = as_dev(xt).reshape(-1)
        m1 = xt.numel()
        if self.dim == 2:
            yt = as_dev(yt).reshape(-1)
            m2 = yt.numel()
        else:
            m2 = 1
        key = (m1, m2)
        if key not in self._pred_work:
            nbytes = self.lib.gphm_predict_work_bytes(self.plan, m1, m2)
            self._pred_work[key] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        out = torch.empty((m1, m2), dtype=DT, device=self.device)
        _lib.check(self.lib.gphm_predict(self.plan, _lib.ptr(st.U), _lib.ptr(st.small), _lib.ptr(xt), m1,
                                         _lib.ptr(yt) if self.dim == 2 else None, m2, _lib.ptr(out),
                                         _lib.ptr(self._pred_work[key]), _lib.stream_ptr()), "gphm_predict")
        return out

    def rel_l2(self, pred, truth):
        """Device scalar ||pred - truth|| / ||truth||."""
        out = torch.empty(1, dtype=DT, device=self.device)
        _lib.check(self.lib.gphm_rel_l2(_lib.ptr(pred), _lib.ptr(truth), pred.numel(), _lib.ptr(out),
                                        _lib.ptr(self._l2_work), _lib.stream_ptr()), "gphm_rel_l2")
        return out


# ---- primitive wrappers used by the secondary reference methods and by tests -------------------
def dgemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, C=None):
    lib = _lib.load()
    A, B = as_dev(A), as_dev(B)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    K2, N = (B.shape[1], B.shape[0]) if transB else B.shape
    if K != K2:
        raise ValueError("dgemm: inner dimensions differ")
    if C is None:
        C = torch.zeros((M, N), dtype=DT, device=A.device)
    _lib.check(lib.gphm_dgemm(int(transA), int(transB), M, N, K, float(alpha), _lib.ptr(A), max(A.shape[1], 1), _lib.ptr(B),
                              max(B.shape[1], 1), float(beta), _lib.ptr(C), max(C.shape[1], 1), _lib.stream_ptr()), "gphm_dgemm")
    return C


def ozaki_dgemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, C=None, slices=0):
    """alpha * op(A) op(B) + beta * C through the tcgen05 int8 Ozaki GEMM (gphm_ozaki_dgemm); returns (C, error factor):
    |C - exact|_ij <= |alpha| * factor * max_k|op(A)_ik| * max_k|op(B)_kj|."""
    lib = _lib.load()
    A, B = as_dev(A).contiguous(), as_dev(B).contiguous()
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    K2, N = (B.shape[1], B.shape[0]) if transB else B.shape
    if K != K2:
        raise ValueError("ozaki_dgemm: inner dimensions differ")
    if C is None:
        C = torch.zeros((M, N), dtype=DT, device=A.device)
    work = torch.empty(lib.gphm_ozaki_work_bytes(M, N, K, slices), dtype=torch.uint8, device=A.device)
    _lib.check(lib.gphm_ozaki_dgemm(int(transA), int(transB), M, N, K, float(alpha), _lib.ptr(A), max(A.shape[1], 1), _lib.ptr(B),
                                    max(B.shape[1], 1), float(beta), _lib.ptr(C), max(C.shape[1], 1), int(slices), _lib.ptr(work),
                                    work.numel(), _lib.stream_ptr()), "gphm_ozaki_dgemm")
    return C, float(lib.gphm_ozaki_error_factor(K, slices))


def potrf_inv(K):
    """(L, Linv, logdet, status) of an SPD matrix; K is not modified."""
    lib = _lib.load()
    Kc = as_dev(K).clone()
    n = Kc.shape[0]
    L, Linv = torch.empty_like(Kc), torch.empty_like(Kc)
    logdet = torch.empty(1, dtype=DT, device=Kc.device)
    status = torch.zeros(1, dtype=torch.int32, device=Kc.device)
    work = torch.empty(lib.gphm_potrf_work_bytes(n), dtype=torch.uint8, device=Kc.device)
    _lib.check(lib.gphm_potrf_inv(_lib.ptr(Kc), n, _lib.ptr(L), _lib.ptr(Linv), _lib.ptr(logdet), _lib.ptr(status),
                                  _lib.ptr(work), _lib.stream_ptr()), "gphm_potrf_inv")
    return L, Linv, logdet, status


def solve_spd(K, B):
    """K^-1 B through libgphm's Cholesky + L^-1 (the reference's jnp.linalg.solve call sites)."""
    _, Linv, _, _ = potrf_inv(K)
    return dgemm(Linv, dgemm(Linv, B), transA=True)


def toeplitz_solve(t, B=None):
    """(X, g, sKinv, logdet, status) for the SPD Toeplitz matrix with first column t: X[r] = K^-1 B[r]
    for every row of B (None: no right-hand sides), g = K^-1 e_0, sKinv = diagonal sums of K^-1."""
    lib = _lib.load()
    t = as_dev(t).reshape(-1).contiguous()
    n = t.numel()
    Bd = None if B is None else as_dev(B).reshape(-1, n).contiguous()
    rows = 0 if Bd is None else Bd.shape[0]
    X = None if Bd is None else torch.empty_like(Bd)
    g, sK = torch.empty_like(t), torch.empty_like(t)
    logdet = torch.empty(1, dtype=DT, device=t.device)
    status = torch.zeros(1, dtype=torch.int32, device=t.device)
    nbytes = lib.gphm_toeplitz_work_bytes(n, rows)
    if nbytes == 0:
        raise ValueError("toeplitz_solve: n=%d not supported (1 <= n <= 4096)" % n)
    work = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
    _lib.check(lib.gphm_toeplitz_solve(_lib.ptr(t), n, _lib.ptr(Bd), rows, _lib.ptr(X), _lib.ptr(g), _lib.ptr(sK),
                                       _lib.ptr(logdet), _lib.ptr(status), _lib.ptr(work), _lib.stream_ptr()),
               "gphm_toeplitz_solve")
    return X, g, sK, logdet, status
