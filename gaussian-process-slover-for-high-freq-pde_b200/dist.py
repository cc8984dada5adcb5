"""Multi-GPU (one process per GPU) version of the 2-D step: U, the source term and the Adam state
are sharded by row blocks; the per-axis operators (Gram factors) are factored once per axis by one
half of the ranks and broadcast.

Why it shards: left-multiplications by axis-1 operators (K1^-1, D1) act on every
column of the N1 x N2 field independently, right-multiplications by axis-2 operators (K2^-1, D2)
on every row.  So axis-2 work runs on the resident row blocks ("R" layout, h x N2, h = N1/P),
axis-1 work on column blocks ("C" layout, N1 x w, w = N2/P), and the only data-path collective is
the block transpose between the two layouts (NCCL all-to-all, N1*N2*8/P bytes per rank), seven
times per step.  Scalars (loss terms, 6Q theta-gradients) use one small all-reduce each.

    R->C : U, G, Bt            C->R : c1*D1*A, A, W, S1          (general grids: 7 exchanges)
    R->Ct: [U], [G, Bt]        Ct->R: [c1*D1*A, A], [V1]         (uniform grids, all-FFT step: 4 exchanges)

The numerical pieces are calls into libgphm through the `ops` object (CudaOps below).  The step
logic itself is backend-agnostic so that tests can drive it with a CPU stand-in over gloo.
"""
import math

import torch
import torch.distributed as dist

from . import _lib

DT = torch.float64


# ------------------------------------------------------------------------------------------------
class CudaOps(object):
    """libgphm-backed primitives on the current CUDA device (the production backend)."""

    def __init__(self, core):
        self.core = core
        self.lib = core.lib
        self.plan = core.plan
        self.device = core.device
        self.Q = core.Q
        self._cache = {}
        import os
        self.use_comm_stream = os.environ.get("GPHM_MG_COMM_STREAM", "1") != "0"

    def _buf(self, tag, shape):
        key = (tag,) + tuple(shape)
        if key not in self._cache:
            self._cache[key] = torch.empty(shape, dtype=DT, device=self.device)
        return self._cache[key]

    def zeros(self, shape, dtype=DT):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def tensor(self, data, dtype=DT):
        return torch.as_tensor(data, dtype=dtype).to(self.device).contiguous()

    def _s(self):
        return _lib.stream_ptr()

    def factor(self, small, axis_mask=3):
        _lib.check(self.lib.gphm_plan_factor(self.plan, _lib.ptr(small), axis_mask, self._s()), "gphm_plan_factor")

    def mat(self, axis, which):
        """(n, n) view of a plan-owned matrix: which = 0 K^-1, 1 D, 2 Linv, 3 L."""
        p = self.lib.gphm_plan_matrix(self.plan, axis, which)
        if not p:
            raise _lib.GphmError("gphm_plan_matrix returned NULL")
        n = self.core.n1 if axis == 0 else self.core.n2
        off = p - self.core.workspace.data_ptr()
        return self.core.workspace[off:off + n * n * 8].view(DT).view(n, n)

    def logdets(self):
        out = self._buf("logdets", (2,))
        _lib.check(self.lib.gphm_plan_logdet(self.plan, _lib.ptr(out), self._s()), "gphm_plan_logdet")
        return out

    def apply_kinv(self, axis, side, X, tag):
        out, tmp = self._buf(tag, X.shape), self._buf("kinv_tmp", X.shape)
        _lib.check(self.lib.gphm_apply_kinv(self.plan, axis, side, _lib.ptr(X), X.shape[0], X.shape[1], _lib.ptr(out),
                                            _lib.ptr(tmp), self._s()), "gphm_apply_kinv")
        return out

    def gemm(self, A, B, tA, tB, alpha, beta, C):
        M, K = (A.shape[1], A.shape[0]) if tA else A.shape
        N = B.shape[0] if tB else B.shape[1]
        _lib.check(self.lib.gphm_dgemm(int(tA), int(tB), M, N, K, float(alpha), _lib.ptr(A), A.stride(0), _lib.ptr(B),
                                       B.stride(0), float(beta), _lib.ptr(C), C.stride(0), self._s()), "gphm_dgemm")
        return C

    def new(self, tag, shape):
        return self._buf(tag, shape)

    def residual(self, R, U, F, A, Bt, small, out=None):
        out = self._buf("res2", (2,)) if out is None else out
        _lib.check(self.lib.gphm_mg_residual(self.plan, _lib.ptr(R), _lib.ptr(U), _lib.ptr(F), _lib.ptr(A), _lib.ptr(Bt),
                                             R.numel(), _lib.ptr(small), _lib.ptr(out), self._s()), "gphm_mg_residual")
        return out

    def boundary(self, U, bidx, bvals, out=None):
        eb, out = self._buf("eb", (max(bidx.numel(), 1),)), (self._buf("bg1", (1,)) if out is None else out)
        _lib.check(self.lib.gphm_mg_boundary(_lib.ptr(U), _lib.ptr(bidx), _lib.ptr(bvals), bidx.numel(), _lib.ptr(eb),
                                             _lib.ptr(out), self._s()), "gphm_mg_boundary")
        return eb, out

    def grad_u(self, U, G, W, S1, S2, bidx, eb, nseg0, small):
        gU, V2 = self._buf("gU", U.shape), self._buf("V2", U.shape)
        _lib.check(self.lib.gphm_mg_grad_u(self.plan, _lib.ptr(U), _lib.ptr(G), _lib.ptr(W), _lib.ptr(S1), _lib.ptr(S2),
                                           U.numel(), _lib.ptr(bidx), _lib.ptr(eb), nseg0, bidx.numel(), _lib.ptr(small),
                                           _lib.ptr(gU), _lib.ptr(V2), self._s()), "gphm_mg_grad_u")
        return gU, V2

    def lincomb(self, a, x, b, y, tag):
        out = self._buf(tag, x.shape)
        _lib.check(self.lib.gphm_lincomb(_lib.ptr(out), float(a), _lib.ptr(x), float(b), _lib.ptr(y), x.numel(), self._s()),
                   "gphm_lincomb")
        return out

    def theta_grad(self, axis, Kbar, Dbar, small, out):
        _lib.check(self.lib.gphm_mg_theta_grad(self.plan, axis, _lib.ptr(Kbar), _lib.ptr(Dbar), _lib.ptr(small),
                                               _lib.ptr(out), self._s()), "gphm_mg_theta_grad")

    def uses_fft(self, axis):
        return bool(self.lib.gphm_plan_uses_fft(self.plan, axis))

    def uses_gs(self, axis):
        return bool(self.lib.gphm_plan_uses_gs(self.plan, axis))

    def transpose(self, X, tag):
        out = self._buf(tag, (X.shape[1], X.shape[0]))
        _lib.check(self.lib.gphm_transpose(_lib.ptr(X), X.shape[0], X.shape[1], _lib.ptr(out), self._s()), "gphm_transpose")
        return out

    def toeplitz_rows(self, axis, transposed, X, alpha, beta, small, out):
        """out[r] = alpha * D x_r (or D^T x_r) + beta * out[r] for every row of X, D the axis' Toeplitz derivative Gram."""
        _lib.check(self.lib.gphm_mg_toeplitz_apply(self.plan, axis, int(transposed), _lib.ptr(X), X.shape[0], float(alpha),
                                                   float(beta), _lib.ptr(small), _lib.ptr(out), self._s()),
                   "gphm_mg_toeplitz_apply")
        return out

    def theta_grad_rows(self, axis, X, Y, G, r0, r1, beta, cD, small, out):
        """theta-gradient of  beta*Linv[r0:r1]^T Linv[r0:r1] - X^T Y  and  cD*G^T Y  via FFT diagonal sums."""
        _lib.check(self.lib.gphm_mg_theta_grad_fft(self.plan, axis, _lib.ptr(X), _lib.ptr(Y), _lib.ptr(G), X.shape[0], r0, r1,
                                                   float(beta), float(cD), _lib.ptr(small), _lib.ptr(out), self._s()),
                   "gphm_mg_theta_grad_fft")

    def kinv_rows(self, axis, X, tag, refine=False):
        """Every row of X (rows x n_axis) times K_axis^-1; refine: one step of iterative refinement (the reverse-pass
        applications V1, V2 - gphm_apply_kinv_rows_refined)."""
        out, tmp = self._buf(tag, X.shape), self._buf("kinv_tmp", X.shape)
        if refine:
            _lib.check(self.lib.gphm_apply_kinv_rows_refined(self.plan, axis, _lib.ptr(X), X.shape[0], _lib.ptr(out),
                                                             _lib.ptr(tmp), self._s()), "gphm_apply_kinv_rows_refined")
        else:
            _lib.check(self.lib.gphm_apply_kinv(self.plan, axis, 1, _lib.ptr(X), X.shape[0], X.shape[1], _lib.ptr(out),
                                                _lib.ptr(tmp), self._s()), "gphm_apply_kinv")
        return out

    def toeplitz_rows_add(self, axis, transposed, X, alpha, beta, add, out, keep):
        """out[r] = alpha * D x_r (D^T x_r) + beta * add[r]; keep: the plan stores the transforms of X's rows."""
        _lib.check(self.lib.gphm_mg_toeplitz_rows(self.plan, axis, int(transposed), _lib.ptr(X), X.shape[0], float(alpha),
                                                  float(beta), _lib.ptr(add), _lib.ptr(out), int(keep), self._s()),
                   "gphm_mg_toeplitz_rows")
        return out

    def theta_grad_pairs(self, axis, V, G, lead, beta, cD, small, out):
        _lib.check(self.lib.gphm_mg_theta_grad_pairs(self.plan, axis, _lib.ptr(V), _lib.ptr(G), V.shape[0], int(lead),
                                                     float(beta), float(cD), _lib.ptr(small), _lib.ptr(out), self._s()),
                   "gphm_mg_theta_grad_pairs")

    def theta_grad_pairs_both(self, V1, G1, V2, G2, lead, beta1, beta2, c1, c2, small, out):
        _lib.check(self.lib.gphm_mg_theta_grad_pairs_both(self.plan, _lib.ptr(V1), _lib.ptr(G1), V1.shape[0], _lib.ptr(V2),
                                                          _lib.ptr(G2), V2.shape[0], int(lead), float(beta1), float(beta2),
                                                          float(c1), float(c2), _lib.ptr(small), _lib.ptr(out), self._s()),
                   "gphm_mg_theta_grad_pairs_both")
        return out

    def grad_u_sum(self, U, G, V1, V2, bidx, eb, nseg0, small):
        gU = self._buf("gU", U.shape)
        _lib.check(self.lib.gphm_mg_grad_u(self.plan, _lib.ptr(U), _lib.ptr(G), _lib.ptr(V1), _lib.ptr(V2), None,
                                           U.numel(), _lib.ptr(bidx), _lib.ptr(eb), nseg0, bidx.numel(), _lib.ptr(small),
                                           _lib.ptr(gU), None, self._s()), "gphm_mg_grad_u")
        return gU

    def pack_transposed(self, Xs, part_cols, tag="pack"):
        """send[d][a] = (X_a^T)[d * part_cols : (d+1) * part_cols]  for the arrays X_a (rows x cols): one tiled transpose
        per array, written straight into the (persistent) all-to-all send buffer."""
        k, (rows, cols) = len(Xs), Xs[0].shape
        parts = cols // part_cols
        send = self._buf(tag + ".send", (parts, k, part_cols, rows))
        blk = part_cols * rows
        for a, X in enumerate(Xs):
            dst = send.view(-1)[a * blk:]
            _lib.check(self.lib.gphm_mg_pack_transposed(_lib.ptr(X), rows, cols, part_cols, k * blk, _lib.ptr(dst), self._s()),
                       "gphm_mg_pack_transposed")
        return send

    def unpack_segments(self, recv, tag="pack"):
        """recv[s][a][r][c] -> out[a][r][s * seg + c]  (persistent output buffer)."""
        parts, k, rows, seg = recv.shape
        out = self._buf(tag + ".out", (k, rows, parts * seg))
        _lib.check(self.lib.gphm_mg_unpack_segments(_lib.ptr(recv), parts, k, rows, seg, _lib.ptr(out), self._s()),
                   "gphm_mg_unpack_segments")
        return out

    def peer_exchange(self, Xs, part_cols, tag, group):
        """Transposing all-to-all of the arrays Xs (rows x cols each) through peer stores; returns (k, part_cols, P*rows)."""
        k, (rows, cols) = len(Xs), Xs[0].shape
        key = ("peer", tag, k, rows, cols)
        ex = self._cache.get(key)
        if ex is None:
            ex = self._cache[key] = PeerExchange(self, k, rows, cols, group)
        return ex.run([X.contiguous() for X in Xs], part_cols)

    def finalize(self, sums3, ld2, small, terms, gsmall):
        """terms[8] and gsmall[6Q:6Q+2] from the all-reduced [eq_gap, quad, boundary_gap] and the log-dets (one kernel)."""
        _lib.check(self.lib.gphm_mg_finalize(self.plan, _lib.ptr(sums3), _lib.ptr(ld2), _lib.ptr(small), _lib.ptr(terms),
                                             _lib.ptr(gsmall), self._s()), "gphm_mg_finalize")

    # ---- communication stream: an exchange (packing copies + all-to-all) issued with fork() runs beside the
    # kernels the main stream launches until join() ----
    def fork(self, fn):
        if not self.use_comm_stream:
            return fn(), None
        if getattr(self, "_comm", None) is None:
            self._comm = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self._comm.wait_stream(cur)                       # everything issued so far is visible to the exchange
        with torch.cuda.stream(self._comm):
            res = fn()
            ev = torch.cuda.Event()
            ev.record(self._comm)
        return res, ev

    def join(self, handle):
        res, ev = handle
        if ev is not None:
            # the exchange buffers are persistent (one set per exchange of the step), so stream order is all that is
            # needed: no allocator hand-over (record_stream) and no allocation inside the step
            torch.cuda.current_stream(self.device).wait_event(ev)
        return res

    def adam(self, p, g, m, v, count, lr):
        _lib.check(self.lib.gphm_adam_update(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), p.numel(), _lib.ptr(count),
                                             float(lr), self._s()), "gphm_adam_update")

    def adam_inc(self, p, g, m, v, count, lr):
        _lib.check(self.lib.gphm_adam_update_inc(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), p.numel(), _lib.ptr(count),
                                                 float(lr), self._s()), "gphm_adam_update_inc")


# ------------------------------------------------------------------------------------------------
class PeerExchange(object):
    """One exchange of the sharded step through NVLink peer memory (csrc/peer.cu): a buffer in cudaIpc-shared device
    memory on every rank, the peers' mappings of it, and a sequence counter.  `run(Xs, part_cols)` is collective:
    it stores this rank's tiles straight into every destination's buffer (transposed, final layout) and enqueues the
    wait for all sources; the returned views alias the local buffer (valid until the next run of THIS exchange)."""

    def __init__(self, ops, k, rows, cols, group):
        import ctypes
        self.ops, self.lib, self.group = ops, ops.lib, group
        self.P, self.me = dist.get_world_size(group), dist.get_rank(group)
        self.k, self.rows, self.cols = k, rows, cols
        self.data_doubles = k * rows * cols
        base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.gphm_mg_peer_alloc(self.data_doubles, ctypes.byref(base), handle), "gphm_mg_peer_alloc")
        self.base = base.value
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=ops.device)
        allh = [torch.empty_like(mine) for _ in range(self.P)]
        dist.all_gather(allh, mine, group=group)
        self.bases = (ctypes.c_void_p * self.P)()
        self._opened = []
        for r in range(self.P):
            if r == self.me:
                self.bases[r] = self.base
                continue
            hb = (ctypes.c_ubyte * 64)(*allh[r].cpu().tolist())
            pb = ctypes.c_void_p()
            _lib.check(self.lib.gphm_mg_peer_open(hb, ctypes.byref(pb)), "gphm_mg_peer_open")
            self.bases[r] = pb.value
            self._opened.append(pb.value)
        self.seq = 0
        self.status = torch.zeros(1, dtype=torch.int32, device=ops.device)
        self._in = (ctypes.c_void_p * k)()
        dist.barrier(group=group)                      # every rank has mapped every buffer before the first store

    def view(self, pc):
        """(k, pc, P * rows) float64 view of the local result."""
        return _tensor_from_ptr(self.base + 512, (self.k, pc, self.P * self.rows), self.ops.device)

    def run(self, Xs, part_cols):
        self.seq += 1
        for a, X in enumerate(Xs):
            self._in[a] = X.data_ptr()
        _lib.check(self.lib.gphm_mg_peer_exchange(self._in, self.k, self.rows, self.cols, part_cols, self.bases, self.P, self.me,
                                                  self.seq, self.data_doubles, _lib.ptr(self.status), _lib.stream_ptr()),
                   "gphm_mg_peer_exchange")
        return self.view(part_cols)

    def close(self):
        for pb in self._opened:
            self.lib.gphm_mg_peer_close(pb)
        self._opened = []
        if self.base:
            self.lib.gphm_mg_peer_free(self.base)
            self.base = None


def _tensor_from_ptr(ptr, shape, device):
    """float64 CUDA tensor over raw device memory owned by libgphm (__cuda_array_interface__)."""
    class _Raw(object):
        pass
    raw = _Raw()
    n = 1
    for d in shape:
        n *= d
    raw.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(raw, device=device).view(*shape)


# ------------------------------------------------------------------------------------------------
def local_boundary(rank, P, N1, N2, bvals):
    """Rank-local boundary points of the row block [rank*h, (rank+1)*h): flat indices into the
    (h, N2) block and target values, in two segments with unique indices each (row edges first,
    then column edges) so that corner contributions are added in a fixed order.
    bvals order: U[0,:], U[-1,:], U[:,0], U[:,-1] (model_GP_solver_2d.py:127)."""
    h = N1 // P
    idx, val = [], []
    if rank == 0:
        idx += list(range(N2)); val += list(bvals[0:N2])
    if rank == P - 1:
        idx += [(h - 1) * N2 + j for j in range(N2)]; val += list(bvals[N2:2 * N2])
    nseg0 = len(idx)
    r0 = rank * h
    idx += [i * N2 for i in range(h)]; val += list(bvals[2 * N2 + r0:2 * N2 + r0 + h])
    idx += [i * N2 + N2 - 1 for i in range(h)]; val += list(bvals[2 * N2 + N1 + r0:2 * N2 + N1 + r0 + h])
    return idx, val, nseg0


class ShardedSolver2D(object):
    """Row-sharded GP_solver_2d_single step (Poisson / Allen-Cahn / advection).  All ranks must
    construct it with identical arguments; `step()` is collective."""

    def __init__(self, kernel_name, eq_name, x, y, src, bvals, llk_weight, logdet, beta, jitter, Q, lr, ops=None,
                 group=None, force_general=0):
        import numpy as np
        self.group = group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        x, y = np.asarray(x, dtype=np.float64).reshape(-1), np.asarray(y, dtype=np.float64).reshape(-1)
        self.N1, self.N2, self.Q = x.size, y.size, int(Q)
        if self.N1 % self.P or self.N2 % self.P:
            raise ValueError("sharded step needs N1 and N2 divisible by the number of ranks (%d x %d, P=%d)"
                             % (self.N1, self.N2, self.P))
        self.h, self.w = self.N1 // self.P, self.N2 // self.P
        src = np.asarray(src, dtype=np.float64).reshape(self.N1, self.N2)
        bvals = np.asarray(bvals, dtype=np.float64).reshape(-1)
        self.eq_name, self.llk_weight, self.logdet, self.lr = eq_name, float(llk_weight), float(logdet), float(lr)
        self.c1 = float(beta) if eq_name == "advection" else 1.0
        if ops is None:
            from .solver_core import SolverCore
            core = SolverCore(2, kernel_name, eq_name, x, y, src, bvals, None, llk_weight, logdet, beta, jitter, Q,
                              force_general=force_general)
            ops = CudaOps(core)
        self.ops = ops
        r0 = self.rank * self.h
        self.F = ops.tensor(src[r0:r0 + self.h, :])
        idx, val, self.nseg0 = local_boundary(self.rank, self.P, self.N1, self.N2, bvals)
        self.bidx = ops.tensor(idx, dtype=torch.int32)
        self.bvals = ops.tensor(val)
        self.Nb, self.Nc = 2 * self.N1 + 2 * self.N2, self.N1 * self.N2
        ns = 6 * self.Q + 2
        self.U = ops.zeros((self.h, self.N2))
        self.mU, self.vU = ops.zeros((self.h, self.N2)), ops.zeros((self.h, self.N2))
        self.small, self.msmall, self.vsmall = ops.zeros((ns,)), ops.zeros((ns,)), ops.zeros((ns,))
        self.gsmall = ops.zeros((ns,))
        self.count = ops.zeros((1,), dtype=torch.int64)
        self.terms = ops.zeros((8,))
        self.acc = ops.zeros((3 + ns,))
        self.bytes_exchanged = 0
        import os
        # exchange: "peer" = one kernel with NVLink peer stores (csrc/peer.cu; default on CUDA with P > 1), "nccl" = pack ->
        # all_to_all_single -> unpack (round 1; also what the gloo CPU tests drive)
        mode = os.environ.get("GPHM_MG_EXCHANGE", "peer")
        self.exchange = "peer" if (mode == "peer" and self.P > 1 and hasattr(self.ops, "peer_exchange")) else "nccl"
        # hooks of step_host into the all-FFT step (None otherwise)
        self._u_ready, self._on_gu, self._copy = None, None, None

    # ---- state ---------------------------------------------------------------------------------
    def init_state(self, freq_scale):
        """Reference initial state (model_GP_solver_2d.py:245-261), identical on every rank."""
        Q = self.Q
        s = torch.zeros(6 * Q + 2, dtype=DT)
        for a in range(2):
            s[(3 * a) * Q:(3 * a + 1) * Q] = math.log(1.0 / Q)
            s[(3 * a + 2) * Q:(3 * a + 3) * Q] = torch.linspace(0, 1, Q, dtype=DT) * freq_scale
        self.set_state(torch.zeros(self.N1, self.N2, dtype=DT), s)

    def set_state(self, U_full, small):
        r0 = self.rank * self.h
        self.U.copy_(torch.as_tensor(U_full, dtype=DT)[r0:r0 + self.h, :])
        self.small.copy_(torch.as_tensor(small, dtype=DT))
        for t in (self.mU, self.vU, self.msmall, self.vsmall):
            t.zero_()
        self.count.zero_()

    def gather_U(self):
        """Full (N1, N2) U on every rank (for checks / prediction)."""
        if self.P == 1:
            return self.U.clone()
        parts = [torch.empty_like(self.U) for _ in range(self.P)]
        dist.all_gather(parts, self.U.contiguous(), group=self.group)
        return torch.cat(parts, 0)

    def last_loss(self):
        return self.terms[0]

    def exchange_name(self):
        return ("one transposing kernel per exchange with NVLink peer stores (cudaIpc), flags for completion; NCCL only for the "
                "3+6Q all-reduce") if self.exchange == "peer" else "NCCL all-to-all"

    # ---- layout exchanges ------------------------------------------------------------------------
    def _a2a(self, send, tag=None):
        recv = torch.empty_like(send) if tag is None else self.ops.new(tag + ".recv", tuple(send.shape))
        if self.P == 1:
            recv.copy_(send)
        else:
            dist.all_to_all_single(recv, send, group=self.group)
            self.bytes_exchanged += send.numel() * 8 * (self.P - 1) // self.P
        return recv

    def r2c(self, X):
        """(h, N2) row block -> (N1, w) column block."""
        P, h, w = self.P, self.h, self.w
        send = X.reshape(h, P, w).transpose(0, 1).contiguous()          # chunk s = my rows, columns of rank s
        return self._a2a(send).reshape(self.N1, w)                      # chunk s = rows of rank s, my columns

    def c2r(self, X):
        """(N1, w) column block -> (h, N2) row block."""
        P, h, w = self.P, self.h, self.w
        recv = self._a2a(X.contiguous().reshape(P, h, w))               # chunk s = my rows, columns of rank s
        return recv.transpose(0, 1).reshape(h, self.N2).contiguous()

    def _pack_t(self, Xs, part_cols, tag):
        """[dest][array][c][r] = X_array[r][dest * part_cols + c]  (backend kernel, else torch)."""
        f = getattr(self.ops, "pack_transposed", None)
        if f is not None:
            return f([X.contiguous() for X in Xs], part_cols, tag)
        k, (rows, cols) = len(Xs), Xs[0].shape
        return torch.stack(Xs).reshape(k, rows, cols // part_cols, part_cols).permute(2, 0, 3, 1).contiguous()

    def _unpack(self, recv, tag):
        """[src][array][r][c] -> [array][r][src * seg + c]."""
        f = getattr(self.ops, "unpack_segments", None)
        if f is not None:
            return f(recv, tag)
        parts, k, rows, seg = recv.shape
        return recv.permute(1, 2, 0, 3).reshape(k, rows, parts * seg).contiguous()

    def r2ct(self, Xs, tag="r2ct"):
        """Row blocks (h, N2) -> TRANSPOSED column blocks (w, N1): row j holds column rank*w + j of the field.
        Several arrays travel in one all-to-all.  `tag` names the persistent buffer set of this exchange."""
        tag = "%s%d" % (tag, len(Xs))
        if self.exchange == "peer":
            out = self.ops.peer_exchange(Xs, self.w, tag, self.group)
            self.bytes_exchanged += len(Xs) * Xs[0].numel() * 8 * (self.P - 1) // self.P
            return [out[a] for a in range(len(Xs))]
        send = self._pack_t(Xs, self.w, tag)                              # [dest][array][j][i]
        out = self._unpack(self._a2a(send, tag), tag)                     # recv [src][array][j][i]
        return [out[a] for a in range(len(Xs))]

    def ct2r(self, Ys, tag="ct2r"):
        """Transposed column blocks (w, N1) -> row blocks (h, N2)."""
        tag = "%s%d" % (tag, len(Ys))
        if self.exchange == "peer":
            out = self.ops.peer_exchange(Ys, self.h, tag, self.group)
            self.bytes_exchanged += len(Ys) * Ys[0].numel() * 8 * (self.P - 1) // self.P
            return [out[a] for a in range(len(Ys))]
        send = self._pack_t(Ys, self.h, tag)                              # [dest][array][i][j]
        out = self._unpack(self._a2a(send, tag), tag)                     # recv [src][array][i][j]
        return [out[a] for a in range(len(Ys))]

    def _allreduce(self, t):
        if self.P > 1:
            dist.all_reduce(t, group=self.group)
        return t

    # ---- factorisation: the two axes are independent, so rank halves take one axis each ----------
    def _factor(self, small, fft1, fft2):
        """Gram + Cholesky + L^-1 (+ K^-1) of both axes on every rank; returns [log|K1|, log|K2|].
        With P >= 2 ranks [0, P/2) factor axis 1 and ranks [P/2, P) axis 2, then each axis' D, Linv
        (and K^-1 when needed) is broadcast from the first rank of its half."""
        o = self.ops
        skip_kinv = fft1 and fft2
        skip = 4 if skip_kinv else 0
        gs = getattr(o, "uses_gs", None)
        if self.P == 1 or (gs is not None and gs(0) and gs(1)):
            # uniform grids: the Toeplitz inverse generators cost O(n^2) - every rank computes both, no broadcast
            o.factor(small, 3 | skip)
            return o.logdets()
        half = self.P // 2
        mine = 0 if self.rank < half else 1
        o.factor(small, (1 << mine) | skip)
        ld = o.logdets().clone()
        for axis, root in ((0, 0), (1, half)):
            src = dist.get_global_rank(self.group, root) if self.group is not None else root
            fft_axis = fft1 if axis == 0 else fft2       # FFT axes need neither D (Toeplitz table instead) nor K^-1
            which_list = [2] + ([] if fft_axis else [1]) + ([] if skip_kinv else [0])
            for which in which_list:
                dist.broadcast(o.mat(axis, which), src=src, group=self.group)
            dist.broadcast(ld[axis:axis + 1], src=src, group=self.group)
        return ld

    # ---- one iteration ---------------------------------------------------------------------------
    def value_and_grad_fft(self):
        """All-FFT step (both axes on the Toeplitz inverse generator): four K^-1 applications
        (V1 = K1^-1 (c1 D1^T G + Bt/2), V2 = (G D2 + A/2) K2^-1, dU = V1 + V2 + ...), axis-1 operands
        kept as transposed column blocks, four all-to-alls:  [U] R->Ct, [c1 D1 A, A] Ct->R, [G, Bt] R->Ct, [V1] Ct->R.
        The first exchange runs on the communication stream beside the (serial, 4-CTA) Schur recursion, which
        needs only theta.  (Keeping the later exchanges in flight behind axis-2 work was measured and rejected:
        an NCCL kernel and a one-CTA-per-SM FFT kernel evict each other, 10.7 vs 4.8 ms per step at P = 2.)
        The loss sums do not feed the reverse pass, so they share ONE all-reduce with the theta-gradients."""
        o, Q, c1 = self.ops, self.Q, self.c1
        N1, N2 = self.N1, self.N2
        small, U_r = self.small, self.U
        fork = getattr(o, "fork", None) or (lambda fn: (fn(), None))
        join = getattr(o, "join", None) or (lambda h: h[0])
        def first_exchange():
            if self._u_ready is not None:           # step_host: this rank's rows of U are still on their way up
                torch.cuda.current_stream(U_r.device).wait_event(self._u_ready)
            return self.r2ct([U_r])
        hU = fork(first_exchange)
        o.factor(small, 3)                          # O(n^2) generators + spectra, local to every rank
        ld = o.logdets()
        (U_ct,) = join(hU)
        At = o.kinv_rows(0, U_ct, "At")                                           # (K1^-1 U)^T      (Ct)
        Rt = o.toeplitz_rows_add(0, False, At, c1, 0.0, None, o.new("Rt", At.shape), True)      # (c1 D1 A)^T
        R_r, A_r = self.ct2r([Rt, At])
        Bt_r = o.kinv_rows(1, U_r, "Bt_r")                                        # U K2^-1          (R)
        eb, _ = o.boundary(U_r, self.bidx, self.bvals, out=self.acc[2:3])
        o.toeplitz_rows_add(1, False, Bt_r, 1.0, 1.0, R_r, R_r, True)             # + Bt D2^T
        acc = self.acc                               # [eqgap, quad, bgap | 6Q theta-gradients | d/dlog_tau, d/dlog_v]
        o.residual(R_r, U_r, self.F, A_r, Bt_r, small, out=acc[0:2])              # R_r <- G_r ; [eqgap, quad] (no torch op in the step)
        G_r = R_r
        gs = acc[3:]
        lead = self.rank == 0                        # the K^-1 (log-det) term is added once
        # backward
        G_ct, Btt = self.r2ct([G_r, Bt_r])
        T0 = o.toeplitz_rows_add(0, True, G_ct, c1, 0.5, Btt, o.new("T0", G_ct.shape), False)   # (c1 D1^T G + Bt/2)^T
        V1t = o.kinv_rows(0, T0, "V1t", refine=True)
        (V1_r,) = self.ct2r([V1t])
        P2 = o.toeplitz_rows_add(1, True, G_r, 1.0, 0.5, A_r, o.new("P2", G_r.shape), False)    # G D2 + A/2
        V2_r = o.kinv_rows(1, P2, "V2_r", refine=True)
        gU_r = o.grad_u_sum(U_r, G_r, V1_r, V2_r, self.bidx, eb, self.nseg0, small)
        if self._on_gu is not None:                  # step_host: nothing below reads U or dL/dU
            self._on_gu(gU_r)
        both = getattr(o, "theta_grad_pairs_both", None)       # (the CPU stand-in of the gloo tests has the per-axis call only)
        if both is not None:
            both(V1t, G_ct, V2_r, G_r, lead, 0.5 * self.logdet * N2, 0.5 * self.logdet * N1, c1, 1.0, small, gs[0:6 * Q])
        else:
            o.theta_grad_pairs(0, V1t, G_ct, lead, 0.5 * self.logdet * N2, c1, small, gs[0:3 * Q])
            o.theta_grad_pairs(1, V2_r, G_r, lead, 0.5 * self.logdet * N1, 1.0, small, gs[3 * Q:6 * Q])
        self._allreduce(acc[0:3 + 6 * Q])            # every entry of the slice was written above (nothing to zero)
        o.finalize(acc[0:3], ld, small, self.terms, gs)                           # terms[8]; gs[6Q], gs[6Q+1]
        return self.terms, gU_r, gs

    def value_and_grad(self):
        """Collective.  Returns (terms[8], gU_r (h,N2), gsmall (6Q+2)) - same layout as gphm_logjoint_grad."""
        gs_fn = getattr(self.ops, "uses_gs", None)
        if gs_fn is not None and gs_fn(0) and gs_fn(1):
            return self.value_and_grad_fft()
        o, Q, c1 = self.ops, self.Q, self.c1
        N1, N2, h, w = self.N1, self.N2, self.h, self.w
        small, U_r = self.small, self.U
        fft1, fft2 = o.uses_fft(0), o.uses_fft(1)
        ld = self._factor(small, fft1, fft2)
        D1 = None if fft1 else o.mat(0, 1)
        D2 = None if fft2 else o.mat(1, 1)
        # forward
        Bt_r = o.apply_kinv(1, 1, U_r, "Bt_r")                           # U K2^-1            (R)
        U_c = self.r2c(U_r)
        A_c = o.apply_kinv(0, 0, U_c, "A_c")                             # K1^-1 U            (C)
        if fft1:      # Toeplitz D1: FFT convolution of this rank's columns
            At = o.transpose(A_c, "At")
            Uxx_c = o.transpose(o.toeplitz_rows(0, False, At, c1, 0.0, small, o.new("Uxx_t", (w, N1))), "Uxx_c")
        else:
            At = None
            Uxx_c = o.gemm(D1, A_c, False, False, c1, 0.0, o.new("Uxx_c", (N1, w)))
        R_r = self.c2r(Uxx_c)
        A_r = self.c2r(A_c)
        if fft2:
            o.toeplitz_rows(1, False, Bt_r, 1.0, 1.0, small, R_r)               # + Bt D2^T          (R)
        else:
            o.gemm(Bt_r, D2, False, True, 1.0, 1.0, R_r)
        red = torch.empty(3, dtype=DT, device=R_r.device)
        red[0:2] = o.residual(R_r, U_r, self.F, A_r, Bt_r, small)        # R_r <- G_r ; [eqgap, quad]
        G_r = R_r
        eb, bg = o.boundary(U_r, self.bidx, self.bvals)
        red[2:3] = bg
        self._allreduce(red)
        eq, quad, bgap = red[0], red[1], red[2]
        tau, v = small[6 * Q], small[6 * Q + 1]
        loss = (0.5 * self.logdet * (N2 * ld[0] + N1 * ld[1]) + 0.5 * quad
                - self.llk_weight * (0.5 * self.Nb * tau - 0.5 * torch.exp(tau) * bgap)
                - (0.5 * self.Nc * v - 0.5 * torch.exp(v) * eq))
        gtau = -self.llk_weight * (0.5 * self.Nb - 0.5 * torch.exp(tau) * bgap)
        gv = -(0.5 * self.Nc - 0.5 * torch.exp(v) * eq)
        self.terms.copy_(torch.stack((loss, ld[0], ld[1], quad, bgap, eq, gtau, gv)))
        # backward, axis-1 work in C layout
        G_c = self.r2c(G_r)
        Bt_c = self.r2c(Bt_r)
        W_c = o.apply_kinv(0, 0, Bt_c, "W_c")
        if fft1:
            Gt = o.transpose(G_c, "Gt")
            P_c = o.transpose(o.toeplitz_rows(0, True, Gt, c1, 0.0, small, o.new("P_t", (w, N1))), "P_c")
        else:
            Gt = None
            P_c = o.gemm(D1, G_c, True, False, c1, 0.0, o.new("P_c", (N1, w)))
        S1_c = o.apply_kinv(0, 0, P_c, "S1_c")
        V1_c = o.lincomb(1.0, S1_c, 0.5, W_c, "V1_c")
        lead = 1.0 if self.rank == 0 else 0.0                            # the K^-1 (log-det) term is added once
        gs = self.gsmall
        gs.zero_()
        if fft1:      # uniform grid: diagonal sums by FFT over this rank's columns; Linv rows are split over ranks
            r0, r1 = self.rank * N1 // self.P, (self.rank + 1) * N1 // self.P
            o.theta_grad_rows(0, o.transpose(V1_c, "V1t"), At, Gt, r0, r1, 0.5 * self.logdet * N2, c1, small, gs[0:3 * Q])
        else:
            Kinv1 = o.mat(0, 0)
            o.gemm(V1_c, A_c, False, True, -1.0, lead * 0.5 * self.logdet * N2, Kinv1)      # Kbar1 partial
            Dbar1 = o.gemm(G_c, A_c, False, True, c1, 0.0, o.new("Dbar1", (N1, N1)))
            o.theta_grad(0, Kinv1, Dbar1, small, gs[0:3 * Q])
        # axis-2 work in R layout
        if fft2:
            P_r = o.toeplitz_rows(1, True, G_r, 1.0, 0.0, small, o.new("P_r", (h, N2)))
        else:
            P_r = o.gemm(G_r, D2, False, False, 1.0, 0.0, o.new("P_r", (h, N2)))
        S2_r = o.apply_kinv(1, 1, P_r, "S2_r")
        W_r = self.c2r(W_c)
        S1_r = self.c2r(S1_c)
        gU_r, V2_r = o.grad_u(U_r, G_r, W_r, S1_r, S2_r, self.bidx, eb, self.nseg0, small)
        if fft2:
            r0, r1 = self.rank * N2 // self.P, (self.rank + 1) * N2 // self.P
            o.theta_grad_rows(1, V2_r, Bt_r, G_r, r0, r1, 0.5 * self.logdet * N1, 1.0, small, gs[3 * Q:6 * Q])
        else:
            Kinv2 = o.mat(1, 0)
            o.gemm(V2_r, Bt_r, True, False, -1.0, lead * 0.5 * self.logdet * N1, Kinv2)     # Kbar2 partial
            Dbar2 = o.gemm(G_r, Bt_r, True, False, 1.0, 0.0, o.new("Dbar2", (N2, N2)))
            o.theta_grad(1, Kinv2, Dbar2, small, gs[3 * Q:6 * Q])
        self._allreduce(gs)
        gs[6 * Q] = gtau
        gs[6 * Q + 1] = gv
        return self.terms, gU_r, gs

    def step_host(self, hU, hsmall, hloss=None):
        """Collective: one step on HOST params - this rank's row block of U (h x N2) and the replicated small params, both in
        pinned memory, updated in place; hloss (1,) receives the loss.  The Adam state stays on the device (like optax's
        state in the reference's loop).  All-FFT plans on CUDA hide the copies: the upload of U runs on a copy stream beside
        the factor stage (which needs only theta) and gates the first exchange; once dL/dU is complete Adam(U) and the
        download of U run on the copy stream beside the theta-gradient tail, the all-reduce and the small leaves' update.
        Same kernels on the same data as step(): bitwise the same result."""
        o = self.ops
        on_cuda = self.U.is_cuda and getattr(o, "uses_gs", None) is not None and o.uses_gs(0) and o.uses_gs(1)
        adam_inc = getattr(o, "adam_inc", None)
        if not on_cuda:                              # general path / CPU stand-in: plain semantics
            self.U.copy_(hU, non_blocking=True)
            self.small.copy_(hsmall, non_blocking=True)
            self.step()
            hU.copy_(self.U, non_blocking=True)
            hsmall.copy_(self.small, non_blocking=True)
            if hloss is not None:
                hloss.copy_(self.last_loss().reshape(1), non_blocking=True)
            if self.U.is_cuda:
                torch.cuda.synchronize(self.U.device)
            return
        dev = self.U.device
        cur = torch.cuda.current_stream(dev)
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=dev)
        cs = self._copy
        self.small.copy_(hsmall, non_blocking=True)  # tiny, and the factor stage needs it
        cs.wait_stream(cur)                          # earlier readers of U are done
        with torch.cuda.stream(cs):
            self.U.copy_(hU, non_blocking=True)
            self._u_ready = torch.cuda.Event()
            self._u_ready.record(cs)
        done = []

        def on_gu(gU_r):
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            cs.wait_event(ev)
            with torch.cuda.stream(cs):
                o.adam(self.U, gU_r, self.mU, self.vU, self.count, self.lr)
                e2 = torch.cuda.Event()
                e2.record(cs)                            # Adam(U) has read the count
                hU.copy_(self.U, non_blocking=True)
            done.append(e2)

        self._on_gu = on_gu
        try:
            _, gU_r, gs = self.value_and_grad()
        finally:
            self._u_ready, self._on_gu = None, None
        if done:
            cur.wait_event(done[0])                  # Adam(U) has read the count (the download may still be running)
        else:
            o.adam(self.U, gU_r, self.mU, self.vU, self.count, self.lr)
            hU.copy_(self.U, non_blocking=True)
        if adam_inc is not None:
            adam_inc(self.small, gs, self.msmall, self.vsmall, self.count, self.lr)
        else:
            o.adam(self.small, gs, self.msmall, self.vsmall, self.count, self.lr)
            self.count += 1
        hsmall.copy_(self.small, non_blocking=True)
        if hloss is not None:
            hloss.copy_(self.last_loss().reshape(1), non_blocking=True)
        cur.synchronize()
        cs.synchronize()

    def step(self):
        """Collective: value_and_grad + Adam on the local U rows and on the (replicated) small params."""
        _, gU_r, gs = self.value_and_grad()
        o = self.ops
        o.adam(self.U, gU_r, self.mU, self.vU, self.count, self.lr)
        adam_inc = getattr(o, "adam_inc", None)
        if adam_inc is not None:                     # Adam on the small leaves and ++count in one launch
            adam_inc(self.small, gs, self.msmall, self.vsmall, self.count, self.lr)
        else:
            o.adam(self.small, gs, self.msmall, self.vsmall, self.count, self.lr)
            self.count += 1
