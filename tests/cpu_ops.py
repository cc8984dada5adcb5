"""torch-CPU stand-in for dist.CudaOps (TEST INFRASTRUCTURE): lets the sharded step's host logic
(layouts, exchanges, partial reductions) run over gloo on a CPU-only box.  Numerics come from the
oracle's formulas; nothing here is shipped."""
import math

import torch

from oracle import gphm_oracle as O

DT = torch.float64


class CpuOps(object):
    def __init__(self, kernel, eq_name, x, y, llk_weight, Q, beta=1.0, jitter=1e-6, gs=True):
        self.kernel, self.eq_name, self.Q, self.jitter = kernel, eq_name, Q, jitter
        self.x, self.y = torch.as_tensor(x, dtype=DT), torch.as_tensor(y, dtype=DT)
        self.order = 1 if eq_name == "advection" else 2
        self.llk_weight = llk_weight
        self.m = {}
        self.ld = torch.zeros(2, dtype=DT)
        self.gs = gs                 # stand in for the all-FFT step (uniform grids) or the general one
        self.kept = {}

    def zeros(self, shape, dtype=DT):
        return torch.zeros(shape, dtype=dtype)

    def tensor(self, data, dtype=DT):
        return torch.as_tensor(data, dtype=dtype).contiguous()

    def new(self, tag, shape):
        return torch.empty(shape, dtype=DT)

    def _theta(self, small, a):
        Q = self.Q
        return {"log-w": small[(3 * a) * Q:(3 * a + 1) * Q], "log-ls": small[(3 * a + 1) * Q:(3 * a + 2) * Q],
                "freq": small[(3 * a + 2) * Q:(3 * a + 3) * Q]}

    def factor(self, small, axis_mask=3):
        for a, xs in enumerate((self.x, self.y)):
            n = xs.numel()
            for which in (0, 1, 2, 3):               # persistent buffers: broadcasts write into them in place
                self.m.setdefault((a, which), torch.zeros(n, n, dtype=DT))
            if not (axis_mask >> a) & 1:
                continue
            K, D = O._gram_pair(self.kernel, xs, self._theta(small, a), self.order, self.jitter)
            L = torch.linalg.cholesky(K)
            self.m[(a, 3)].copy_(L)
            self.m[(a, 1)].copy_(D)
            self.m[(a, 2)].copy_(torch.linalg.solve_triangular(L, torch.eye(n, dtype=DT), upper=False))
            if not axis_mask & 4:
                self.m[(a, 0)].copy_(torch.cholesky_inverse(L))
            self.ld[a] = 2.0 * torch.log(torch.diagonal(L)).sum()

    def mat(self, axis, which):
        return self.m[(axis, which)]

    def logdets(self):
        return self.ld.clone()

    def apply_kinv(self, axis, side, X, tag):
        Li = self.m[(axis, 2)]                        # like the product: two triangular products with L^-1
        if side == 0:
            return Li.T @ (Li @ X)
        return (X @ Li.T) @ Li

    def uses_fft(self, axis):
        return O.is_uniform(self.x if axis == 0 else self.y)

    # ---- the all-FFT step's primitives (uniform grids) ----
    def uses_gs(self, axis):
        return self.gs and O.is_uniform(self.x if axis == 0 else self.y)

    def kinv_rows(self, axis, X, tag, refine=False):
        Li = self.m[(axis, 2)]
        return (X @ Li.T) @ Li

    def toeplitz_rows_add(self, axis, transposed, X, alpha, beta, add, out, keep):
        D = self.m[(axis, 1)]
        res = alpha * (X @ (D if transposed else D.T))
        if beta != 0.0:
            res = res + beta * (add if add is not None else out)
        if keep:
            self.kept[axis] = X.clone()
        out.copy_(res)
        return out

    def theta_grad_pairs(self, axis, V, G, lead, beta, cD, small, out):
        Y = self.kept[axis]
        Li = self.m[(axis, 2)]
        self.theta_grad(axis, (beta if lead else 0.0) * (Li.T @ Li) - V.T @ Y, cD * (G.T @ Y), small, out)

    def grad_u_sum(self, U, G, V1, V2, bidx, eb, nseg0, small):
        g = V1 + V2
        if self.eq_name == "allencahn":
            g = g + G * (3.0 * U * U - 1.0)
        s = self.llk_weight * torch.exp(small[6 * self.Q])
        g.reshape(-1).index_add_(0, bidx.long(), s * eb)
        return g

    def transpose(self, X, tag):
        return X.T.contiguous()

    def toeplitz_rows(self, axis, transposed, X, alpha, beta, small, out):
        # like the product, D comes from `small`, not from a (possibly un-broadcast) factorisation
        _, D = O._gram_pair(self.kernel, self.x if axis == 0 else self.y, self._theta(small, axis), self.order, self.jitter)
        res = alpha * (X @ (D if transposed else D.T))          # row x -> D^T x (transposed) or D x
        out.copy_(res + (beta * out if beta != 0.0 else 0.0))
        return out

    def theta_grad_rows(self, axis, X, Y, G, r0, r1, beta, cD, small, out):
        Li = self.m[(axis, 2)][r0:r1]
        self.theta_grad(axis, beta * (Li.T @ Li) - X.T @ Y, cD * (G.T @ Y), small, out)

    def gemm(self, A, B, tA, tB, alpha, beta, C):
        prod = (A.T if tA else A) @ (B.T if tB else B)
        C.copy_(alpha * prod + (beta * C if beta != 0.0 else 0.0))
        return C

    def residual(self, R, U, F, A, Bt, small, out=None):
        r = R - F
        if self.eq_name == "allencahn":
            r = r + U * (U * U - 1.0)
        res = torch.stack(((r * r).sum(), (A * Bt).sum()))
        R.copy_(torch.exp(small[6 * self.Q + 1]) * r)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def boundary(self, U, bidx, bvals, out=None):
        eb = U.reshape(-1)[bidx.long()] - bvals
        bg = (eb * eb).sum().reshape(1)
        if out is not None:
            out.copy_(bg)
            return eb, out
        return eb, bg

    def grad_u(self, U, G, W, S1, S2, bidx, eb, nseg0, small):
        g = W + S1 + S2
        if self.eq_name == "allencahn":
            g = g + G * (3.0 * U * U - 1.0)
        s = self.llk_weight * torch.exp(small[6 * self.Q])
        g.reshape(-1).index_add_(0, bidx.long(), s * eb)
        return g, S2 + 0.5 * W

    def lincomb(self, a, x, b, y, tag):
        return a * x + b * y

    def theta_grad(self, axis, Kbar, Dbar, small, out):
        xs = self.x if axis == 0 else self.y
        g = O._theta_grad_axis(self.kernel, xs, self._theta(small, axis), self.order, Kbar, Dbar)
        out.copy_(torch.cat((g["log-w"], g["log-ls"], g["freq"])))

    def finalize(self, sums3, ld2, small, terms, gsmall):
        Q = self.Q
        e, q, b = sums3[0], sums3[1], sums3[2]
        tau, v = small[6 * Q], small[6 * Q + 1]
        N1, N2 = self.x.numel(), self.y.numel()
        Nb, Nc = 2 * N1 + 2 * N2, N1 * N2
        loss = (0.5 * (N2 * ld2[0] + N1 * ld2[1]) + 0.5 * q - self.llk_weight * (0.5 * Nb * tau - 0.5 * torch.exp(tau) * b)
                - (0.5 * Nc * v - 0.5 * torch.exp(v) * e))
        gtau = -self.llk_weight * (0.5 * Nb - 0.5 * torch.exp(tau) * b)
        gv = -(0.5 * Nc - 0.5 * torch.exp(v) * e)
        terms.copy_(torch.stack((loss, ld2[0], ld2[1], q, b, e, gtau, gv)))
        gsmall[6 * Q], gsmall[6 * Q + 1] = gtau, gv

    def adam(self, p, g, m, v, count, lr):
        t = int(count) + 1
        m.mul_(0.9).add_(0.1 * g)
        v.mul_(0.999).add_(0.001 * g * g)
        p.sub_(lr * (m / (1 - 0.9 ** t)) / (torch.sqrt(v / (1 - 0.999 ** t)) + 1e-8))
