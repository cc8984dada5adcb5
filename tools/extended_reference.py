"""Extended-precision reference for ONE value_and_grad of the 2-D Poisson problem on a uniform grid (test infrastructure,
CPU only).  Every K^-1 application is a float64 Cholesky solve followed by iterative refinement with the residual
formed in 80-bit long double through FFT Toeplitz products; derivative-Gram products, diagonal sums and the final
contractions run in long double too.  The result is accurate to ~1e-13 on every gradient leaf, i.e. it can arbitrate
between two FP64 implementations (the CPU oracle's Cholesky route and the GPU's Schur/Gohberg-Semencul route) that
disagree at the 1e-6 level when cond(K) ~ 5e7.

    python tools/extended_reference.py N [gpu_grad.npz]      (N = 1024 ... 4096)
"""
import math
import os
import sys
import time

import numpy as np
import scipy.linalg as sla
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gphm_oracle as O

LD = np.longdouble


def toep_mul_ld(tab, V, antisym=False):
    """T V (columns) in long double, T[i,j] = tab[|i-j|] (* sign(i-j) if antisym)."""
    n = len(tab); L = 2 * n
    c = np.zeros(L, dtype=LD); c[:n] = tab
    c[L - n + 1:] = (-tab[:0:-1] if antisym else tab[:0:-1])
    if antisym:
        c[0] = 0
    out = np.empty(V.shape, dtype=LD)
    fc = np.fft.rfft(c)
    for j0 in range(0, V.shape[1], 256):
        blk = V[:, j0:j0 + 256].astype(LD)
        out[:, j0:j0 + 256] = np.fft.irfft(fc[:, None] * np.fft.rfft(blk, L, axis=0), L, axis=0)[:n]
    return out


def solve_refined(tabK, cf, B, iters=3):
    """K^-1 B to ~long-double accuracy (K SPD Toeplitz with first column tabK, cf its float64 Cholesky factor)."""
    X = sla.cho_solve(cf, B).astype(LD)
    Bl = B.astype(LD)
    for _ in range(iters):
        R = Bl - toep_mul_ld(tabK, X)
        X = X + sla.cho_solve(cf, R.astype(np.float64)).astype(LD)
    return X


def xcorr_diag_sums_ld(X, Y, antisym=False):
    """s[d] = sum_i (X Y^T)[i, i+d] (+/-) (X Y^T)[i+d, i]  for d >= 0 (d = 0 counted once), X, Y: n x m, long double."""
    n = X.shape[0]; L = 2 * n
    acc = np.zeros(L // 2 + 1, dtype=np.clongdouble)
    for j0 in range(0, X.shape[1], 256):
        fx = np.fft.rfft(X[:, j0:j0 + 256].astype(LD), L, axis=0)
        fy = np.fft.rfft(Y[:, j0:j0 + 256].astype(LD), L, axis=0)
        acc += (np.conj(fx) * fy).sum(axis=1)
    r = np.fft.irfft(acc, L)               # r[d] = sum_i sum_j X[i,j] Y[i+d,j]  (d >= 0), r[L-d] for negative lags
    up = r[:n].copy()                      # entries (i, i+d)
    lo = np.zeros(n, dtype=LD); lo[1:] = r[L - 1:L - n:-1]      # entries (i+d, i)
    s = up - lo if antisym else up + lo
    if not antisym:
        s[0] = up[0]
    else:
        s[0] = 0
    return s


def reference(p, params):
    U = params["U"].numpy()
    N1, N2 = U.shape
    tau, v = float(params["log_tau"]), float(params["log_v"])
    lam, ld = p.llk_weight, float(p.logdet)
    ax = []
    for x, th in ((p.x, params["kernel_paras_1"]), (p.y, params["kernel_paras_2"])):
        d = (x - x[0]).abs()
        tK = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], 0).sum(-1).numpy().copy()
        tD = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], 2).sum(-1).numpy().copy()
        tK[0] += p.jitter
        K = sla.toeplitz(tK)
        cf = sla.cho_factor(K, lower=True)
        ax.append(dict(tK=tK.astype(LD), tD=tD.astype(LD), cf=cf, x=x, th=th))
    k1 = lambda B: solve_refined(ax[0]["tK"], ax[0]["cf"], B)
    k2 = lambda B: solve_refined(ax[1]["tK"], ax[1]["cf"], B.T).T
    A = k1(U); Bt = k2(U)
    R = toep_mul_ld(ax[0]["tD"], A) + toep_mul_ld(ax[1]["tD"], Bt.T).T - p.src.numpy().astype(LD)
    G = LD(math.exp(v)) * R
    P1 = toep_mul_ld(ax[0]["tD"], G) + LD(0.5) * Bt
    P2 = toep_mul_ld(ax[1]["tD"], G.T).T + LD(0.5) * A
    V1 = k1(P1.astype(np.float64)) + 0                      # the refinement starts from the float64 right-hand side:
    V1 = V1 + k1((P1 - P1.astype(np.float64)).astype(np.float64))      # add the part of P1 lost by rounding it
    V2 = k2(P2.astype(np.float64)); V2 = V2 + k2((P2 - P2.astype(np.float64)).astype(np.float64))
    eb = O.boundary_vector_2d(torch.from_numpy(U)).numpy() - p.bvals.reshape(-1).numpy()
    gU = V1 + V2
    s = lam * math.exp(tau)
    gU[0, :] += s * eb[:N2]; gU[-1, :] += s * eb[N2:2 * N2]; gU[:, 0] += s * eb[2 * N2:2 * N2 + N1]; gU[:, -1] += s * eb[2 * N2 + N1:]
    grads = {"U": gU}
    I = np.eye(N1)
    for a, (V, Y, Gm, nb) in enumerate(((V1, A, G, N2), (V2.T, Bt.T, G.T, N1))):
        X = ax[a]
        Kinv = solve_refined(X["tK"], X["cf"], np.eye(len(X["tK"])))
        n = Kinv.shape[0]
        sKinv = np.array([np.trace(Kinv, d) for d in range(n)], dtype=LD)
        sKinv[1:] *= 2
        xVY = xcorr_diag_sums_ld(V, Y)
        sK = LD(0.5 * ld * nb) * sKinv - xVY
        sD = xcorr_diag_sums_ld(Gm, Y)
        grads["pieces_%d" % (a + 1)] = dict(sKinv=sKinv, xVY=xVY, sD=sD, g=Kinv[:, 0].copy())
        th = X["th"]; d = (X["x"] - X["x"][0]).abs()
        _, pK = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], 0, True)
        _, pD = O.kernel_terms(p.kernel, d, th["log-w"], th["log-ls"], th["freq"], 2, True)
        g = [(sK[:, None] * pk.numpy().astype(LD)).sum(0) + (sD[:, None] * pd.numpy().astype(LD)).sum(0) for pk, pd in zip(pK, pD)]
        grads["kernel_paras_%d" % (a + 1)] = {"log-w": g[0], "log-ls": g[1], "freq": g[2]}
    return grads


def rel(a, b):
    a = np.asarray(a, dtype=LD).reshape(-1); b = np.asarray(b, dtype=LD).reshape(-1)
    return float(np.sqrt(((a - b) ** 2).sum()) / np.sqrt((b ** 2).sum()))


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    KERNEL = sys.argv[3] if len(sys.argv) > 3 else "Matern52_Cos_1d"
    SCALE = float(sys.argv[4]) if len(sys.argv) > 4 else 2 * math.pi
    torch.set_num_threads(os.cpu_count() or 1)
    p, _, _ = O.make_problem_2d("poisson_2d-sin_add_cos", KERNEL, N, SCALE, M=8)
    s1 = O.state_S1(p)
    cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "extref_%d%s.npz" % (N, "" if len(sys.argv) <= 3 else "_%s_s%g" % (KERNEL, SCALE)))
    if os.path.exists(cache):
        z = np.load(cache)
        ref = {"U": z["U"]}
        for a in (1, 2):
            ref["kernel_paras_%d" % a] = {l: z["kp%d_%s" % (a, l)] for l in ("log-w", "log-ls", "freq")}
        print("# extended-precision reference at N=%d: cached (%s; long double rounded to float64)" % (N, cache))
    else:
        t0 = time.time()
        ref = reference(p, s1)
        print("# extended-precision reference at N=%d: %.0f s" % (N, time.time() - t0))
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        flat = {"U": ref["U"].astype(np.float64)}
        for a in (1, 2):
            for l in ("log-w", "log-ls", "freq"):
                flat["kp%d_%s" % (a, l)] = np.asarray(ref["kernel_paras_%d" % a][l], dtype=np.float64)
        np.savez(cache, **flat)
    _, ge = O.loss_and_grad_efficient(p, s1)
    gpu = np.load(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
    print("# leaf | oracle (FP64 Cholesky route) vs reference | GPU (Schur + Gohberg-Semencul route) vs reference | GPU vs oracle")
    leaves = [("U", ge["U"].numpy(), ref["U"])]
    for a in (1, 2):
        for l in ("log-w", "log-ls", "freq"):
            leaves.append(("kernel_paras_%d/%s" % (a, l), ge["kernel_paras_%d" % a][l].numpy(), ref["kernel_paras_%d" % a][l]))
    for name, o, r in leaves:
        line = "%-22s %.2e" % (name, rel(o, r))
        if gpu is not None:
            gname = name.replace("/", "_")
            if name == "U":                      # only two rows of dL/dU travel back from the GPU box
                n = o.shape[0]
                g2 = np.stack([gpu["U_row0"], gpu["U_rowmid"]]); o, r = o[[0, n // 2]], r[[0, n // 2]]
                line = "%-22s %.2e" % ("U (rows 0, N/2)", rel(o, r))
                line += "   %.2e   %.2e" % (rel(g2, r), rel(g2, o))
            else:
                line += "   %.2e   %.2e" % (rel(gpu[gname], r), rel(gpu[gname], o))
        print(line)
