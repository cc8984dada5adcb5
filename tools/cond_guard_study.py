"""Conditioning study for the Schur/Levinson + Gohberg-Semencul route: the reference's shipped
configs with the plain kernels SE_1d / Matern52_1d at their initial state (log-ls = 0, jitter 1e-6), N_col = 400 / 900,
scale 1 or 2 pi.  Prints cond(K), max |kappa|, min(1 - kappa^2), |g0| growth and the relative error of K^-1 v (GS route,
then after ONE refinement step) against a longdouble-refined Cholesky solve.    python tools/cond_guard_study.py"""
import os, sys, math
import numpy as np, scipy.linalg as sla, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import gphm_oracle as O
from toeplitz_numerics_gs import gs_apply, refine
from toeplitz_numerics_schur import schur_lattice, toep_sym_apply

rel = lambda a, b: float(np.linalg.norm((a - b).astype(np.float64)) / np.linalg.norm(b.astype(np.float64)))
rng = np.random.default_rng(0)
print("# kernel N scale | cond(K) | max|kappa| | min(1-kappa^2) | g0*r0 | err GS | err GS+1 refine | err chol (fp64)")
for kernel in ("SE_1d", "Matern52_1d", "SE_Cos_1d", "Matern52_Cos_1d"):
    for N, scale in ((400, 2 * math.pi), (400, 1.0), (900, 1.0), (900, 2 * math.pi), (200, 1.0)):
        x = torch.linspace(0, 1, N, dtype=torch.float64) * scale
        th = O.init_params_1d(N, 30, 20.0)["kernel_paras"]
        K = O.gram(kernel, x, x, th, 0, 1e-6).numpy()
        r = K[:, 0].copy()
        cf = sla.cho_factor(K, lower=True)
        V = rng.standard_normal((N, 4))
        Xt = refine(K, cf, V)
        g, ld, ks = schur_lattice(r)
        Y = gs_apply(g, V)
        Y1 = Y + gs_apply(g, V - toep_sym_apply(r, Y))
        Yc = sla.cho_solve(cf, V)
        ev = np.linalg.eigvalsh(K)
        print("%-16s %4d %5.2f | %.1e | %.9f | %.2e | %.2e | %.1e | %.1e | %.1e" % (
            kernel, N, scale, ev[-1] / ev[0], np.abs(ks).max(), (1 - ks * ks).min(), g[0] * r[0], rel(Y, Xt), rel(Y1, Xt), rel(Yc, Xt)))
