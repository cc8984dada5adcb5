"""Development diagnostic: per-leaf relative errors of the GPU path vs the CPU oracle for the
theta-gradient variants (force_general bits: 0 default FFT sums, 2 GEMM + direct sums, 4 FFT incl. K^-1)."""
import math
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gphm_b200 as G
from oracle import gphm_oracle as O
from helpers import tree_flatten

CASES = [("poisson_2d-sin_add_cos", "Matern52_Cos_1d", 150, 131, 30, 20.0, 2 * math.pi, 1.0),
         ("allencahn_2d-mix-sincos", "SE_Cos_1d", 130, 130, 9, 10.0, 1.0, 1.0),
         ("advection-sin", "Matern52_Cos_1d", 131, 90, 8, 4.0, 1.0, 200.0),
         ("poisson_2d-sin_sin", "Matern52_Cos_1d", 400, 400, 30, 20.0, 2 * math.pi, 1.0)]
for eq, ker, N1, N2, Q, fs, scale, beta in CASES:
    p, (xt, yt), ut = O.make_problem_2d(eq, ker, N1, scale, beta=beta, M=8, N2=N2)
    for state_name, params in (("S1", O.state_S1(p, Q=Q, freq_scale=fs)), ("S0", O.init_params_2d(N1, N2, Q, fs))):
        _, want = O.loss_and_grad_efficient(p, params) if N1 > 200 else O.loss_and_grad_literal(p, params)
        want = dict(tree_flatten(want))
        for mode in (0, 8, 2):
            tp = {"equation": eq, "kernel": ker, "Q": Q, "freq_scale": fs, "N_col": N1, "llk_weight": 200.0, "lr": 0.01,
                  "logdet": True, "nepoch": 1, "tol": -1, "beta": beta, "force_general": mode}
            cls = G.GP_solver_2d_single_advection if eq.startswith("advection") else G.GP_solver_2d_single
            m = cls(p.bvals.numpy(), (p.x.numpy(), p.y.numpy()), p.src.numpy(), 1e-6, (xt.numpy(), yt.numpy()), ut.numpy(), tp)
            _, g = m.value_and_grad(params)
            errs = {k: float((v.reshape(-1) - want[k].reshape(-1)).norm() / max(float(want[k].norm()), 1e-300))
                    for k, v in tree_flatten(g) if "kernel_paras" in k}
            print("%-26s %-16s %s mode %d  max rel err %.2e  %s" % (eq, ker, state_name, mode, max(errs.values()),
                  " ".join("%s=%.1e" % (k.split("/")[0][-1] + k.split("/")[1][:5], v) for k, v in errs.items())), flush=True)
