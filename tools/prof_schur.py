"""Times the Schur/Levinson recursion of one 4096-point axis pair (plan factor stage) and checks g against torch's Cholesky.
GPHM_SCHUR_SPLIT = 1 | 2 | 4 selects the CTAs per role.     python tools/prof_schur.py [n]"""
import ctypes
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gphm_b200 as G
from oracle import gphm_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.linspace(0, 1, n, dtype=torch.float64) * 2 * math.pi
th = O.init_params_2d(8, 8, 30, 20.0)["kernel_paras_1"]
K = O.gram("Matern52_Cos_1d", x, x, th, 0, 1e-6)
X, g, sK, logdet, status = G.solver_core.toeplitz_solve(K[:, 0])
torch.cuda.synchronize()
L = torch.linalg.cholesky(K)
g_ref = torch.cholesky_solve(torch.eye(n, dtype=torch.float64)[:, :1], L)[:, 0]
ld_ref = float(2 * torch.log(torch.diagonal(L)).sum())
err_g = float((g.cpu() - g_ref).norm() / g_ref.norm())
lib = G._lib.load()
NF = 8
ms = (ctypes.c_double * NF)(); fl = (ctypes.c_double * NF)(); by = (ctypes.c_double * NF)(); nl = (ctypes.c_longlong * NF)()
reps = 20
lib.gphm_profile_start()
for _ in range(reps):
    G.solver_core.toeplitz_solve(K[:, 0])
lib.gphm_profile_stop(ms, fl, by, nl)
print("GPHM_SCHUR_SPLIT=%s n=%d: recursion %.4f ms per launch (%d launches), status %d, g rel err vs Cholesky %.2e, logdet rel err %.2e"
      % (os.environ.get("GPHM_SCHUR_SPLIT", "default"), n, ms[2] / max(nl[2], 1), nl[2], int(status), err_g, abs(float(logdet) - ld_ref) / abs(ld_ref)))
