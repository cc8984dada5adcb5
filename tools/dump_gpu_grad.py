"""GPU side of tools/extended_reference.py: the theta-gradient leaves (and the norm of dL/dU) of one value_and_grad at
state S1 on both GPU routes (default Toeplitz inverse generator; force_general=16 blocked Cholesky), written to
gpurun_out/gpu_grad_<N>[_chol].npz.      python tools/dump_gpu_grad.py N"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gphm_b200 as G
from oracle import gphm_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
KERNEL = sys.argv[2] if len(sys.argv) > 2 else "Matern52_Cos_1d"
SCALE = float(sys.argv[3]) if len(sys.argv) > 3 else 2 * math.pi
p, _, _ = O.make_problem_2d("poisson_2d-sin_add_cos", KERNEL, N, SCALE, M=8)
s1 = O.state_S1(p)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
TAG = "" if (KERNEL == "Matern52_Cos_1d" and len(sys.argv) <= 3) else "_%s_s%g" % (KERNEL, SCALE)
for mode, tag in ((0, ""), (16, "_chol"), (32, "_norefine"), (128, "_refine")):
    core = G.solver_core.SolverCore(2, KERNEL, "poisson", p.x.numpy(), p.y.numpy(), p.src.numpy(), p.bvals.numpy(), None,
                                    p.llk_weight, 1.0, 1.0, 1e-6, 30, force_general=mode)
    st = core.new_state(s1)
    terms, gU, gs = core.value_and_grad(st)
    core.raise_on_bad_status()
    tree = core.unpack_tree(gU, gs)
    out = {"terms": terms.cpu().numpy(), "U_norm": float(gU.norm()), "U_row0": tree["U"][0].cpu().numpy(), "U_rowmid": tree["U"][N // 2].cpu().numpy()}
    for a in (1, 2):
        for l in ("log-w", "log-ls", "freq"):
            out["kernel_paras_%d_%s" % (a, l)] = tree["kernel_paras_%d" % a][l].cpu().numpy()
    np.savez(os.path.join(ROOT, "gpurun_out", "gpu_grad_%d%s%s.npz" % (N, TAG, tag)), **out)
    print("wrote", N, tag or "default", float(terms[0]))
    del core, st
    torch.cuda.empty_cache()
