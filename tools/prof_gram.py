"""General-grid Gram pair (K, D = k'') on a non-uniform 4096-point grid, Q = 30: the full kernel (rectangular call shape)
against the symmetric-half kernel the plans use.      python tools/prof_gram.py [N] [--once]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gphm_b200 as G
from helpers import nonuniform_grid, theta_state

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 4096
Q = 30
lib = G._lib.load()
x = nonuniform_grid(N, 2 * math.pi, 1).cuda()
x2 = x.clone()                                     # a different pointer: the rectangular (full) kernel
th = theta_state(Q, 20.0)
theta = torch.cat([th["log-w"], th["log-ls"], th["freq"]]).cuda()
K = torch.empty(N, N, dtype=torch.float64, device="cuda")
D = torch.empty_like(K)
P, sp = G._lib.ptr, G._lib.stream_ptr
core = G.solver_core.SolverCore(2, "Matern52_Cos_1d", "poisson", x.cpu().numpy(), x.cpu().numpy(), torch.zeros(N, N).numpy(),
                                torch.zeros(4 * N).numpy(), None, 200.0, 1.0, 1.0, 1e-6, Q)
small = torch.zeros(6 * Q + 2, dtype=torch.float64, device="cuda")
small[:3 * Q] = theta; small[3 * Q:6 * Q] = theta


def full():
    G._lib.check(lib.gphm_gram(1, 2, P(x), N, P(x2), N, P(theta), Q, 0.0, P(D), sp()), "gram")


def plan_pairs():                                  # both axes: 2 x gram_symmetric_kernel (+ Cholesky etc. not timed here)
    pass


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if "--once" in sys.argv:
    st = core.new_state()
    st.small.copy_(small)
    core.value_and_grad(st, forward_only=True)     # general path of a non-uniform plan: launches gram_symmetric_kernel twice
    torch.cuda.synchronize()
    sys.exit(0)
t_full = timed(full)
NF = 8
import ctypes
ms = (ctypes.c_double * NF)(); fl = (ctypes.c_double * NF)(); by = (ctypes.c_double * NF)(); nl = (ctypes.c_longlong * NF)()
st = core.new_state(); st.small.copy_(small)
core.value_and_grad(st, forward_only=True); torch.cuda.synchronize()
lib.gphm_profile_start()
core.value_and_grad(st, forward_only=True)
lib.gphm_profile_stop(ms, fl, by, nl)
print("N=%d Q=%d Matern52_Cos_1d, non-uniform grid" % (N, Q))
print("full kernel, one matrix (k''): %.3f ms = %.1f GB/s of output, %.2f G transcendental pairs/s" % (t_full, 8.0 * N * N / t_full / 1e6, N * N * Q / t_full / 1e6))
print("plan Gram family (2 axes x (K, D) by gram_symmetric_kernel): %.3f ms for %d launches = %.3f ms per (K, D) pair, %.1f GB/s of output"
      % (ms[0], nl[0], ms[0] / max(nl[0], 1), by[0] / ms[0] / 1e6))
