"""Times one K^-1 application (gs_apply_fused_kernel, 4096 rows of length 4096) of a 4096^2 plan; GPHM_GS_VARIANT selects the
kernel variant.  Prints ms per launch and a checksum / residual so that variants can be compared across processes."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gphm_b200 as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tp, bvals, X_col, src, X_test, u_test = bench.build_inputs(n)
core = G.solver_core.SolverCore(2, bench.KERNEL, "poisson", X_col[0], X_col[1], src, bvals, None, bench.LLK, 1.0, 1.0, 1e-6, bench.Q)


class _M:
    trick_paras, N1, N2 = tp, n, n


st = core.new_state(G.GP_solver_2d_single.init_params(_M))
lib = core.lib
small = st.small
G._lib.check(lib.gphm_plan_factor(core.plan, G._lib.ptr(small), 3, G._lib.stream_ptr()), "factor")
g = torch.Generator().manual_seed(0)
X = torch.randn(n, n, generator=g, dtype=torch.float64).cuda()
out, tmp = torch.empty_like(X), torch.empty_like(X)


def apply():
    G._lib.check(lib.gphm_apply_kinv(core.plan, 1, 1, G._lib.ptr(X), n, n, G._lib.ptr(out), G._lib.ptr(tmp), G._lib.stream_ptr()), "kinv")


for _ in range(3):
    apply()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 20
for _ in range(reps):
    apply()
e1.record()
torch.cuda.synchronize()
print("GPHM_GS_VARIANT=%s  %.4f ms per K^-1 application   checksum %.15e  |out| %.6e" %
      (os.environ.get("GPHM_GS_VARIANT", "default"), e0.elapsed_time(e1) / reps, float(out.sum()), float(out.norm())))
