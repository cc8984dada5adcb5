"""ncu target: a few steps of the bench workload (N=4096 by default) on the default path."""
import sys
sys.path.insert(0, ".")
import bench
import torch
import gphm_b200 as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
tp, bvals, X_col, src, X_test, u_test = bench.build_inputs(n)
core = G.solver_core.SolverCore(2, bench.KERNEL, "poisson", X_col[0], X_col[1], src, bvals, None, bench.LLK, 1.0, 1.0, 1e-6, bench.Q)
class _M:
    trick_paras, N1, N2 = tp, n, n
st = core.new_state(G.GP_solver_2d_single.init_params(_M))
for _ in range(steps):
    core.step_inplace(st, 0.01)
torch.cuda.synchronize()
print("ok", float(st.terms[0]))
