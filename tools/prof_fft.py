"""ncu target: a few steps at N=4096 so that the FFT kernels (xcorr_spectrum, toeplitz_apply) and the DGEMM show up."""
import sys
sys.path.insert(0, ".")
import subprocess
subprocess.run  # noqa
import bench
import torch
import gphm_b200 as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tp, bvals, X_col, src, X_test, u_test = bench.build_inputs(n)
core = G.solver_core.SolverCore(2, bench.KERNEL, "poisson", X_col[0], X_col[1], src, bvals, None, bench.LLK, 1.0, 1.0, 1e-6, bench.Q)
class _M:
    trick_paras, N1, N2 = tp, n, n
st = core.new_state(G.GP_solver_2d_single.init_params(_M))
for _ in range(2):
    core.step_inplace(st, 0.01)
torch.cuda.synchronize()
print("ok", float(st.terms[0]))
