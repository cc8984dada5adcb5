"""Development diagnostics (run under gpurun): kernel timings vs cuBLAS/cuSOLVER comparison baselines."""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
import gphm_b200 as G
from oracle import gphm_oracle as O

DT = torch.float64


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096]
    print(torch.cuda.get_device_name(0))
    for N in sizes:
        A = torch.randn(N, N, dtype=DT, device="cuda"); B = torch.randn(N, N, dtype=DT, device="cuda")
        C = torch.zeros(N, N, dtype=DT, device="cuda")
        fl = 2.0 * N ** 3
        for tA in (False, True):
            for tB in (False, True):
                t, _ = timeit(lambda: G.solver_core.dgemm(A, B, tA, tB, C=C))
                print("N=%d gphm dgemm tA=%d tB=%d: %.3f ms  %.1f TFLOP/s" % (N, tA, tB, t, fl / t / 1e9))
        t, _ = timeit(lambda: torch.matmul(A, B, out=C))
        print("N=%d cuBLAS dgemm: %.3f ms  %.1f TFLOP/s" % (N, t, fl / t / 1e9))
        x = torch.linspace(0, 1, N, dtype=DT) * 2 * math.pi
        th = O.state_S1(O.Problem2D("Matern52_Cos_1d", "poisson_2d", x, x, None, None))["kernel_paras_1"] if False else None
        q = torch.arange(30, dtype=DT)
        th = {"log-w": math.log(1 / 30) - 0.05 * torch.cos(q), "log-ls": 0.1 * torch.sin(q), "freq": 20 * q / 29}
        K = G.Matern52_Cos_1d().gram(x, x, th, 0, 1e-6)
        t, _ = timeit(lambda: G.Matern52_Cos_1d().gram(x, x, th, 2, 0.0))
        print("N=%d general gram: %.3f ms" % (N, t))
        t, _ = timeit(lambda: G.solver_core.potrf_inv(K))
        print("N=%d gphm potrf+trtri: %.3f ms" % (N, t))
        t, _ = timeit(lambda: torch.linalg.cholesky(K))
        print("N=%d cuSOLVER potrf: %.3f ms" % (N, t))
        # full step
        p, (xt, yt), ut = O.make_problem_2d("poisson_2d-sin_add_cos", "Matern52_Cos_1d", N, 2 * math.pi, M=300)
        tp = {"equation": "poisson_2d-sin_add_cos", "kernel": "Matern52_Cos_1d", "Q": 30, "freq_scale": 20.0, "N_col": N,
              "llk_weight": 200.0, "lr": 0.01, "logdet": True, "nepoch": 1, "tol": -1}
        m = G.GP_solver_2d_single(p.bvals.numpy(), (p.x.numpy(), p.y.numpy()), p.src.numpy(), 1e-6,
                                  (xt.numpy(), yt.numpy()), ut.numpy(), tp)
        st = m.core.new_state(m.init_params())
        t, med = timeit(lambda: m.core.step_inplace(st, 0.01), n=5, warm=3)
        print("N=%d full step: best %.3f ms median %.3f ms -> %.2f it/s, %.1f TFLOP/s of 28N^3; status %s" % (
            N, t, med, 1000 / med, 28.0 * N ** 3 / med / 1e9, m.core.status()))
        t, med = timeit(lambda: m.core.value_and_grad(st, forward_only=True), n=3, warm=1)
        print("N=%d forward only: %.3f ms" % (N, med))
        t, med = timeit(lambda: m.core.lib.gphm_plan_factor(m.core.plan, G._lib.ptr(st.small), 3, G._lib.stream_ptr()), n=3, warm=1)
        print("N=%d factor both axes (gram+chol+trtri+kinv): %.3f ms" % (N, med))
        del m, st, A, B, C, K
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
