"""Throughput of the tcgen05 Ozaki GEMM against the native-FP64 DMMA kernel and cuBLAS DGEMM (one GPU).
    python tools/prof_ozaki.py [N] [slices]        prints one line per kernel; run under ncu with --once for a single launch"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gphm_b200 as G

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 4096
S = int(args[1]) if len(args) > 1 else 8
once = "--once" in sys.argv
lib = G._lib.load()
g = torch.Generator().manual_seed(0)
A = torch.randn(N, N, generator=g, dtype=torch.float64).cuda()
B = torch.randn(N, N, generator=g, dtype=torch.float64).cuda()
C = torch.zeros(N, N, dtype=torch.float64, device="cuda")
work = torch.empty(lib.gphm_ozaki_work_bytes(N, N, N, S), dtype=torch.uint8, device="cuda")
P, sp = G._lib.ptr, G._lib.stream_ptr


def oz():
    G._lib.check(lib.gphm_ozaki_dgemm(0, 0, N, N, N, 1.0, P(A), N, P(B), N, 0.0, P(C), N, S, P(work), work.numel(), sp()), "ozaki")


def native():
    G._lib.check(lib.gphm_dgemm(0, 0, N, N, N, 1.0, P(A), N, P(B), N, 0.0, P(C), N, sp()), "dgemm")


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if once:
    oz()
    torch.cuda.synchronize()
    sys.exit(0)
fl = 2.0 * N ** 3
t_oz = timed(oz, 5)
err = float(((C - A @ B).abs().max()) / (A.abs().amax(1).max() * B.abs().amax(0).max()))
t_nat = timed(native, 5)
t_cub = timed(lambda: torch.matmul(A, B, out=C), 5)
pairs = S * (S + 1) // 2
print("N=%d slices=%d (%d int8 digit products)" % (N, S, pairs))
print("ozaki tcgen05 (split + GEMM): %.3f ms  = %.1f TFLOP/s FP64-equivalent, %.0f TOP/s int8 issued;  max err / (rowmax colmax) %.2e (bound %.2e)"
      % (t_oz, fl / t_oz / 1e9, pairs * fl / t_oz / 1e9, err, lib.gphm_ozaki_error_factor(N, S)))
print("native DMMA dgemm_kernel     : %.3f ms  = %.1f TFLOP/s" % (t_nat, fl / t_nat / 1e9))
print("cuBLAS DGEMM (torch.matmul)  : %.3f ms  = %.1f TFLOP/s" % (t_cub, fl / t_cub / 1e9))
