"""Per-role cycle counts of the Schur/Levinson kernel (generator CTA vs lattice CTA) through gphm_toeplitz_solve."""
import os, sys, math
os.environ["GPHM_SCHUR_CYCLES"] = "1"
sys.path.insert(0, ".")
import torch, ctypes
import gphm_b200 as G
from oracle import gphm_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.linspace(0, 1, n, dtype=torch.float64) * 2 * math.pi
th = O.init_params_2d(8, 8, 30, 20.0)["kernel_paras_1"]
d = (x - x[0]).abs()
t = O.kernel_terms("Matern52_Cos_1d", d, th["log-w"], th["log-ls"], th["freq"], 0).sum(-1)
t[0] += 1e-6
lib = G._lib.load()
td = t.cuda()
B = torch.zeros(2, n, dtype=torch.float64, device="cuda")      # rows >= 1 so that the scratch (used as debug buffer) exists
X = torch.empty_like(B); g = torch.empty_like(td); sK = torch.empty_like(td)
ld = torch.empty(1, dtype=torch.float64, device="cuda"); stt = torch.zeros(1, dtype=torch.int32, device="cuda")
work = torch.zeros(lib.gphm_toeplitz_work_bytes(n, 2), dtype=torch.uint8, device="cuda")
P = G._lib.ptr
for it in range(3):
    # rows = 0: only the recursion + spectra run, the scratch keeps the cycle counters
    G._lib.check(lib.gphm_toeplitz_solve(P(td), n, None, 0, None, P(g), P(sK), P(ld), P(stt), P(work), G._lib.stream_ptr()), "solve")
    torch.cuda.synchronize()
    L = 2
    while L < 2 * n: L <<= 1
    off = ((2 * L * 8 + 255) // 256 * 256) + ((8 * L * 8 + 255) // 256 * 256) + 256
    cyc = work[off:off + 16].view(torch.int64).tolist()
    tk = work[off + 64:off + 104].view(torch.int64).tolist()
    print("   generator warps: led batch %.0f cycles (%d), bulk batch %.0f cycles (%d), barrier wait per warp-period %.0f" %
          (tk[0] / max(tk[1], 1), tk[1], tk[2] / max(tk[3], 1), tk[3], tk[4] / max((n // 8 + max(n // 256, 1)) * max(n // 256, 1), 1)))
    work[off:off + 128].zero_()
    print("n=%d generator %d cycles (%.0f/step), lattice %d cycles (%.0f/step), status %d" % (n, cyc[0], cyc[0] / n, cyc[1], cyc[1] / n, int(stt)))
