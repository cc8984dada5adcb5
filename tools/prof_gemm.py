"""ncu target: one launch each of libgphm's DGEMM (NN, NT, TN) and the cuBLAS DGEMM at N=4096."""
import sys
import torch
sys.path.insert(0, ".")
import gphm_b200 as G
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A = torch.randn(N, N, dtype=torch.float64, device="cuda"); B = torch.randn(N, N, dtype=torch.float64, device="cuda")
C = torch.zeros(N, N, dtype=torch.float64, device="cuda")
for tA, tB in ((False, False), (False, True), (True, False)):
    G.solver_core.dgemm(A, B, tA, tB, C=C)
torch.cuda.synchronize()
torch.matmul(A, B, out=C)
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
