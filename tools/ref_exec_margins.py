"""CPU emulation (numpy FP64) of the GPU's K^-1 route - Schur/lattice generator + Gohberg-Semencul application
(tools/toeplitz_numerics_*.py) - on the Gram matrices of the executed-reference "g" cases (tests/golden/ref_exec.npz):
distance of K^-1 v and log|K| from LU.  Margin check for tests/test_gpu_zz_reference_exec.py (bound 1e-6).
    python tools/ref_exec_margins.py > profiles/r01_ref_exec_margins.txt"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, d)
from oracle import gphm_oracle as O                                      # noqa: E402
from toeplitz_numerics_gs import gs_apply                                # noqa: E402
from toeplitz_numerics_schur import schur_lattice                        # noqa: E402

T = importlib.import_module("test_oracle_ref_exec")
worst = 0.0
print("# case | axis | cond(K) | rel err of K^-1 v (Schur + Gohberg-Semencul vs LU) | rel err of log|K| | max |kappa|")
for tag in [t for t in T.TAGS if t.startswith("g")]:
    p, xte, like = T._problem(O, tag)
    params = T._tree(tag, "params0/", like)
    two = tag.split("|")[0].endswith("2d")
    for key, x in ([("kernel_paras_1", p.x), ("kernel_paras_2", p.y)] if two else [("kernel_paras", p.x)]):
        K = O.gram(p.kernel, x, x, params[key], 0, p.jitter).numpy()
        g, ld, ks = schur_lattice(K[:, 0].copy())
        V = np.random.default_rng(0).standard_normal((K.shape[0], 5))
        want = np.linalg.solve(K, V)
        e = float(np.linalg.norm(gs_apply(g, V) - want) / np.linalg.norm(want))
        ldw = np.linalg.slogdet(K)[1]
        worst = max(worst, e)
        print("%-46s %-15s %.1e  %.1e  %.1e  %.6f" % (tag, key, np.linalg.cond(K), e, abs(ld - ldw) / abs(ldw), np.abs(ks).max()))
print("# worst K^-1 v error %.1e (test bound 1e-6)" % worst)
