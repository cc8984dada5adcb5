"""Golden vectors for the problem set-up step (manufactured source term, meshes, boundary values) produced by
EXECUTING the reference's own get_source_val / get_mesh_data / get_boundary_vals
(/root/reference/code/model_GP_solver_{1d,2d,advection}.py:299-307, 355-379, 354-379) through the torch-backed jax
stand-in of tests/golden/ref_exec_shim/ - with this package's equation lambdas as the solution u (the reference
keeps its own inside test(); the formulas are the same by inspection and two of them are pinned by the shipped
result logs).  Writes tests/golden/ref_setup.npz; tests/test_host_setup_ref_exec.py holds the host code to it.

    python tests/golden/make_ref_setup_golden.py        # needs /root/reference"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
N2D, N1D, M = 9, 15, 6


def _wrap(u):
    return lambda *a: u(*[torch.as_tensor(t) for t in a])


def main():
    sys.path.insert(0, "/root/reference/code")
    sys.path.insert(0, os.path.join(HERE, "ref_exec_shim"))
    sys.path.insert(0, ROOT)
    os.chdir("/tmp")
    import model_GP_solver_2d as R2
    import model_GP_solver_1d as R1
    import model_GP_solver_advection as RA
    import gphm_b200 as G
    out = {}
    arr = lambda t: np.asarray(torch.as_tensor(t).detach().numpy(), dtype=np.float64)
    for name, u in G.model_GP_solver_2d.equation_dict.items():
        eq_type = name.split("-")[0]
        for scale in (1.0, 2 * math.pi):
            x, y, um = R2.get_mesh_data(_wrap(u), N2D, N2D - 2, scale)
            tag = "2d|%s|%.6f" % (name, scale)
            out[tag + "|x"], out[tag + "|y"], out[tag + "|u_mesh"] = arr(x), arr(y), arr(um)
            out[tag + "|bvals"] = arr(R2.get_boundary_vals(um))
            out[tag + "|src"] = arr(R2.get_source_val(_wrap(u), x, y, eq_type))
    for name, u in G.model_GP_solver_1d.equation_dict.items():
        eq_type = name.split("-")[0]
        for scale in (1.0, 2 * math.pi):
            x = np.linspace(0, 1, num=N1D) * scale
            out["1d|%s|%.6f|src" % (name, scale)] = arr(R1.get_source_val(_wrap(u), x, eq_type))
    for beta in (2.0, 200.0):
        u = G.model_GP_solver_advection.make_equation_dict(beta)["advection-sin"]
        x, y, um = RA.get_mesh_data(_wrap(u), N2D, N2D + 1, 1.0)
        tag = "adv|advection-sin|%.1f" % beta
        out[tag + "|x"], out[tag + "|y"], out[tag + "|u_mesh"] = arr(x), arr(y), arr(um)
        out[tag + "|bvals"] = arr(RA.get_boundary_vals(um))
        out[tag + "|src"] = arr(RA.get_source_val(_wrap(u), x, y, "advection", beta))
    np.savez_compressed(os.path.join(HERE, "ref_setup.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
