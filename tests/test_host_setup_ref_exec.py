"""CPU: this package's problem set-up functions (manufactured source, meshes, boundary values: host code that runs
once before the hot path) against the outputs of the reference's own get_source_val / get_mesh_data /
get_boundary_vals, executed by tests/golden/make_ref_setup_golden.py.  Covers every equation of the three
equation tables, both scales, the row-major flattening and the edge order of hstack(U[0,:],U[-1,:],U[:,0],U[:,-1])."""
import math
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_setup.npz"))
N2D, N1D = 9, 15


def _scale(txt):                                # the tag carries six decimals of the scale
    v = float(txt)
    return 2 * math.pi if abs(v - 2 * math.pi) < 1e-5 else v


def _close(a, b, tol=1e-11):
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("tag", sorted({"|".join(k.split("|")[:3]) for k in GOLD.files if k.startswith("2d")}))
def test_setup_2d(gphm, tag):
    m = gphm.model_GP_solver_2d
    _, name, scale = tag.split("|")
    u, scale = m.equation_dict[name], _scale(scale)
    x, y, um = m.get_mesh_data(u, N2D, N2D - 2, scale)
    _close(x, GOLD[tag + "|x"]); _close(y, GOLD[tag + "|y"]); _close(um, GOLD[tag + "|u_mesh"])
    _close(m.get_boundary_vals(um), GOLD[tag + "|bvals"])
    # second derivatives of sin(100 x)-type solutions are ~1e4: relative to the largest entry
    _close(m.get_source_val(u, x, y, name.split("-")[0]), GOLD[tag + "|src"], 1e-12)


@pytest.mark.parametrize("tag", sorted({"|".join(k.split("|")[:3]) for k in GOLD.files if k.startswith("1d")}))
def test_setup_1d(gphm, tag):
    m = gphm.model_GP_solver_1d
    _, name, scale = tag.split("|")
    x = np.linspace(0, 1, num=N1D) * _scale(scale)
    _close(m.get_source_val(m.equation_dict[name], x, name.split("-")[0]), GOLD[tag + "|src"], 1e-12)


@pytest.mark.parametrize("beta", [2.0, 200.0])
def test_setup_advection(gphm, beta):
    m, m2 = gphm.model_GP_solver_advection, gphm.model_GP_solver_2d
    tag = "adv|advection-sin|%.1f" % beta
    u = m.make_equation_dict(beta)["advection-sin"]
    x, y, um = m2.get_mesh_data(u, N2D, N2D + 1, 1.0)
    _close(um, GOLD[tag + "|u_mesh"]); _close(m2.get_boundary_vals(um), GOLD[tag + "|bvals"])
    _close(m.get_source_val(u, x, y, "advection", beta), GOLD[tag + "|src"], 1e-12)
