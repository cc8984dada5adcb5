""""1-D advection" on a 2-D (x, t) grid: beta*u_x + u_y = 0 with first-derivative cross-kernels -
mirror of the reference's model_GP_solver_advection.py (class :31-85, hot path :87-179, problem
setup :354-410, evals :466-513)."""
import numpy as np
import torch

from . import model_GP_solver_2d as m2d
from .infras.exp_config import ExpConfig
from .kernel_matrix import DT


class GP_solver_2d_single_advection(m2d.GP_solver_2d_single):
    """beta * u_x + u_y = src; trick_paras['beta'] (model_GP_solver_advection.py:69)."""
    EQ_TYPES = ("advection",)
    _deriv_order = 1
    _early_stop_enabled = False          # commented out in the reference (:323-328)


def make_equation_dict(beta):
    """model_GP_solver_advection.py:385-388."""
    return {"advection-sin": lambda x, y: torch.sin(x - beta * y)}


def get_source_val(u, x_pos, y_pos, equation_type, beta):
    """model_GP_solver_advection.py:354-362."""
    if equation_type != "advection":
        raise NotImplementedError
    X, Y = torch.meshgrid(torch.as_tensor(x_pos, dtype=DT), torch.as_tensor(y_pos, dtype=DT), indexing="ij")
    return (beta * m2d._derivs(u, [X, Y], 0, 1) + m2d._derivs(u, [X, Y], 1, 1)).reshape(-1).numpy()


def get_boundary_vals_only_init(u_mesh):
    """model_GP_solver_advection.py:377-379 (defined, unused by the reference's test())."""
    return np.hstack((u_mesh[:, 0],))


def build_problem(trick_paras, M=300):
    beta = trick_paras["beta"]
    u = make_equation_dict(beta)[trick_paras["equation"]]
    scale, N = trick_paras["scale"], trick_paras["N_col"]
    x_te, y_te, u_test = m2d.get_mesh_data(u, M, M, scale)
    x_tr, y_tr, u_mh = m2d.get_mesh_data(u, N, N, scale)
    bvals = m2d.get_boundary_vals(u_mh)
    src = get_source_val(u, x_tr, y_tr, "advection", beta).reshape(x_tr.size, y_tr.size)
    return bvals, (x_tr, y_tr), src, (x_te, y_te), u_test


def test(trick_paras):
    return m2d.test(trick_paras, GP_solver_2d_single_advection, build_problem(trick_paras))


def evals(**kwargs):
    """model_GP_solver_advection.py:466-509."""
    args = ExpConfig().parse(kwargs)
    config = m2d.make_config(args.equation, args.kernel, args.nepoch, allowed=["advection-sin"],
                             suffix_fn=lambda c: "-beta-%d" % c["beta"])          # :507
    return test(config)


if __name__ == "__main__":
    m2d._main(evals)
