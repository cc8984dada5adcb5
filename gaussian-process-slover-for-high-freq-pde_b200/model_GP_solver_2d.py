"""2-D GP-HM solver (Poisson / Allen-Cahn on a Kronecker grid) - same class, method names,
argument meaning, params pytree and return tuples as the reference's model_GP_solver_2d.py; the
numerical core (value_and_grad of the log-joint + Adam) runs in libgphm on the GPU.

  GP_solver_2d_single.__init__                model_GP_solver_2d.py:40-85
  value_and_grad_kernel / boundary_and_eq_gap :87-143
  loss / step                                 :145-183
  preds / compute_early_stopping              :185-233
  train                                       :235-352
  get_source_val / get_mesh_data / get_boundary_vals / test / evals   :355-514
"""
import time

import numpy as np
import torch

from . import configs, utils
from .infras.exp_config import ExpConfig
from .kernel_matrix import DT, KERNELS, Kernel_matrix, as_dev
from .solver_core import SolverCore, dgemm, solve_spd

try:
    import tqdm
    _progress = lambda n: tqdm.tqdm(range(n))
except Exception:                                                    # pragma: no cover
    _progress = range


class GP_solver_2d_single(object):
    """u_xx + u_yy [+ u(u^2-1)] = f on a tensor grid.
    bvals: hstack(U[0,:],U[-1,:],U[:,0],U[:,-1]); X_col = (x_pos, y_pos); src_vals: N1 x N2;
    X_test = (x_test, y_test); u_test: M1 x M2; trick_paras: the config dict."""
    EQ_TYPES = ("poisson_2d", "allencahn_2d")

    def __init__(self, bvals, X_col, src_vals, jitter, X_test, u_test, trick_paras=None, fix_dict=None):
        self.bvals = np.asarray(_np(bvals), dtype=np.float64).reshape(-1)
        self.X_col = (np.asarray(_np(X_col[0]), dtype=np.float64).reshape(-1),
                      np.asarray(_np(X_col[1]), dtype=np.float64).reshape(-1))
        self.jitter = jitter
        self.Nb = self.bvals.size
        self.N1, self.N2 = self.X_col[0].size, self.X_col[1].size
        self.Nc = self.N1 * self.N2
        self.src_vals = np.asarray(_np(src_vals), dtype=np.float64).reshape(self.N1, self.N2)
        self.trick_paras = trick_paras
        self.lr = trick_paras["lr"]
        self.llk_weight = trick_paras["llk_weight"]
        kernel = trick_paras["kernel"]
        self.cov_func = (KERNELS[kernel] if isinstance(kernel, str) else kernel)()
        self.kernel_matrix = Kernel_matrix(self.jitter, self.cov_func)
        self.Xte = (as_dev(_np(X_test[0])).reshape(-1), as_dev(_np(X_test[1])).reshape(-1))
        self.ute = as_dev(_np(u_test))
        self.params = None
        self.pred_func = None
        self.eq_type = trick_paras["equation"].split("-")[0]
        assert self.eq_type in self.EQ_TYPES
        self.beta = float(trick_paras.get("beta", 1.0))
        self.core = SolverCore(2, type(self.cov_func).__name__, self._eq_name(), self.X_col[0], self.X_col[1],
                               self.src_vals, self.bvals, None, self.llk_weight, float(trick_paras["logdet"]),
                               self.beta, self.jitter, trick_paras["Q"],
                               force_general=int(trick_paras.get("force_general", 0)))
        print("equation is: ", self.trick_paras["equation"])
        print("kernel is:", self.cov_func.__class__.__name__)

    def _eq_name(self):
        return {"poisson_2d": "poisson", "allencahn_2d": "allencahn", "advection": "advection"}[self.eq_type]

    _deriv_order = 2

    # ---- secondary reference methods (not on the per-iteration path) ---------------------------
    def value_and_grad_kernel(self, params, key=None):
        """(K1, K2, K1inv_U, K2inv_Ut, U_xx, U_yy)  - model_GP_solver_2d.py:87-121."""
        U = as_dev(params["U"])
        th1, th2 = params["kernel_paras_1"], params["kernel_paras_2"]
        x, y = self.X_col
        K1 = self.cov_func.gram(x, x, th1, 0, self.jitter)
        K2 = self.cov_func.gram(y, y, th2, 0, self.jitter)
        K1inv_U = solve_spd(K1, U)
        K2inv_Ut = solve_spd(K2, U.T.contiguous())
        U_xx = dgemm(self.cov_func.gram(x, x, th1, self._deriv_order), K1inv_U)
        U_yy = dgemm(self.cov_func.gram(y, y, th2, self._deriv_order), K2inv_Ut).T
        return K1, K2, K1inv_U, K2inv_Ut, U_xx, U_yy

    def boundary_and_eq_gap(self, U, U_xx, U_yy):
        """model_GP_solver_2d.py:123-143."""
        U, U_xx, U_yy = as_dev(U), as_dev(U_xx), as_dev(U_yy)
        u_b = torch.cat((U[0, :], U[-1, :], U[:, 0], U[:, -1]))
        boundary_gap = torch.sum(torch.square(u_b - as_dev(self.bvals)))
        src = as_dev(self.src_vals)
        if self.eq_type == "poisson_2d":
            eq_gap = torch.sum(torch.square(U_xx + U_yy - src))
        elif self.eq_type == "allencahn_2d":
            eq_gap = torch.sum(torch.square(U_xx + U_yy + U * (U ** 2 - 1) - src))
        elif self.eq_type == "advection":
            eq_gap = torch.sum(torch.square(self.beta * U_xx + U_yy - src))
        else:
            raise NotImplementedError
        return boundary_gap, eq_gap

    # ---- the hot path --------------------------------------------------------------------------
    def loss_terms(self, params, key=None):
        """dict of device scalars: loss, logdet1, logdet2, quad, boundary_gap, eq_gap."""
        st = self.core.new_state(params)
        terms, _, _ = self.core.value_and_grad(st, forward_only=True)
        return dict(zip(("loss", "logdet1", "logdet2", "quad", "boundary_gap", "eq_gap"), terms[:6]))

    def loss(self, params, key=None):
        """-log joint  - model_GP_solver_2d.py:145-174."""
        return self.loss_terms(params)["loss"]

    def value_and_grad(self, params, key=None):
        """jax.value_and_grad(self.loss)(params, key): (loss, grads pytree)."""
        st = self.core.new_state(params)
        terms, gU, gs = self.core.value_and_grad(st)
        return terms[0], self.core.unpack_tree(gU, gs)

    def step(self, params, opt_state, key=None):
        """(params, opt_state, loss) - functional, as model_GP_solver_2d.py:176-183."""
        st = self.core.new_state(params, opt_state)
        self.core.step_inplace(st, self.lr)
        new_params = self.core.unpack_tree(st.U, st.small)
        new_opt = {"count": st.count.reshape(()).clone(), "mu": self.core.unpack_tree(st.mU, st.msmall),
                   "nu": self.core.unpack_tree(st.vU, st.vsmall)}
        return new_params, new_opt, st.terms[0].clone()

    def preds(self, params):
        """(U_pred on the test grid, None) - model_GP_solver_2d.py:185-220."""
        st = self.core.new_state(params)
        return self.core.predict(st, self.Xte[0], self.Xte[1]), None

    def compute_early_stopping(self, params, key=None):
        """boundary_gap / Nb + eq_gap / Nc - model_GP_solver_2d.py:222-233."""
        t = self.loss_terms(params)
        return t["boundary_gap"] / self.Nb + t["eq_gap"] / self.Nc

    def init_params(self):
        """model_GP_solver_2d.py:245-261."""
        Q, fs = self.trick_paras["Q"], self.trick_paras["freq_scale"]
        kp = lambda: {"log-w": np.log(1 / Q) * np.ones(Q), "log-ls": np.zeros(Q), "freq": np.linspace(0, 1, Q) * fs}
        return {"log_tau": 0.0, "log_v": 0.0, "kernel_paras_1": kp(), "kernel_paras_2": kp(),
                "U": np.zeros((self.N1, self.N2))}

    def train(self, nepoch, seed=0):
        """model_GP_solver_2d.py:235-352: same logging cadence (`i % (nepoch/20) == 0`), same
        log_dict keys, same return tuple.  The per-iteration work is one in-place gphm_step on a
        persistent device state; the host only synchronises at the 20 evaluation points."""
        early_stopping = {"flag": False, "epoch": self.trick_paras["nepoch"]}
        st = self.core.new_state(self.init_params())
        self.core.check_conditioning(st)
        log = {k: [] for k in ("loss_list", "err_list", "w_list_k1", "freq_list_k1", "ls_list_k1", "w_list_k2",
                               "freq_list_k2", "ls_list_k2", "epoch_list")}
        min_err = 2.0
        self.pred_func = self.preds
        for i in _progress(nepoch):
            self.core.step_inplace(st, self.lr)
            if i % (nepoch / 20) == 0:
                loss = float(st.terms[0])
                pred = self.core.predict(st, self.Xte[0], self.Xte[1])
                err = float(self.core.rel_l2(pred, self.ute))
                self.core.raise_on_bad_status()
                min_err = min(min_err, err)
                print("It ", i, "  loss = %g " % loss, " Relative L2 error", err, " min error", min_err)
                params = self.core.unpack_tree(st.U, st.small)
                log["loss_list"].append(np.log(loss) if loss > 1 else loss)
                log["err_list"].append(err)
                for a in ("1", "2"):
                    kp = params["kernel_paras_" + a]
                    log["w_list_k" + a].append(torch.exp(kp["log-w"]).cpu().numpy())
                    log["freq_list_k" + a].append(kp["freq"].cpu().numpy())
                    log["ls_list_k" + a].append(torch.exp(kp["log-ls"]).cpu().numpy())
                log["epoch_list"].append(i)
                terms, _, _ = self.core.value_and_grad(st, forward_only=True)
                criterion = float(terms[4]) / self.Nb + float(terms[5]) / self.Nc
                print("criterion = %g" % criterion)
                if self._early_stop_enabled and self.trick_paras["tol"] > 0 and criterion < self.trick_paras["tol"]:
                    print("early stop at epoch %d" % i)
                    early_stopping["flag"], early_stopping["epoch"] = True, i
                    break
        print("finish training ...")
        self.params = self.core.unpack_tree(st.U, st.small)
        self.state = st
        return log, early_stopping, min_err

    _early_stop_enabled = True


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


def _derivs(fn, args, wrt, order):
    args = [torch.as_tensor(a, dtype=DT).clone().requires_grad_(True) for a in args]
    out = fn(*args)
    for _ in range(order):
        (out,) = torch.autograd.grad(out.sum(), args[wrt], create_graph=True)
    return out.detach()


equation_dict = {
    "poisson_2d-sin_sin": lambda x, y: torch.sin(100 * x) * torch.sin(100 * y),
    "poisson_2d-sin_cos": lambda x, y: torch.sin(100 * x) * torch.cos(100 * y),
    "poisson_2d-sin_add_cos": lambda x, y: torch.sin(6 * x) * torch.cos(20 * x) + torch.sin(6 * y) * torch.cos(20 * y),
    "allencahn_2d-mix-sincos": lambda x, y: (torch.sin(x) + 0.1 * torch.sin(20 * x) + torch.cos(100 * x)) *
                                            (torch.sin(y) + 0.1 * torch.sin(20 * y) + torch.cos(100 * y)),
}


def get_source_val(u, x_pos, y_pos, equation_type):
    """Manufactured source on the collocation mesh, flattened row-major (model_GP_solver_2d.py:355-366)."""
    X, Y = torch.meshgrid(torch.as_tensor(x_pos, dtype=DT), torch.as_tensor(y_pos, dtype=DT), indexing="ij")
    lap = _derivs(u, [X, Y], 0, 2) + _derivs(u, [X, Y], 1, 2)
    if equation_type == "poisson_2d":
        return lap.reshape(-1).numpy()
    if equation_type == "allencahn_2d":
        uv = u(X, Y)
        return (lap + uv * (uv ** 2 - 1)).reshape(-1).numpy()
    raise NotImplementedError


def get_mesh_data(u, M1, M2, scale):
    """model_GP_solver_2d.py:369-374."""
    x_coor = np.linspace(0, 1, num=M1) * scale
    y_coor = np.linspace(0, 1, num=M2) * scale
    X, Y = torch.meshgrid(torch.as_tensor(x_coor), torch.as_tensor(y_coor), indexing="ij")
    return x_coor, y_coor, u(X, Y).numpy()


def get_boundary_vals(u_mesh):
    """model_GP_solver_2d.py:377-379."""
    return np.hstack((u_mesh[0, :], u_mesh[-1, :], u_mesh[:, 0], u_mesh[:, -1]))


def build_problem(trick_paras, M=300):
    """Grids, boundary values and source term exactly as test() builds them (:398-416)."""
    u = equation_dict[trick_paras["equation"]]
    eq_type = trick_paras["equation"].split("-")[0]
    scale, N = trick_paras["scale"], trick_paras["N_col"]
    x_te, y_te, u_test = get_mesh_data(u, M, M, scale)
    x_tr, y_tr, u_mh = get_mesh_data(u, N, N, scale)
    bvals = get_boundary_vals(u_mh)
    src = get_source_val(u, x_tr, y_tr, eq_type).reshape(x_tr.size, y_tr.size)
    return bvals, (x_tr, y_tr), src, (x_te, y_te), u_test


def test(trick_paras, solver_cls=None, problem=None):
    """model_GP_solver_2d.py:382-464."""
    solver_cls = solver_cls or GP_solver_2d_single
    bvals, X_col, src_vals, X_test, u_test = problem or build_problem(trick_paras)
    err_list, stop_list = [], []
    start = time.time()
    model = None
    for fold in range(trick_paras["num_fold"]):
        print("fold %d training" % fold)
        model = solver_cls(bvals, X_col, src_vals, 1e-6, X_test, u_test, trick_paras)
        log_dict, early_stopping, min_err = model.train(trick_paras["nepoch"], fold)
        err_list.append(min_err)
        stop_list.append(early_stopping["epoch"])
        if fold == 0:
            utils.store_model(model, log_dict, trick_paras)
    used = time.time() - start
    err_dict = {"mean": np.mean(err_list), "std": np.std(err_list), "err_list": err_list,
                "stop_epoch_mean": np.mean(stop_list), "used_time": used, "avg_time": used / trick_paras["num_fold"]}
    utils.wrirte_log(model, err_dict, trick_paras)
    print("finish writing log ...")
    return model, err_dict


EQUATIONS = ["poisson_2d-sin_cos", "poisson_2d-sin_sin", "poisson_2d-sin_add_cos", "allencahn_2d-mix-sincos"]


def make_config(equation, kernel, nepoch=None, allowed=None, config_dir="./config", suffix_fn=None):
    """The config-merging half of evals() (model_GP_solver_2d.py:472-508)."""
    assert equation in (allowed or EQUATIONS)
    config = configs.load_config(equation, config_dir)
    config["equation"] = equation
    config["init_u_trick"] = "zeros"
    config["kernel_extra"] = None
    config["scale"] = 2 * np.pi if config["scale"] == "2pi" else 1.0
    if nepoch is not None:
        config["nepoch"] = nepoch
    if kernel not in KERNELS:
        raise Exception("Invalid Kernel")
    config["kernel"] = KERNELS[kernel]
    print("equation: %s, kernel: %s, freq_scale: %d" % (config["equation"], config["kernel"].__name__, config["freq_scale"]))
    extra = suffix_fn(config) if suffix_fn else ""
    config["other_paras"] = config["other_paras"] + extra + "-Ncol-%d" % config["N_col"]
    return config


def evals(**kwargs):
    """model_GP_solver_2d.py:467-510:  evals(equation=..., kernel=..., nepoch=...)."""
    args = ExpConfig().parse(kwargs)
    return test(make_config(args.equation, args.kernel, args.nepoch))


def _main(evals_fn):
    """`python -m ... -equation=poisson_2d-sin_sin -kernel=Matern52_Cos_1d -nepoch=1000` (fire-style flags)."""
    import sys
    kw = {}
    for a in sys.argv[1:]:
        k, _, v = a.lstrip("-").partition("=")
        kw[k] = int(v) if v.lstrip("-").isdigit() else v
    evals_fn(**kw)


if __name__ == "__main__":
    _main(evals)
