"""1-D GP-HM solver (Poisson / Allen-Cahn) - mirror of the reference's model_GP_solver_1d.py:
  GP_solver_1d_single.__init__                :38-78
  value_and_grad_kernel / boundary_and_eq_gap :80-121
  loss / step / preds / compute_early_stopping:123-191
  train                                       :193-296
  get_source_val / test / evals               :299-447
The per-iteration core runs in libgphm (same kernels as the 2-D path with a single axis)."""
import time

import numpy as np
import torch

from . import model_GP_solver_2d as m2d
from . import utils
from .infras.exp_config import ExpConfig
from .kernel_matrix import DT, KERNELS, Kernel_matrix, as_dev
from .solver_core import SolverCore, dgemm, solve_spd

_np = m2d._np


class GP_solver_1d_single(object):
    """u_xx [+ u(u^2-1)] = f.  Xind: indices of X_col that are boundary points; y: their values;
    X_col: (N,1) collocation points; src_col: source at X_col."""

    def __init__(self, Xind, y, X_col, src_col, jitter, X_test, Y_test, trick_paras=None, fix_dict=None):
        self.Xind = np.asarray(_np(Xind)).reshape(-1).astype(np.int64)
        self.y = np.asarray(_np(y), dtype=np.float64).reshape(-1)
        self.X_col = np.asarray(_np(X_col), dtype=np.float64)
        self.src_col = np.asarray(_np(src_col), dtype=np.float64).reshape(-1)
        self.jitter = jitter
        self.X_con = self.X_col
        self.N = self.Xind.shape[0]
        self.N_con = self.X_con.shape[0]
        self.trick_paras = trick_paras
        self.lr = trick_paras["lr"]
        self.llk_weight = trick_paras["llk_weight"]
        kernel = trick_paras["kernel"]
        self.cov_func = (KERNELS[kernel] if isinstance(kernel, str) else kernel)()
        self.kernel_matrix = Kernel_matrix(self.jitter, self.cov_func)
        self.Xte = as_dev(_np(X_test)).reshape(-1)
        self.yte = as_dev(_np(Y_test)).reshape(-1, 1)
        self.params = None
        self.pred_func = None
        self.eq_type = trick_paras["equation"].split("-")[0]
        assert self.eq_type in ["poisson_1d", "allencahn_1d"]
        self.core = SolverCore(1, type(self.cov_func).__name__,
                               {"poisson_1d": "poisson", "allencahn_1d": "allencahn"}[self.eq_type],
                               self.X_col.reshape(-1), None, self.src_col, self.y, self.Xind, self.llk_weight,
                               float(trick_paras["logdet"]), 1.0, self.jitter, trick_paras["Q"],
                               force_general=int(trick_paras.get("force_general", 0)))
        print("equation is: ", self.trick_paras["equation"])
        print("kernel is:", self.cov_func.__class__.__name__)

    def value_and_grad_kernel(self, params, key=None):
        """(K, Kinv_u, u_xx) - model_GP_solver_1d.py:80-99."""
        u = as_dev(params["u"]).reshape(-1, 1)
        th = params["kernel_paras"]
        x = self.X_con.reshape(-1)
        K = self.cov_func.gram(x, x, th, 0, self.jitter)
        Kinv_u = solve_spd(K, u)
        u_xx = dgemm(self.cov_func.gram(x, x, th, 2), Kinv_u)
        return K, Kinv_u, u_xx

    def boundary_and_eq_gap(self, u, u_xx):
        """model_GP_solver_1d.py:101-121."""
        u, u_xx = as_dev(u).reshape(-1), as_dev(u_xx).reshape(-1)
        idx = torch.as_tensor(self.Xind, device=u.device)
        boundary_gap = torch.sum(torch.square(u[idx] - as_dev(self.y)))
        src = as_dev(self.src_col)
        if self.eq_type == "poisson_1d":
            eq_gap = torch.sum(torch.square(u_xx - src))
        elif self.eq_type == "allencahn_1d":
            eq_gap = torch.sum(torch.square(u_xx + u * (u ** 2 - 1) - src))
        else:
            raise NotImplementedError
        return boundary_gap, eq_gap

    def loss_terms(self, params, key=None):
        st = self.core.new_state(params)
        terms, _, _ = self.core.value_and_grad(st, forward_only=True)
        return dict(zip(("loss", "logdet1", "logdet2", "quad", "boundary_gap", "eq_gap"), terms[:6]))

    def loss(self, params, key=None):
        """model_GP_solver_1d.py:123-149."""
        return self.loss_terms(params)["loss"]

    def value_and_grad(self, params, key=None):
        st = self.core.new_state(params)
        terms, gU, gs = self.core.value_and_grad(st)
        return terms[0], self.core.unpack_tree(gU, gs)

    def step(self, params, opt_state, key=None):
        """(params, opt_state, loss) - model_GP_solver_1d.py:151-158."""
        st = self.core.new_state(params, opt_state)
        self.core.step_inplace(st, self.lr)
        new_opt = {"count": st.count.reshape(()).clone(), "mu": self.core.unpack_tree(st.mU, st.msmall),
                   "nu": self.core.unpack_tree(st.vU, st.vsmall)}
        return self.core.unpack_tree(st.U, st.small), new_opt, st.terms[0].clone()

    def preds(self, params, Xte=None):
        """(preds (M,1), K) - model_GP_solver_1d.py:160-180."""
        st = self.core.new_state(params)
        xt = self.Xte if Xte is None else as_dev(_np(Xte)).reshape(-1)
        x = self.X_con.reshape(-1)
        return self.core.predict(st, xt), self.cov_func.gram(x, x, params["kernel_paras"], 0, self.jitter)

    def compute_early_stopping(self, params, key=None):
        """model_GP_solver_1d.py:182-191."""
        t = self.loss_terms(params)
        return t["boundary_gap"] / self.N + t["eq_gap"] / self.N_con

    def init_params(self):
        """model_GP_solver_1d.py:203-213."""
        Q, fs = self.trick_paras["Q"], self.trick_paras["freq_scale"]
        return {"log_tau": 0.0, "log_v": 0.0,
                "kernel_paras": {"log-w": np.log(1 / Q) * np.ones(Q), "log-ls": np.zeros(Q),
                                 "freq": np.linspace(0, 1, Q) * fs},
                "u": np.zeros((self.N_con, 1))}

    def train(self, nepoch, seed=0):
        """model_GP_solver_1d.py:193-296 (same cadence, log_dict keys and return tuple)."""
        early_stopping = {"flag": False, "epoch": self.trick_paras["nepoch"]}
        st = self.core.new_state(self.init_params())
        self.core.check_conditioning(st)
        log = {k: [] for k in ("loss_list", "err_list", "w_list", "freq_list", "ls_list", "epoch_list")}
        min_err = 2.0
        self.pred_func = self.preds
        for i in m2d._progress(nepoch):
            self.core.step_inplace(st, self.lr)
            if i % (nepoch / 20) == 0:
                loss = float(st.terms[0])
                err = float(self.core.rel_l2(self.core.predict(st, self.Xte), self.yte))
                self.core.raise_on_bad_status()
                min_err = min(min_err, err)
                print("It ", i, "  loss = %g " % loss, " Relative L2 error", err, " min error", min_err)
                kp = self.core.unpack_tree(st.U, st.small)["kernel_paras"]
                log["loss_list"].append(np.log(loss) if loss > 1 else loss)
                log["err_list"].append(err)
                log["w_list"].append(torch.exp(kp["log-w"]).cpu().numpy())
                log["freq_list"].append(kp["freq"].cpu().numpy())
                log["ls_list"].append(torch.exp(kp["log-ls"]).cpu().numpy())
                log["epoch_list"].append(i)
                terms, _, _ = self.core.value_and_grad(st, forward_only=True)
                print("criterion = %g" % (float(terms[4]) / self.N + float(terms[5]) / self.N_con))
        print("finish training ...")
        self.params = self.core.unpack_tree(st.U, st.small)
        self.state = st
        return log, early_stopping, min_err


equation_dict = {
    "poisson_1d-mix_sin": lambda x: torch.sin(x) + 0.1 * torch.sin(20 * x) + 0.05 * torch.sin(100 * x),
    "poisson_1d-single_sin": lambda x: torch.sin(100 * x),
    "poisson_1d-sin_cos": lambda x: torch.sin(6 * x) * torch.cos(100 * x),
    "poisson_1d-x_time_sinx": lambda x: x * torch.sin(200 * x),
    "poisson_1d-x2_add_sinx": lambda x: torch.sin(500 * x) - 2 * (x - 0.5) ** 2,
    "allencahn_1d-sin_cos": lambda x: torch.sin(6 * x) * torch.cos(100 * x),
    "allencahn_1d-single_sin": lambda x: torch.sin(100 * x),
    "poisson_1d-x_time_sinx_scale": lambda x: x * torch.sin(200 * x * np.pi),
}

EQUATIONS = ["poisson_1d-mix_sin", "poisson_1d-single_sin", "poisson_1d-sin_cos", "poisson_1d-x_time_sinx",
             "poisson_1d-x2_add_sinx", "allencahn_1d-sin_cos", "allencahn_1d-single_sin"]


def get_source_val(u, x_vec, equation_type):
    """model_GP_solver_1d.py:299-307."""
    x = torch.as_tensor(x_vec, dtype=DT)
    uxx = m2d._derivs(u, [x], 0, 2)
    if equation_type == "poisson_1d":
        return uxx.numpy()
    if equation_type == "allencahn_1d":
        return (uxx + u(x) * (u(x) ** 2 - 1)).numpy()
    raise NotImplementedError


def build_problem(trick_paras, M=300):
    """model_GP_solver_1d.py:334-354."""
    u = equation_dict[trick_paras["equation"]]
    scale, N_col = trick_paras["scale"], trick_paras["N_col"]
    X_test = np.linspace(0, 1, num=M).reshape(-1, 1) * scale
    Y_test = u(torch.as_tensor(X_test)).numpy()
    X_col = np.linspace(0, 1, num=N_col).reshape(-1, 1) * scale
    Xind = np.array([0, X_col.shape[0] - 1])
    y = u(torch.as_tensor(X_col[Xind].reshape(-1))).numpy()
    src = get_source_val(u, X_col.reshape(-1), trick_paras["equation"].split("-")[0])
    return Xind, y, X_col, src, X_test, Y_test


def test(trick_paras):
    """model_GP_solver_1d.py:310-391."""
    Xind, y, X_col, src, X_test, Y_test = build_problem(trick_paras)
    err_list, stop_list = [], []
    start = time.time()
    model = None
    for fold in range(trick_paras["num_fold"]):
        print("fold %d training" % fold)
        model = GP_solver_1d_single(Xind, y, X_col, src, 1e-6, X_test, Y_test, trick_paras)
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


def evals(**kwargs):
    """model_GP_solver_1d.py:396-447."""
    args = ExpConfig().parse(kwargs)
    return test(m2d.make_config(args.equation, args.kernel, args.nepoch, allowed=EQUATIONS))


if __name__ == "__main__":
    m2d._main(evals)
