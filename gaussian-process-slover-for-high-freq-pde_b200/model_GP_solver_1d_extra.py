"""Two-stage "extra GP" 1-D solver - mirror of the reference's model_GP_solver_1d_extra.py:
a first GP (spectral-mixture kernel) is trained up to `change_point * nepoch`, frozen, and a second
GP with `kernel_extra` (Matern52_1d, one component) is fitted to what is left of the equation.

  GP_solver_1d_extra.__init__                       model_GP_solver_1d_extra.py:31-55
  value_and_grad_kernel_extra / boundary_and_eq_gap_extra / loss_extra   :57-141
  step_extra / preds_extra / compute_early_stopping_extra                :143-199
  train                                             :201-339
  evals                                             :443-496

Stage 2 is again a 1-D log-joint: with the first GP frozen, loss_extra(params_extra) equals the
1-D loss of a problem whose source is  f - u_xx(stage 1),  whose boundary targets are
y - u[Xind],  and whose nonlinearity is evaluated on  u(stage 1) + u_extra  (the base-field hook of
libgphm, gphm_plan_set_base_field).  So it runs on the same CUDA kernels through a second plan.
"""
import copy
import ctypes
import time

import numpy as np
import torch

from . import _lib
from . import model_GP_solver_1d as m1d
from . import model_GP_solver_2d as m2d
from . import utils
from .infras.exp_config import ExpConfig
from .kernel_matrix import KERNELS, Kernel_matrix, as_dev
from .solver_core import SolverCore, dgemm, solve_spd


class GP_solver_1d_extra(m1d.GP_solver_1d_single):

    def __init__(self, Xind, y, X_col, src_col, jitter, X_test, Y_test, trick_paras=None, fix_dict=None):
        super().__init__(Xind, y, X_col, src_col, jitter, X_test, Y_test, trick_paras, fix_dict)
        kernel_extra = trick_paras["kernel_extra"]
        self.cov_func_extra = (KERNELS[kernel_extra] if isinstance(kernel_extra, str) else kernel_extra)()
        self.kernel_matrix_extra = Kernel_matrix(self.jitter, self.cov_func_extra)
        self.params = None
        self.params_extra = None
        self.core_extra = None
        print("using extra GP with kernel:", self.cov_func_extra.__class__.__name__)

    # ---- stage-2 plan: built once the first GP is frozen -----------------------------------------
    def freeze_first_stage(self, params):
        """self.params <- params (model_GP_solver_1d_extra.py:258-262) and build the stage-2 plan."""
        self.params = copy.deepcopy(params)
        _, _, u_xx = self.value_and_grad_kernel(self.params)
        u = as_dev(self.params["u"]).reshape(-1)
        src2 = as_dev(self.src_col) - u_xx.reshape(-1)
        y2 = as_dev(self.y) - u[torch.as_tensor(self.Xind, device=u.device)]
        self.core_extra = SolverCore(1, type(self.cov_func_extra).__name__,
                                     {"poisson_1d": "poisson", "allencahn_1d": "allencahn"}[self.eq_type],
                                     self.X_col.reshape(-1), None, src2.cpu().numpy(), y2.cpu().numpy(), self.Xind,
                                     self.llk_weight, float(self.trick_paras["logdet"]), 1.0, self.jitter, self._Q_extra())
        if self.eq_type == "allencahn_1d":
            base = np.ascontiguousarray(u.cpu().numpy())
            _lib.check(self.core_extra.lib.gphm_plan_set_base_field(self.core_extra.plan,
                                                                    base.ctypes.data_as(ctypes.c_void_p)),
                       "gphm_plan_set_base_field")
        self.core_extra._kp_keys = lambda: ("kernel_paras",)
        return self.core_extra

    def _Q_extra(self):
        return 1                                      # params_extra['kernel_paras'] has one component (:266-269)

    def init_params_extra(self, params):
        """model_GP_solver_1d_extra.py:263-271."""
        return {"log_tau": copy.deepcopy(params["log_tau"]), "log_v": 0.0,
                "kernel_paras": {"log-w": np.zeros(1), "log-ls": np.zeros(1)}, "u": np.zeros((self.N_con, 1))}

    @staticmethod
    def _with_freq(params_extra):
        """The stage-2 kernel has no 'freq' leaf in the reference; the packed layout carries a zero
        one whose gradient is exactly 0 (Adam leaves it untouched)."""
        kp = dict(params_extra["kernel_paras"])
        kp.setdefault("freq", np.zeros(np.asarray(_np(kp["log-w"])).size))
        out = dict(params_extra)
        out["kernel_paras"] = kp
        return out

    @staticmethod
    def _drop_freq(tree):
        kp = {k: v for k, v in tree["kernel_paras"].items() if k != "freq"}
        out = dict(tree)
        out["kernel_paras"] = kp
        return out

    def _state_extra(self, params_extra, opt_state=None):
        if self.core_extra is None:
            raise RuntimeError("freeze_first_stage(params) must run before the extra-GP stage")
        p = self._with_freq(params_extra)
        p["u"] = as_dev(_np(p["u"])).reshape(self.N_con, -1).sum(1, keepdim=True)       # sum over trick (:113)
        if opt_state is not None:
            opt_state = {"count": opt_state["count"], "mu": self._with_freq(opt_state["mu"]),
                         "nu": self._with_freq(opt_state["nu"])}
        return self.core_extra.new_state(p, opt_state)

    # ---- reference methods ---------------------------------------------------------------------
    def value_and_grad_kernel_extra(self, params_extra, key=None):
        """(K, Kinv_u, u_xx) of the extra GP - model_GP_solver_1d_extra.py:57-78."""
        u = as_dev(_np(params_extra["u"])).reshape(self.N_con, -1).sum(1, keepdim=True)
        th = self._with_freq(params_extra)["kernel_paras"]
        x = self.X_con.reshape(-1)
        K = self.cov_func_extra.gram(x, x, th, 0, self.jitter)
        Kinv_u = solve_spd(K, u)
        return K, Kinv_u, dgemm(self.cov_func_extra.gram(x, x, th, 2), Kinv_u)

    def boundary_and_eq_gap_extra(self, u, u_extra, u_xx, u_xx_extra):
        """model_GP_solver_1d_extra.py:80-105."""
        u, ue = as_dev(_np(u)).reshape(-1), as_dev(_np(u_extra)).reshape(-1)
        uxx, uxxe = as_dev(u_xx).reshape(-1), as_dev(u_xx_extra).reshape(-1)
        idx = torch.as_tensor(self.Xind, device=u.device)
        boundary_gap = torch.sum(torch.square(u[idx] + ue[idx] - as_dev(self.y)))
        src = as_dev(self.src_col)
        if self.eq_type == "poisson_1d":
            eq_gap = torch.sum(torch.square(uxx + uxxe - src))
        elif self.eq_type == "allencahn_1d":
            t = u + ue
            eq_gap = torch.sum(torch.square(uxx + uxxe + t * (t ** 2 - 1) - src))
        else:
            raise NotImplementedError
        return boundary_gap, eq_gap

    def loss_extra(self, params_extra, key=None):
        """model_GP_solver_1d_extra.py:107-141."""
        terms, _, _ = self.core_extra.value_and_grad(self._state_extra(params_extra), forward_only=True)
        return terms[0]

    def value_and_grad_extra(self, params_extra, key=None):
        st = self._state_extra(params_extra)
        terms, gU, gs = self.core_extra.value_and_grad(st)
        return terms[0], self._drop_freq(self.core_extra.unpack_tree(gU, gs))

    def step_extra(self, params_extra, opt_state, key=None):
        """(params_extra, opt_state, loss) - model_GP_solver_1d_extra.py:143-152."""
        st = self._state_extra(params_extra, opt_state)
        self.core_extra.step_inplace(st, self.lr)
        c = self.core_extra
        new_opt = {"count": st.count.reshape(()).clone(), "mu": self._drop_freq(c.unpack_tree(st.mU, st.msmall)),
                   "nu": self._drop_freq(c.unpack_tree(st.vU, st.vsmall))}
        return self._drop_freq(c.unpack_tree(st.U, st.small)), new_opt, st.terms[0].clone()

    def preds_extra(self, params_extra, Xte=None):
        """preds of both GPs added - model_GP_solver_1d_extra.py:154-184."""
        preds, _ = self.preds(self.params, Xte)
        xt = self.Xte if Xte is None else as_dev(_np(Xte)).reshape(-1)
        return preds + self.core_extra.predict(self._state_extra(params_extra), xt), None

    def compute_early_stopping_extra(self, params_extra, key=None):
        """model_GP_solver_1d_extra.py:186-199."""
        terms, _, _ = self.core_extra.value_and_grad(self._state_extra(params_extra), forward_only=True)
        return terms[4] / self.N + terms[5] / self.N_con

    def init_params(self):
        """model_GP_solver_1d_extra.py:211-224 (the two '*-matern' leaves there are never read)."""
        p = super().init_params()
        p["u"] = np.zeros((self.N_con, self.trick_paras.get("num_u_trick", 1)))
        return p

    def train(self, nepoch, seed=0):
        """model_GP_solver_1d_extra.py:201-339: same switch at `change_point`, logging cadence,
        early-stop rule and return tuple."""
        early_stopping = {"flag": False, "epoch": self.trick_paras["nepoch"]}
        error_increase_count, min_err, threshold = 0, 2.0, 1e-3
        st = self.core.new_state(m1d.GP_solver_1d_single.init_params(self))
        self.core.check_conditioning(st)
        st2 = None
        log = {k: [] for k in ("loss_list", "err_list", "w_list", "freq_list", "ls_list", "epoch_list")}
        change_point = int(nepoch * self.trick_paras["change_point"])
        params = None
        for i in m2d._progress(nepoch):
            if i <= change_point:
                self.core.step_inplace(st, self.lr)
                loss_t = st.terms
            else:
                self.core_extra.step_inplace(st2, self.lr)
                loss_t = st2.terms
            if i == change_point:
                print("start to train the extra matern kernel")
                params = self.core.unpack_tree(st.U, st.small)
                self.freeze_first_stage(params)
                st2 = self._state_extra(self.init_params_extra(params))
                self.core_extra.check_conditioning(st2)      # Matern52_1d, Q = 1: the plain-kernel case of the guard
            if i % (nepoch / 20) == 0:
                loss = float(loss_t[0])
                pred = self.core.predict(st, self.Xte)
                if i > change_point:
                    pred = pred + self.core_extra.predict(st2, self.Xte)
                err = float(self.core.rel_l2(pred, self.yte))
                (self.core_extra if i > change_point else self.core).raise_on_bad_status()
                if err < min_err:
                    min_err = err
                elif err - min_err > threshold:
                    error_increase_count += 1
                print("It ", i, "  loss = %g " % loss, " Relative L2 error", err, " min error", min_err)
                kp = self.core.unpack_tree(st.U, st.small)["kernel_paras"]
                log["loss_list"].append(np.log(loss) if loss > 1 else loss)
                log["err_list"].append(err)
                log["w_list"].append(torch.exp(kp["log-w"]).cpu().numpy())
                log["freq_list"].append(kp["freq"].cpu().numpy())
                log["ls_list"].append(torch.exp(kp["log-ls"]).cpu().numpy())
                log["epoch_list"].append(i)
                terms, _, _ = self.core.value_and_grad(st, forward_only=True)
                criterion = float(terms[4]) / self.N + float(terms[5]) / self.N_con
                print("criterion = %g" % criterion)
                if i > 0 and (criterion < self.trick_paras["tol"] or error_increase_count > 7):
                    print("early stop at epoch %d" % i)
                    early_stopping["flag"], early_stopping["epoch"] = True, i
                    break
        print("finish training ...")
        if self.params is None:
            self.params = self.core.unpack_tree(st.U, st.small)
        self.params_extra = None if st2 is None else self._drop_freq(self.core_extra.unpack_tree(st2.U, st2.small))
        self.state, self.state_extra = st, st2
        return log, early_stopping, min_err


_np = m2d._np
equation_dict = m1d.equation_dict
get_source_val = m1d.get_source_val


def test(trick_paras):
    """model_GP_solver_1d_extra.py:354-440."""
    Xind, y, X_col, src, X_test, Y_test = m1d.build_problem(trick_paras)
    err_list, stop_list = [], []
    start = time.time()
    model = None
    for fold in range(trick_paras["num_fold"]):
        print("fold %d training" % fold)
        model = GP_solver_1d_extra(Xind, y, X_col, src, 1e-6, X_test, Y_test, trick_paras)
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
    """model_GP_solver_1d_extra.py:443-496."""
    args = ExpConfig().parse(kwargs)
    config = m2d.make_config(args.equation, args.kernel, args.nepoch, allowed=m1d.EQUATIONS)
    config["kernel_extra"] = KERNELS["Matern52_1d"]       # extra GP kernel to speed up the convergence (:463)
    config.setdefault("change_point", 0.5)
    config["other_paras"] = config["other_paras"] + "change_point-%.1f" % config["change_point"] + "-extra-GP"   # :491-492
    return test(config)


if __name__ == "__main__":
    m2d._main(evals)
