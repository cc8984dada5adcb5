"""Golden vectors produced by EXECUTING the reference's own, unmodified solver sources
(/root/reference/code/{kernel_matrix,model_GP_solver_1d,model_GP_solver_2d,model_GP_solver_advection}.py)
with torch standing in for the jax / optax API (tests/golden/ref_exec_shim/: JAX cannot be installed here).

    python tests/golden/make_ref_exec_golden.py            # writes tests/golden/ref_exec.npz  (needs /root/reference)

For every case the reference classes are constructed exactly as their test() functions do and the reference
methods `loss`, `jax.value_and_grad(loss)`, `step` (x2) and `preds` are called on a deterministic non-degenerate
state; inputs and outputs go into the fixture.  tests/test_oracle_ref_exec.py then holds the oracle to them - this
pins the kernels / equations the two shipped result logs do not cover (SE_Cos_1d, Matern52_1d, SE_1d, Allen-Cahn,
advection).  Only array data is stored; nothing from the reference is copied."""
import contextlib
import io
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/code"

CASES_2D = [(eq, k) for eq in ("poisson_2d-sin_add_cos", "allencahn_2d-mix-sincos")
            for k in ("Matern52_Cos_1d", "SE_Cos_1d", "Matern52_1d", "SE_1d")]
CASES_ADV = [("advection-sin", "SE_Cos_1d"), ("advection-sin", "Matern52_Cos_1d")]
CASES_1D = [(eq, k) for eq in ("poisson_1d-sin_cos", "allencahn_1d-single_sin")
            for k in ("Matern52_Cos_1d", "SE_Cos_1d", "Matern52_1d", "SE_1d")]
# two size sets: "s" (tiny, exercises every code path of the oracle cheaply) and "g" (the sizes the GPU parity tests use)
SIZES = {"s": (12, 10, 20, 4), "g": (40, 33, 60, 6)}            # N1, N2, N (1-D), Q
FS, LR, M_TEST = 5.0, 0.01, 7


def _np(t):
    return {k: _np(v) for k, v in t.items()} if isinstance(t, dict) else torch.as_tensor(t).detach().numpy().copy()


def _flat(prefix, tree, out):
    for k, v in tree.items():
        if isinstance(v, dict):
            _flat(prefix + k + "/", v, out)
        else:
            out[prefix + k] = np.asarray(v, dtype=np.float64)


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "ref_exec_shim"))
    sys.path.insert(0, ROOT)
    os.chdir("/tmp")                                  # the reference writes relative paths; nothing is written here
    import jax
    import kernel_matrix as KM
    import model_GP_solver_2d as M2               # first: utils.py imports the solver modules circularly
    import model_GP_solver_1d as M1
    import model_GP_solver_advection as MA
    import model_GP_solver_1d_extra as ME
    from oracle import gphm_oracle as O               # problem data and states only (inputs), not results
    out = {}
    quiet = contextlib.redirect_stdout(io.StringIO())

    def run(tag, model, params, one_d=False):
        val, grads = jax.value_and_grad(model.loss)(params, 0)
        rec = {"loss": np.float64(val)}
        _flat("grad/", _np(grads), rec)
        _flat("params0/", _np(params), rec)
        opt = model.optimizer.init(params)
        pr, losses = params, []
        for _ in range(2):
            pr, opt, l = model.step(pr, opt, 0)
            losses.append(float(l))
        rec["step_losses"] = np.asarray(losses)
        _flat("params2/", _np(pr), rec)
        pred = model.preds(pr, model.Xte)[0] if one_d else model.preds(pr)[0]
        rec["pred2"] = np.asarray(torch.as_tensor(pred).detach().numpy(), dtype=np.float64)
        for k, v in rec.items():
            out[tag + "|" + k] = v
        print(tag, "loss %.12e" % float(val), flush=True)

    for sz, (N1, N2, N1D, Q) in SIZES.items():
        for eq, kname in CASES_2D + CASES_ADV:
            adv = eq.startswith("advection")
            beta = 7.0 if adv else 1.0
            p, (xt, yt), ut = O.make_problem_2d(eq, kname, N1, 1.0 if adv else 2 * math.pi, beta=beta, M=M_TEST, N2=N2,
                                                llk_weight=500.0 if adv else 200.0)
            tp = {"lr": LR, "llk_weight": p.llk_weight, "kernel": getattr(KM, kname), "equation": eq, "logdet": True, "Q": Q,
                  "freq_scale": FS, "nepoch": 2, "beta": beta}
            cls = MA.GP_solver_2d_single_advection if adv else M2.GP_solver_2d_single
            with quiet:
                model = cls(p.bvals.numpy(), (p.x.numpy(), p.y.numpy()), p.src.numpy().reshape(N1, N2), 1e-6,
                            (xt.numpy(), yt.numpy()), ut.numpy(), tp)
            model.bvals, model.src_vals = torch.as_tensor(model.bvals), torch.as_tensor(model.src_vals)   # data only
            tag = "%s2d|%s|%s" % (sz, eq, kname)
            out[tag + "|beta"] = np.float64(beta)
            out[tag + "|src"], out[tag + "|bvals"] = p.src.numpy().reshape(N1, N2), p.bvals.numpy()
            run(tag, model, O.state_S1(p, Q=Q, freq_scale=FS))

        for eq, kname in CASES_1D:
            p, xte, yte = O.make_problem_1d(eq, kname, N1D, 2 * math.pi, M=M_TEST)
            tp = {"lr": LR, "llk_weight": p.llk_weight, "kernel": getattr(KM, kname), "equation": eq, "logdet": True, "Q": Q,
                  "freq_scale": FS, "nepoch": 2}
            with quiet:
                model = M1.GP_solver_1d_single(p.xind.numpy(), p.yb.numpy(), p.x.numpy().reshape(-1, 1),
                                               p.src.numpy().reshape(-1, 1), 1e-6, xte.numpy().reshape(-1, 1),
                                               yte.numpy().reshape(-1, 1), tp)
            model.y, model.src_col = torch.as_tensor(model.y), torch.as_tensor(model.src_col)
            tag = "%s1d|%s|%s" % (sz, eq, kname)
            out[tag + "|src"], out[tag + "|yb"] = p.src.numpy(), p.yb.numpy()
            run(tag, model, state_1d(O, p, N1D, Q), one_d=True)

    # two-stage extra-GP solver (model_GP_solver_1d_extra.py:107-152): loss_extra with the first GP frozen
    for eq in ("poisson_1d-sin_cos", "allencahn_1d-single_sin"):
        N1D, Q = SIZES["g"][2], SIZES["g"][3]
        p, xte, yte = O.make_problem_1d(eq, "SE_Cos_1d", N1D, 2 * math.pi, M=M_TEST)
        tp = {"lr": LR, "llk_weight": p.llk_weight, "kernel": KM.SE_Cos_1d, "kernel_extra": KM.Matern52_1d, "equation": eq,
              "logdet": True, "Q": Q, "freq_scale": FS, "nepoch": 2}
        with quiet:
            model = ME.GP_solver_1d_extra(p.xind.numpy(), p.yb.numpy(), p.x.numpy().reshape(-1, 1), p.src.numpy().reshape(-1, 1),
                                          1e-6, xte.numpy().reshape(-1, 1), yte.numpy().reshape(-1, 1), tp)
        model.y, model.src_col = torch.as_tensor(model.y), torch.as_tensor(model.src_col)
        model.params = state_1d(O, p, N1D, Q)                                     # the frozen first stage
        pe = {"u": (0.2 * torch.cos(2 * p.x) - 0.1 * torch.sin(p.x)).reshape(-1, 1),
              "kernel_paras": {"log-w": torch.tensor([0.1], dtype=torch.float64), "log-ls": torch.tensor([-0.3], dtype=torch.float64)},
              "log_tau": torch.tensor(0.2, dtype=torch.float64), "log_v": torch.tensor(-0.1, dtype=torch.float64)}
        tag = "x1d|%s|SE_Cos_1d+Matern52_1d" % eq
        val, grads = jax.value_and_grad(model.loss_extra)(pe, 0)
        rec = {"loss": np.float64(val), "src": p.src.numpy(), "yb": p.yb.numpy()}
        _flat("grad/", _np(grads), rec)
        _flat("params0/", _np(pe), rec)
        _flat("first/", _np(model.params), rec)
        opt, pr, losses = model.optimizer_extra.init(pe), pe, []
        for _ in range(2):
            pr, opt, l = model.step_extra(pr, opt, 0)
            losses.append(float(l))
        rec["step_losses"] = np.asarray(losses)
        _flat("params2/", _np(pr), rec)
        rec["pred2"] = np.asarray(torch.as_tensor(model.preds_extra(pr, model.Xte)[0]).detach().numpy(), dtype=np.float64)
        for k, v in rec.items():
            out[tag + "|" + k] = v
        print(tag, "loss %.12e" % float(val), flush=True)

    np.savez_compressed(os.path.join(HERE, "ref_exec.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_exec.npz"), len(out), "arrays")


def state_1d(O, p, n, Q):
    """Deterministic non-degenerate 1-D state (the 1-D analogue of the oracle's S1 recipe)."""
    params = O.init_params_1d(n, Q, FS)
    q = torch.arange(Q, dtype=torch.float64)
    params["u"] = (0.5 * torch.sin(3 * p.x) + 0.1 * torch.cos(7 * p.x)).reshape(-1, 1)
    params["kernel_paras"] = {"log-w": math.log(1.0 / Q) - 0.05 * torch.cos(q), "log-ls": 0.1 * torch.sin(q),
                              "freq": FS * q / (Q - 1) + 0.05 * torch.sin(2 * q)}
    params["log_tau"], params["log_v"] = torch.tensor(0.3, dtype=torch.float64), torch.tensor(-0.2, dtype=torch.float64)
    return params


if __name__ == "__main__":
    main()
