"""Result formats of the reference (utils.py:550-619): directory scheme, save name, the
(params, log_dict, trick_paras) pickle and the appended log.txt block, so that downstream readers
of the reference's result_log/ tree keep working.  Arrays are stored as numpy (the reference
stores jax arrays); kernel classes inside trick_paras are stored by name."""
import os
import pickle

import numpy as np
import torch


def _to_numpy(x):
    if isinstance(x, dict):
        return {k: _to_numpy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_to_numpy(v) for v in x]
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, type):
        return x.__name__
    if callable(x):
        return getattr(x, "__name__", repr(x))
    return x


def get_prefix(model, trick_paras):
    kname = model.cov_func.__class__.__name__
    if trick_paras.get("kernel_extra") is not None:            # utils.py:550-556
        kname += "-extra-" + model.cov_func_extra.__class__.__name__
    prefix = ("result_log/" + trick_paras["equation"] + "/kernel_" + kname +
              "/epoch_" + str(trick_paras["nepoch"]) + "/Q" + str(trick_paras["Q"]) + "/")
    os.makedirs(prefix, exist_ok=True)
    return prefix


def _head(trick_paras, sep):
    return ("llk_weight-%.1f%snu-%d-Q-%d-epoch-%d-lr-%.4f-freqscale=%d-logdet-%d" % (
        trick_paras["llk_weight"], sep, trick_paras["num_u_trick"], trick_paras["Q"], trick_paras["nepoch"],
        trick_paras["lr"], trick_paras["freq_scale"], trick_paras["logdet"]) + trick_paras["other_paras"])


def get_save_name(trick_paras):
    return _head(trick_paras, "-")


def store_model(model, log_dict, trick_paras):
    path = get_prefix(model, trick_paras) + get_save_name(trick_paras) + ".pkl"
    with open(path, "wb") as f:
        if trick_paras.get("kernel_extra") is not None:        # (params, params_extra, log_dict, trick_paras), utils.py:587-589
            data = (_to_numpy(model.params), _to_numpy(model.params_extra), _to_numpy(log_dict), _to_numpy(trick_paras))
        else:
            data = (_to_numpy(model.params), _to_numpy(log_dict), _to_numpy(trick_paras))
        pickle.dump(data, f)
    print("save model, log_dict, trick_paras to ", path)
    return path


def wrirte_log(model, err_dict, trick_paras):
    path = get_prefix(model, trick_paras) + "log.txt"
    with open(path, "a+") as f:
        f.write(_head(trick_paras, "--") + "\n")
        f.write("err_mean: %.4f, err_std: %.4f, used_time: %.4f, avg_time: %.4f, avg_epochs %d \n" % (
            err_dict["mean"], err_dict["std"], err_dict["used_time"], err_dict["avg_time"],
            err_dict["stop_epoch_mean"]))
        f.write("err_list: " + str([float(np.asarray(e)) for e in err_dict["err_list"]]) + "\n\n\n")
    print("write log to ", path)
    return path


# ---- model rebuilders for the notebooks (utils.py:622-837): (params[, params_extra], trick_paras) of a stored pickle
# -> a live solver object + its prediction on the (optionally finer) test grid -------------------------------------------
def load_model(path):
    """The tuple store_model wrote: (params, log_dict, trick_paras) or (params, params_extra, log_dict, trick_paras)."""
    with open(path, "rb") as f:
        return pickle.load(f)


def _with_scale(trick_paras):
    tp = dict(trick_paras)
    if "scale" not in tp and "x_scale" in tp:            # the reference's notebooks pass 'x_scale' (utils.py:647)
        tp["scale"] = tp["x_scale"]
    return tp


def get_model_1d(params, trick_paras, new_test=False):
    """utils.py:622-677 -> (model, preds (M,1), Xtr)."""
    from . import model_GP_solver_1d as m1d
    tp = _with_scale(trick_paras)
    Xind, y, X_col, src, X_test, Y_test = m1d.build_problem(tp, M=new_test if new_test else 300)
    model = m1d.GP_solver_1d_single(Xind, y, X_col, src, 1e-6, X_test, Y_test, tp)
    model.params = params
    preds, _ = model.preds(params, model.Xte)
    return model, preds, np.asarray(X_col)[Xind]


def get_model_1d_extra(params, params_extra, trick_paras, new_test=False):
    """utils.py:679-740 -> (model, preds of the two-stage solver, Xtr)."""
    from . import model_GP_solver_1d as m1d, model_GP_solver_1d_extra as mex
    tp = _with_scale(trick_paras)
    Xind, y, X_col, src, X_test, Y_test = m1d.build_problem(tp, M=new_test if new_test else 300)
    model = mex.GP_solver_1d_extra(Xind, y, X_col, src, 1e-6, X_test, Y_test, tp)
    model.freeze_first_stage(params)
    model.params_extra = params_extra
    preds, _ = model.preds_extra(params_extra, model.Xte)
    return model, preds, np.asarray(X_col)[Xind]


def get_model_2d(params, trick_paras, new_test=False):
    """utils.py:742-790 -> (model, preds (M,M))."""
    from . import model_GP_solver_2d as m2d
    tp = _with_scale(trick_paras)
    prob = m2d.build_problem(tp, M=new_test if new_test else 300)
    model = m2d.GP_solver_2d_single(prob[0], prob[1], prob[2], 1e-6, prob[3], prob[4], tp)
    model.params = params
    preds, _ = model.preds(params)
    return model, preds


def get_model_2d_advection(params, trick_paras, new_test=False):
    """utils.py:792-837 -> (model, preds (M,M))."""
    from . import model_GP_solver_advection as madv
    tp = _with_scale(trick_paras)
    prob = madv.build_problem(tp, M=new_test if new_test else 300)
    model = madv.GP_solver_2d_single_advection(prob[0], prob[1], prob[2], 1e-6, prob[3], prob[4], tp)
    model.params = params
    preds, _ = model.preds(params)
    return model, preds
