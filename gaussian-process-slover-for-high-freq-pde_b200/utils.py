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
