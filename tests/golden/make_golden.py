"""Convert the reference's two shipped golden runs into plain .npz + JSON fixtures.

Run ONCE in the builder container (needs /root/reference, which does not exist on the
GPU box):    python tests/golden/make_golden.py

Source artefacts (reference, read-only):
  code/result_log/poisson_1d-single_sin/kernel_Matern52_Cos_1d/epoch_100/Q30/*.pkl
  code/result_log/poisson_2d-sin_sin/kernel_Matern52_Cos_1d/epoch_100/Q30/*.pkl
Tuple layout `(params, log_dict, trick_paras)` is what code/utils.py:585-595 pickles.

Nothing inside the pickle is executed: the unpickler whitelists numpy array
reconstruction, maps jax's array reconstructor onto the wrapped numpy array, and turns
every other global (kernel class, init function inside trick_paras) into an inert string.
Tests only ever read the .npz/.json written here.
"""
import glob
import json
import os
import pickle

import numpy as np

REF = "/root/reference/code/result_log"
OUT = os.path.dirname(os.path.abspath(__file__))


class _Stub:
    def __init__(self, name):
        self.name = name

    def __call__(self, *a, **k):
        return "<stub %s>" % self.name

    def __repr__(self):
        return "<stub %s>" % self.name


class SafeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == ("jax._src.array", "_reconstruct_array"):
            def rebuild(fun, args, arr_state, aval_state):
                arr = fun(*args)
                arr.__setstate__(arr_state)
                return arr
            return rebuild
        if module in ("numpy.core.multiarray", "numpy._core.multiarray") and name in ("_reconstruct", "scalar"):
            import numpy._core.multiarray as m
            return getattr(m, name)
        if (module, name) in (("numpy", "ndarray"), ("numpy", "dtype")):
            return getattr(np, name)
        return _Stub(module + "." + name)


def to_plain(x):
    if isinstance(x, dict):
        return {k: to_plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [to_plain(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    if isinstance(x, (str, int, float, bool)) or x is None:
        return x
    return repr(x)


def convert(tag, pattern, two_d):
    (path,) = glob.glob(pattern)
    with open(path, "rb") as f:
        params, log_dict, trick = SafeUnpickler(f).load()
    arrays = {}
    arrays["log_tau"] = np.asarray(params["log_tau"], dtype=np.float64)
    arrays["log_v"] = np.asarray(params["log_v"], dtype=np.float64)
    if two_d:
        arrays["U"] = np.asarray(params["U"], dtype=np.float64)
        for ax in ("1", "2"):
            for leaf in ("log-w", "log-ls", "freq"):
                arrays["kp%s_%s" % (ax, leaf)] = np.asarray(params["kernel_paras_" + ax][leaf], dtype=np.float64)
    else:
        arrays["u"] = np.asarray(params["u"], dtype=np.float64)
        for leaf in ("log-w", "log-ls", "freq"):
            arrays["kp_%s" % leaf] = np.asarray(params["kernel_paras"][leaf], dtype=np.float64)
    for k, v in log_dict.items():
        arrays["log_" + k] = np.asarray([np.asarray(e, dtype=np.float64) for e in v], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **arrays)
    meta = {"source": os.path.relpath(path, "/root/reference"),
            "trick_paras": to_plain(trick),
            "log_keys": sorted(log_dict.keys())}
    with open(os.path.join(OUT, tag + ".json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(tag, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    convert("poisson_1d_single_sin_matern52cos_e100",
            REF + "/poisson_1d-single_sin/kernel_Matern52_Cos_1d/epoch_100/Q30/*.pkl", False)
    convert("poisson_2d_sin_sin_matern52cos_e100",
            REF + "/poisson_2d-sin_sin/kernel_Matern52_Cos_1d/epoch_100/Q30/*.pkl", True)
