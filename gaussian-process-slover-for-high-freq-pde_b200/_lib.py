"""ctypes binding of libgphm.so (the C-ABI declared in include/gphm.h).

There is no CPU fallback: if the shared library is missing or a call fails this module raises.
Device memory is owned by torch CUDA tensors; only raw pointers and sizes cross the boundary.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgphm.so")

KERNEL_IDS = {"SE_Cos_1d": 0, "Matern52_Cos_1d": 1, "Matern52_1d": 2, "SE_1d": 3}
EQ_IDS = {"poisson": 0, "allencahn": 1, "advection": 2}
FORWARD_ONLY = 1
NOT_SPD, NONFINITE, ILL_CONDITIONED = 1, 2, 3


class ProblemDesc(ctypes.Structure):
    """gphm_problem_desc (include/gphm.h)."""
    _fields_ = [("dim", c_int), ("kernel_id", c_int), ("eq_type", c_int), ("n1", c_int), ("n2", c_int),
                ("Q", c_int), ("nb", c_int), ("force_general", c_int), ("llk_weight", c_double),
                ("logdet", c_double), ("beta", c_double), ("jitter", c_double)]


_SIGS = {
    "gphm_version": (c_int, []),
    "gphm_last_error": (c_char_p, []),
    "gphm_launch_count": (c_longlong, []),
    "gphm_profile_start": (c_int, []),
    "gphm_profile_stop": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "gphm_gram": (c_int, [c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "gphm_kappa_pairs": (c_int, [c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_void_p, c_void_p]),
    "gphm_dgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_double, c_void_p, c_int, c_void_p, c_int, c_double,
                           c_void_p, c_int, c_void_p]),
    "gphm_ozaki_work_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gphm_ozaki_error_factor": (c_double, [c_int, c_int]),
    "gphm_ozaki_dgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_double, c_void_p, c_int, c_void_p, c_int, c_double,
                                 c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "gphm_potrf_work_bytes": (c_size_t, [c_int]),
    "gphm_potrf_inv": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gphm_workspace_bytes": (c_size_t, [POINTER(ProblemDesc)]),
    "gphm_plan_create": (c_int, [POINTER(ProblemDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_size_t, POINTER(c_void_p)]),
    "gphm_plan_destroy": (None, [c_void_p]),
    "gphm_plan_status": (c_int, [c_void_p, POINTER(c_int), c_void_p]),
    "gphm_plan_uses_toeplitz": (c_int, [c_void_p, c_int]),
    "gphm_plan_use_cholesky": (c_int, [c_void_p]),
    "gphm_plan_set_base_field": (c_int, [c_void_p, c_void_p]),
    "gphm_logjoint_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "gphm_adam_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_double, c_void_p]),
    "gphm_adam_update_inc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_double, c_void_p]),
    "gphm_step": (c_int, [c_void_p] * 8 + [c_double, c_void_p, c_void_p]),
    "gphm_step_host": (c_int, [c_void_p] * 8 + [c_double, c_void_p, c_void_p]),
    "gphm_step_host_params": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_double, c_void_p, c_void_p]),
    "gphm_predict_work_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "gphm_predict": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "gphm_rel_l2_work_bytes": (c_size_t, []),
    "gphm_rel_l2": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "gphm_plan_factor": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "gphm_apply_kinv": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "gphm_apply_kinv_rows_refined": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "gphm_plan_matrix": (c_void_p, [c_void_p, c_int, c_int]),
    "gphm_plan_logdet": (c_int, [c_void_p, c_void_p, c_void_p]),
    "gphm_mg_residual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                 c_void_p]),
    "gphm_mg_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gphm_mg_pack_transposed": (c_int, [c_void_p, c_int, c_int, c_int, c_size_t, c_void_p, c_void_p]),
    "gphm_mg_unpack_segments": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gphm_mg_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "gphm_mg_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "gphm_mg_peer_close": (c_int, [c_void_p]),
    "gphm_mg_peer_free": (c_int, [c_void_p]),
    "gphm_mg_peer_exchange": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, ctypes.c_ulonglong, c_size_t,
                                      c_void_p, c_void_p]),
    "gphm_mg_boundary": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "gphm_mg_grad_u": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                               c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gphm_mg_theta_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gphm_plan_uses_fft": (c_int, [c_void_p, c_int]),
    "gphm_plan_uses_gs": (c_int, [c_void_p, c_int]),
    "gphm_toeplitz_work_bytes": (c_size_t, [c_int, c_int]),
    "gphm_toeplitz_solve": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "gphm_mg_toeplitz_apply": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "gphm_transpose": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "gphm_mg_theta_grad_fft": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double,
                                       c_void_p, c_void_p, c_void_p]),
    "gphm_mg_toeplitz_rows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_double, c_double, c_void_p, c_void_p, c_int,
                                      c_void_p]),
    "gphm_mg_theta_grad_pairs": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_void_p,
                                         c_void_p, c_void_p]),
    "gphm_mg_theta_grad_pairs_both": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_double,
                                              c_double, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "gphm_lincomb": (c_int, [c_void_p, c_double, c_void_p, c_double, c_void_p, c_size_t, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGS))
_lib = None


class GphmError(RuntimeError):
    pass


def load():
    """Load libgphm.so once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GphmError("libgphm.so not found at %s - build it with `python %s` (nvcc, sm_100a); "
                        "there is no CPU fallback" % (LIB_PATH, os.path.join(HERE, "build.py")))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().gphm_last_error().decode("utf-8", "replace")


def check(status, what):
    if status != 0:
        msg = last_error()
        if status == -1:
            if msg.startswith("Invalid Kernel"):
                raise Exception("Invalid Kernel")          # model_GP_solver_2d.py:502
            raise ValueError("%s: %s" % (what, msg))
        if status == -3:
            raise MemoryError("%s: %s" % (what, msg))
        raise GphmError("%s failed (status %d): %s" % (what, status, msg))


def ptr(t):
    """Raw device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
