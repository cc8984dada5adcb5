"""gphm_b200: B200-native (sm_100a) log-joint + gradient + Adam path of the GP-HM PDE solver
(reference: xuangu-fang/Gaussian-Process-Slover-for-High-Freq-PDE), behind the reference's own
Python interface.  Host code is thin: every numerical kernel lives in libgphm.so (csrc/), reached
through the C-ABI of include/gphm.h via ctypes.  Import as

    import importlib; gphm = importlib.import_module("gaussian-process-slover-for-high-freq-pde_b200")
or  import gphm_b200 as gphm        (alias module at the repo root)
"""
from . import _lib, configs, kernel_matrix, solver_core, utils          # noqa: F401
from . import model_GP_solver_1d, model_GP_solver_1d_extra, model_GP_solver_2d, model_GP_solver_advection   # noqa: F401
from .kernel_matrix import (Kernel_1d, Kernel_matrix, Matern52_1d, Matern52_Cos_1d, SE_1d,   # noqa: F401
                            SE_Cos_1d)
from .model_GP_solver_1d import GP_solver_1d_single                      # noqa: F401
from .model_GP_solver_1d_extra import GP_solver_1d_extra                 # noqa: F401
from .model_GP_solver_2d import GP_solver_2d_single                      # noqa: F401
from .model_GP_solver_advection import GP_solver_2d_single_advection     # noqa: F401

__version__ = "0.1.0"
