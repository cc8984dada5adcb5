"""jax.ffi host side of experimental/ffi_shim.cc - the XLA-FFI layer BASELINE's north_star names.

Importing this module needs JAX >= 0.4.31 and the built `libgphm_ffi.so`; neither exists in the
image this repo is developed in (no jax wheel, no XLA FFI headers), so this path is UNTESTED here and
never imported by the package, the tests or the bench - the ctypes host (`solver_core.py`) is the
tested one.  It raises, never falls back, when a piece is missing.

    step = make_step(plan_handle, lr)           # plan_handle: SolverCore.plan.value (int)
    U, small, mU, vU, msmall, vsmall, count, terms = step(U, small, mU, vU, msmall, vsmall, count)

is the functional `GP_solver_2d_single.step` (model_GP_solver_2d.py:176-183) on the packed layout of
include/gphm.h; the seven state arrays are donated to their outputs (input_output_aliases), so the
in-place `gphm_step` runs without copies.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
FFI_LIB_PATH = os.path.join(os.path.dirname(HERE), "libgphm_ffi.so")     # build.py writes it beside libgphm.so
_registered = False


def _register():
    global _registered
    if _registered:
        return
    import jax                                                  # ImportError here is the honest failure mode
    if not os.path.exists(FFI_LIB_PATH):
        raise RuntimeError("libgphm_ffi.so not found - build experimental/ffi_shim.cc (see its header) against jax.ffi.include_dir()")
    lib = ctypes.CDLL(FFI_LIB_PATH)
    jax.ffi.register_ffi_target("gphm_step", jax.ffi.pycapsule(lib.GphmStep), platform="CUDA")
    jax.ffi.register_ffi_target("gphm_logjoint_grad", jax.ffi.pycapsule(lib.GphmLogjointGrad), platform="CUDA")
    _registered = True


def make_step(plan_handle, lr):
    import jax
    import jax.numpy as jnp
    import numpy as np
    _register()

    def step(U, small, mU, vU, msmall, vsmall, count):
        outs = [jax.ShapeDtypeStruct(a.shape, a.dtype) for a in (U, small, mU, vU, msmall, vsmall, count)]
        outs.append(jax.ShapeDtypeStruct((8,), jnp.float64))
        call = jax.ffi.ffi_call("gphm_step", outs, input_output_aliases={i: i for i in range(7)})
        return call(U, small, mU, vU, msmall, vsmall, count, plan=np.int64(plan_handle), lr=np.float64(lr))
    return step


def make_value_and_grad(plan_handle):
    import jax
    import jax.numpy as jnp
    import numpy as np
    _register()

    def value_and_grad(U, small):
        outs = [jax.ShapeDtypeStruct((8,), jnp.float64), jax.ShapeDtypeStruct(U.shape, U.dtype),
                jax.ShapeDtypeStruct(small.shape, small.dtype)]
        terms, gU, gsmall = jax.ffi.ffi_call("gphm_logjoint_grad", outs)(U, small, plan=np.int64(plan_handle))
        return terms[0], (gU, gsmall), terms
    return value_and_grad
