// XLA-FFI (jax.ffi custom_call) handlers over the C-ABI of include/gphm.h - the layer BASELINE's
// north_star names for a JAX host.  NOT built in this image: jax / jaxlib and xla/ffi/api/ffi.h are
// absent (build.py compiles this file only when `jax.ffi.include_dir()` resolves), so this source is
// UNTESTED here; the graded boundary is the C-ABI + ctypes path.  Every handler is a plain argument
// adapter: device buffers in, the same libgphm entry point the ctypes host calls, XLA's stream.
//
//   gphm_logjoint_grad  <- jax.value_and_grad(self.loss)(params, key)     model_GP_solver_2d.py:145-174,179
//   gphm_step           <- GP_solver_2d_single.step                        model_GP_solver_2d.py:176-183
//
// Build:  g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())")
//             -I../../include -I/usr/local/cuda/include ffi_shim.cc -L.. -lgphm -o ../libgphm_ffi.so
// The plan handle (gphm_plan_create, eager, in the solver's __init__) travels as an int64 attribute.
#include <cstdint>
#include <cstring>
#include <string>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "gphm.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status_to_error(int rc, const char* what) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + gphm_last_error());
}

// XLA gives every result its own buffer unless the caller declared input_output_aliases; the in-place
// C entry points then need the inputs copied over first.
ffi::Error carry(cudaStream_t stream, const void* src, void* dst, size_t bytes) {
    if (src == dst || bytes == 0) return ffi::Error::Success();
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
        return ffi::Error(ffi::ErrorCode::kInternal, "gphm ffi: device copy failed");
    return ffi::Error::Success();
}

#define GPHM_FFI_TRY(expr)                      \
    do {                                        \
        ffi::Error e_ = (expr);                 \
        if (e_.failure()) return e_;            \
    } while (0)

// (terms[8], gU, gsmall) = value_and_grad at (U, small)
ffi::Error LogjointGradImpl(cudaStream_t stream, int64_t plan, ffi::Buffer<ffi::F64> U, ffi::Buffer<ffi::F64> small,
                            ffi::ResultBuffer<ffi::F64> terms, ffi::ResultBuffer<ffi::F64> gU,
                            ffi::ResultBuffer<ffi::F64> gsmall) {
    if (terms->element_count() != 8 || gU->element_count() != U.element_count() ||
        gsmall->element_count() != small.element_count())
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "gphm_logjoint_grad: result shapes");
    return status_to_error(gphm_logjoint_grad(reinterpret_cast<gphm_plan*>(plan), U.typed_data(), small.typed_data(),
                                              gU->typed_data(), gsmall->typed_data(), terms->typed_data(), 0, stream),
                           "gphm_logjoint_grad");
}

// (U', small', mU', vU', msmall', vsmall', count', terms[8]) = step(U, small, mU, vU, msmall, vsmall, count)
ffi::Error StepImpl(cudaStream_t stream, int64_t plan, double lr, ffi::Buffer<ffi::F64> U, ffi::Buffer<ffi::F64> small,
                    ffi::Buffer<ffi::F64> mU, ffi::Buffer<ffi::F64> vU, ffi::Buffer<ffi::F64> msmall,
                    ffi::Buffer<ffi::F64> vsmall, ffi::Buffer<ffi::S64> count, ffi::ResultBuffer<ffi::F64> U_o,
                    ffi::ResultBuffer<ffi::F64> small_o, ffi::ResultBuffer<ffi::F64> mU_o, ffi::ResultBuffer<ffi::F64> vU_o,
                    ffi::ResultBuffer<ffi::F64> msmall_o, ffi::ResultBuffer<ffi::F64> vsmall_o,
                    ffi::ResultBuffer<ffi::S64> count_o, ffi::ResultBuffer<ffi::F64> terms) {
    if (terms->element_count() != 8 || count.element_count() != 1)
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "gphm_step: terms must hold 8 doubles, count one int64");
    const size_t nf = U.element_count() * sizeof(double), ns = small.element_count() * sizeof(double);
    GPHM_FFI_TRY(carry(stream, U.typed_data(), U_o->typed_data(), nf));
    GPHM_FFI_TRY(carry(stream, mU.typed_data(), mU_o->typed_data(), nf));
    GPHM_FFI_TRY(carry(stream, vU.typed_data(), vU_o->typed_data(), nf));
    GPHM_FFI_TRY(carry(stream, small.typed_data(), small_o->typed_data(), ns));
    GPHM_FFI_TRY(carry(stream, msmall.typed_data(), msmall_o->typed_data(), ns));
    GPHM_FFI_TRY(carry(stream, vsmall.typed_data(), vsmall_o->typed_data(), ns));
    GPHM_FFI_TRY(carry(stream, count.typed_data(), count_o->typed_data(), sizeof(int64_t)));
    static_assert(sizeof(long long) == sizeof(int64_t), "count is a device int64");
    return status_to_error(gphm_step(reinterpret_cast<gphm_plan*>(plan), U_o->typed_data(), small_o->typed_data(),
                                     mU_o->typed_data(), vU_o->typed_data(), msmall_o->typed_data(), vsmall_o->typed_data(),
                                     reinterpret_cast<long long*>(count_o->typed_data()), lr, terms->typed_data(), stream),
                           "gphm_step");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(GphmLogjointGrad, LogjointGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F64>>()    // U
                                  .Arg<ffi::Buffer<ffi::F64>>()    // small
                                  .Ret<ffi::Buffer<ffi::F64>>()    // terms
                                  .Ret<ffi::Buffer<ffi::F64>>()    // gU
                                  .Ret<ffi::Buffer<ffi::F64>>());  // gsmall

XLA_FFI_DEFINE_HANDLER_SYMBOL(GphmStep, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<double>("lr")
                                  .Arg<ffi::Buffer<ffi::F64>>()    // U
                                  .Arg<ffi::Buffer<ffi::F64>>()    // small
                                  .Arg<ffi::Buffer<ffi::F64>>()    // mU
                                  .Arg<ffi::Buffer<ffi::F64>>()    // vU
                                  .Arg<ffi::Buffer<ffi::F64>>()    // msmall
                                  .Arg<ffi::Buffer<ffi::F64>>()    // vsmall
                                  .Arg<ffi::Buffer<ffi::S64>>()    // count
                                  .Ret<ffi::Buffer<ffi::F64>>()    // U'
                                  .Ret<ffi::Buffer<ffi::F64>>()    // small'
                                  .Ret<ffi::Buffer<ffi::F64>>()    // mU'
                                  .Ret<ffi::Buffer<ffi::F64>>()    // vU'
                                  .Ret<ffi::Buffer<ffi::F64>>()    // msmall'
                                  .Ret<ffi::Buffer<ffi::F64>>()    // vsmall'
                                  .Ret<ffi::Buffer<ffi::S64>>()    // count'
                                  .Ret<ffi::Buffer<ffi::F64>>());  // terms
