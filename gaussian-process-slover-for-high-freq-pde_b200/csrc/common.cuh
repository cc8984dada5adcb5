// Shared helpers for libgphm (sm_100a).  Internal header - the public C-ABI is include/gphm.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include "../../include/gphm.h"   // status codes GPHM_OK / GPHM_EINVAL / ...

namespace gphm {

void set_last_error(const char* fmt, ...);

#define GPHM_CUDA_OK(expr)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::gphm::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                                   cudaGetErrorString(_e));                                  \
            return GPHM_ECUDA;                                                       \
        }                                                                                    \
    } while (0)

#define GPHM_LAUNCH_OK()  GPHM_CUDA_OK(cudaGetLastError())

#define GPHM_TRY(expr)                                                                       \
    do {                                                                                     \
        int _s = (expr);                                                                     \
        if (_s != 0) return _s;                                                              \
    } while (0)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// cudaFuncSetAttribute is per DEVICE: one-time kernel set-up keyed by the current device, thread-safe.
//   static DeviceOnce once;  if (once.needed()) { ...set attributes...; once.done(); }   (needed() holds the lock until done())
struct DeviceOnce {
    static constexpr int kMaxDevices = 64;
    std::mutex mu;
    bool ready[kMaxDevices] = {};
    int dev = 0;
    bool needed() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
        mu.lock();
        if (ready[d]) { mu.unlock(); return false; }
        dev = d;
        return true;
    }
    void done(bool ok = true) { ready[dev] = ok; mu.unlock(); }
};
#define GPHM_ONCE_CUDA_OK(once, expr)                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            (once).done(false);                                                              \
            ::gphm::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                                   cudaGetErrorString(_e));                                  \
            return GPHM_ECUDA;                                                               \
        }                                                                                    \
    } while (0)

// Launch accounting / optional CUDA-event profiling per kernel family (gphm_profile_* in gphm.h).
enum : int { CAT_GRAM = 0, CAT_DGEMM = 1, CAT_CHOL_DIAG = 2, CAT_ELEMWISE = 3, CAT_ADAM = 4, CAT_FFT = 5, CAT_GS_APPLY = 6,
              CAT_TOEPLITZ_APPLY = 7, CAT_COUNT = 8 };
inline double fft_flops(int L) { double l = 0; for (int t = 1; t < L; t <<= 1) l += 1.0; return 5.0 * L * l; }   // per complex transform
bool profiling_enabled();
struct LaunchScope {
    int cat; cudaStream_t st; int slot;
    LaunchScope(int cat, cudaStream_t st, double flops = 0.0, double bytes = 0.0);
    ~LaunchScope();
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum (fixed tree): every thread gets the result.  `red` holds >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect `red` from a previous call
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = (lane < nw) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

}  // namespace gphm
