// In-place shared-memory FFT building blocks shared by fft.cu and toeplitz_inv.cu (FP64 complex,
// power-of-two lengths up to FFT_MAX_L, register-blocked radix-8 passes on a padded layout).
#pragma once
#include <cuda_runtime.h>

namespace gphm {

constexpr int FFT_THREADS = 512;
constexpr int FFT_MAX_L = 8192;                       // 144 KB of (padded) complex doubles in shared memory
constexpr int FFT_ACC = FFT_MAX_L / FFT_THREADS;      // spectrum bins per thread

// shared-memory layout: one pad slot per 8 complex values, so that the 8 lanes of a quarter-warp
// always hit 8 different 16-byte bank groups for every stride the passes use
__device__ __forceinline__ int PADI(int i) { return i + (i >> 3); }
inline size_t fft_smem_bytes(int L) { return (size_t)(L + (L >> 3) + 1) * sizeof(double2); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// K radix-2 DIF stages s..s+K-1 fused in registers: each thread owns the 2^K points
// base + m*q (q = L >> (s+K)) of one sub-transform.
template <int K>
__device__ __forceinline__ void dif_pass(double2* xs, int L, int logL, int s, const double2* __restrict__ W, int tid) {
    constexpr int R = 1 << K;
    const int lq = logL - s - K, q = 1 << lq;
    for (int b = tid; b < (L >> K); b += FFT_THREADS) {
        const int j = b & (q - 1);
        const int base = ((b >> lq) << (lq + K)) + j;
        double2 e[R];
#pragma unroll
        for (int m = 0; m < R; ++m) e[m] = xs[PADI(base + m * q)];
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const int span = R >> (t + 1);
            const double2* __restrict__ Ws = W + (L - (L >> (s + t)));
#pragma unroll
            for (int m = 0; m < R; ++m) {
                if (m & span) continue;
                const double2 a = e[m], c = e[m + span];
                e[m] = make_double2(a.x + c.x, a.y + c.y);
                e[m + span] = cmul(make_double2(a.x - c.x, a.y - c.y), Ws[j + (m & (span - 1)) * q]);
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) xs[PADI(base + m * q)] = e[m];
    }
    __syncthreads();
}

// K radix-2 inverse DIT stages s..s+K-1 (half-spans 2^s .. 2^(s+K-1)) fused in registers.
template <int K>
__device__ __forceinline__ void dit_pass(double2* xs, int L, int logL, int s, const double2* __restrict__ W, int tid) {
    constexpr int R = 1 << K;
    const int q = 1 << s;
    for (int b = tid; b < (L >> K); b += FFT_THREADS) {
        const int j = b & (q - 1);
        const int base = ((b >> s) << (s + K)) + j;
        double2 e[R];
#pragma unroll
        for (int m = 0; m < R; ++m) e[m] = xs[PADI(base + m * q)];
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const int span = 1 << t;
            const double2* __restrict__ Ws = W + (L - (2 << (s + t)));      // table of forward stage logL-1-(s+t)
#pragma unroll
            for (int m = 0; m < R; ++m) {
                if (m & span) continue;
                const double2 w = Ws[j + (m & (span - 1)) * q];
                const double2 tt = cmul(e[m + span], make_double2(w.x, -w.y));
                const double2 a = e[m];
                e[m] = make_double2(a.x + tt.x, a.y + tt.y);
                e[m + span] = make_double2(a.x - tt.x, a.y - tt.y);
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) xs[PADI(base + m * q)] = e[m];
    }
    __syncthreads();
}

// natural order in -> bit-reversed order out
__device__ __forceinline__ void fft_dif_inplace(double2* xs, int L, int logL, const double2* __restrict__ W, int tid) {
    int s = 0;
    for (; logL - s >= 3; s += 3) dif_pass<3>(xs, L, logL, s, W, tid);
    if (logL - s == 2) dif_pass<2>(xs, L, logL, s, W, tid);
    else if (logL - s == 1) dif_pass<1>(xs, L, logL, s, W, tid);
}
// bit-reversed order in -> natural order out (unscaled inverse)
__device__ __forceinline__ void fft_dit_inverse_inplace(double2* xs, int L, int logL, const double2* __restrict__ W, int tid) {
    int s = 0;
    for (; logL - s >= 3; s += 3) dit_pass<3>(xs, L, logL, s, W, tid);
    if (logL - s == 2) dit_pass<2>(xs, L, logL, s, W, tid);
    else if (logL - s == 1) dit_pass<1>(xs, L, logL, s, W, tid);
}

}  // namespace gphm
