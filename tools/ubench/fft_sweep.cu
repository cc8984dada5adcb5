// Decomposition of one radix-8 shared-memory FFT sweep (the unit the fused row kernels spend 75 % of the step in):
// what does a sweep cost with only its shared-memory traffic, only its FP64 work, both, and with the
// synchronisation variants?  Standalone (no libgphm): includes the production pass functions from fft_core.cuh.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../gaussian-process-slover-for-high-freq-pde_b200/csrc \
//        -o fft_sweep fft_sweep.cu && ./fft_sweep
//
// Every variant runs REPS x (3 middle forward passes + 3 middle inverse passes) on one resident data set per CTA,
// one CTA per SM (or two for the half-size variant), and reports cycles per sweep (clock64 of CTA 0).
// Variants:
//   full_cta      production passes, __syncthreads after every pass
//   full_groups   production passes, octant-group barriers
//   smem_only     the loads / stores of a pass, butterfly replaced by a register shuffle (no FP64)
//   fp64_only     the butterfly on registers, no shared-memory traffic (one load before, one store after)
//   no_twiddle    production pass with constant twiddles (no table loads)
//   half_2cta     L = 4096 on 256 threads, two CTAs per SM (the shape of the planned one-row-per-CTA design)
#include <cstdio>
#include <cuda_runtime.h>
#include "fft_core.cuh"

using namespace gphm;

constexpr int REPS = 64;

template <int MODE>
__device__ __forceinline__ void pass_variant(double2* xs, int L, int logL, int s, const double2* tw, int tid, bool inverse) {
    const int nt = blockDim.x;
    const int lq = inverse ? s : logL - s - 3, q = 1 << lq;
    for (int b = tid; b < (L >> 3); b += nt) {
        const int j = b & (q - 1);
        const int base = ((b >> lq) << (lq + 3)) + j;
        double2 e[8];
        if (MODE != 2) {
#pragma unroll
            for (int m = 0; m < 8; ++m) e[m] = xs[PADI(base + m * q)];
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) e[m] = make_double2(1.0 + m + tid, 0.5 * m);
        }
        if (MODE == 1) {                       // smem only: keep the data dependent on the loads, no FP64
#pragma unroll
            for (int m = 0; m < 4; ++m) { const double2 t = e[m]; e[m] = e[7 - m]; e[7 - m] = t; }
        } else {
            double2 w1, w2, w4;
            if (MODE == 3 || MODE == 2) { w1 = make_double2(0.6, 0.8); w2 = make_double2(-0.28, 0.96); w4 = make_double2(0.8, -0.6); }
            else { w1 = tw[j]; w2 = tw[q + j]; w4 = tw[2 * q + j]; }
            if (inverse) bfly8_dit_inv(e, w4, w2, w1); else bfly8_dif(e, w1, w2, w4);
        }
        if (MODE != 2) {
#pragma unroll
            for (int m = 0; m < 8; ++m) xs[PADI(base + m * q)] = e[m];
        } else if (e[0].x == 123.456) xs[0] = e[3];          // keep the butterfly alive
    }
}

// MODE 0 production, 1 smem only, 2 fp64 only, 3 no twiddle loads;  GROUPS: octant-group barriers
template <int MODE, bool GROUPS>
__global__ void __launch_bounds__(FFT_THREADS, 1) sweep_kernel(int L, int logL, const double2* __restrict__ W, long long* cyc, double* sink) {
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < fft_data_slots(L); i += nt) xs[i] = make_double2(1e-3 * (i % 97), -1e-3 * (i % 89));
    double2* tw = fft_twiddles(xs, L);
    for (int i = tid; i < fft_twiddle_slots(L); i += nt) tw[i] = make_double2(0.6, 0.8);
    __syncthreads();
    const int np8 = (logL - (logL % 3 == 0 ? 3 : logL % 3)) / 3;
    const long long t0 = clock64();
    for (int r = 0; r < REPS; ++r) {
        const double2* t = tw + 3 * (L >> 3);
        for (int p = 1; p < np8; ++p) {
            pass_variant<MODE>(xs, L, logL, 3 * p, t, tid, false);
            fft_pass_sync<GROUPS>(tid, logL - 6);
            t += 3 * (L >> (3 * p + 3));
        }
        for (int p = np8 - 1; p >= 1; --p) {
            t -= 3 * (L >> (3 * p + 3));
            pass_variant<MODE>(xs, L, logL, logL - 3 - 3 * p, t, tid, true);
            fft_pass_sync<GROUPS>(tid, logL - 6);
        }
        if (GROUPS) __syncthreads();
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) cyc[0] = (t1 - t0) / (REPS * 2 * (np8 - 1));
    if (xs[PADI(tid)].x == 123.456) sink[0] = xs[PADI(tid)].y;
}

template <int MODE, bool GROUPS>
static void run(const char* name, int L, int threads, int ctas_per_sm, const double2* W, long long* cyc, double* sink) {
    int logL = 0;
    while ((1 << logL) < L) ++logL;
    const size_t smem = fft_smem_bytes(L);
    cudaFuncSetAttribute(sweep_kernel<MODE, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    sweep_kernel<MODE, GROUPS><<<sms * ctas_per_sm, threads, smem>>>(L, logL, W, cyc, sink);
    long long c = 0;
    cudaError_t e = cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-12s L=%5d threads=%3d CTAs/SM=%d  smem=%6zu B  cycles/sweep %6lld  (%.2f cycles per point and CTA)  %s\n", name, L,
           threads, ctas_per_sm, smem, c, (double)c / L, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* cyc; double* sink; double2* W;
    cudaMalloc(&cyc, 64); cudaMalloc(&sink, 64); cudaMalloc(&W, 16);
    run<0, false>("full_cta", 8192, 512, 1, W, cyc, sink);
    run<0, true>("full_groups", 8192, 512, 1, W, cyc, sink);
    run<1, false>("smem_only", 8192, 512, 1, W, cyc, sink);
    run<1, true>("smem_groups", 8192, 512, 1, W, cyc, sink);
    run<2, false>("fp64_only", 8192, 512, 1, W, cyc, sink);
    run<3, false>("no_twiddle", 8192, 512, 1, W, cyc, sink);
    run<3, true>("no_tw_groups", 8192, 512, 1, W, cyc, sink);
    run<0, false>("half_1cta", 4096, 256, 1, W, cyc, sink);
    run<0, false>("half_2cta", 4096, 256, 2, W, cyc, sink);
    run<0, true>("half_2cta_gr", 4096, 256, 2, W, cyc, sink);
    run<0, false>("half_512thr", 4096, 512, 1, W, cyc, sink);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
