// Latency micro-benchmarks (one warp): dependent DFMA chain, MUFU.RCP64H, shuffle, shared-memory round trip, barrier.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int threads_active) {
    __shared__ double sm[64];
    double x = out[0], y = out[1];
    long long t0, t1;
    // 1) dependent DFMA
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); x = fma(x, y, y); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 1024;
    // 2) dependent MUFU.RCP64H
    double r = x;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r)); asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r));
        asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r)); asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r));
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 1024;
    // 3) dependent shuffle
    double s = r;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { s = __shfl_up_sync(0xffffffffu, s, 1); s = __shfl_up_sync(0xffffffffu, s, 1); s = __shfl_up_sync(0xffffffffu, s, 1); s = __shfl_up_sync(0xffffffffu, s, 1); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 1024;
    // 4) shared-memory store -> load dependent
    double v = s;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) { sm[threadIdx.x & 63] = v; __syncwarp(); v = sm[(threadIdx.x + 1) & 63] + 1.0; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 1024;
    // 5) __syncthreads
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 1024;
    // 6) store -> barrier -> load (the kappa broadcast)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) { if (threadIdx.x == (i & 31)) sm[0] = v; __syncthreads(); v = sm[0] + 1.0; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0) / 1024;
    // 7) dependent DADD and DMUL
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) { x = x + y; x = x * y; x = x + y; x = x * y; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0) / 1024;
    // 8) independent DFMA throughput, one warp: 8 chains
    double a0 = x, a1 = y, a2 = v, a3 = s, a4 = r, a5 = x + 1, a6 = y + 1, a7 = v + 1;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        a0 = fma(a0, y, y); a1 = fma(a1, y, y); a2 = fma(a2, y, y); a3 = fma(a3, y, y);
        a4 = fma(a4, y, y); a5 = fma(a5, y, y); a6 = fma(a6, y, y); a7 = fma(a7, y, y);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0) * 100 / 2048;      // x100 cycles per independent DFMA (warp instruction)
    out[2 + threadIdx.x] = x + r + s + v + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 64);
    double h[2] = {0.5, 0.999};
    cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice);
    for (int threads : {32, 256, 512}) {
        k<<<1, threads>>>(out, cyc, threads);
        long long c[8];
        cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads %4d: DFMA dep %lld | MUFU.RCP64H dep %lld | SHFL dep %lld | STS->LDS %lld | __syncthreads %lld | STS->BAR->LDS %lld | DADD/DMUL dep %lld | indep DFMA x100 %lld  (cycles)\n",
               threads, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
