// In-place shared-memory FFT building blocks shared by fft.cu and toeplitz_inv.cu (FP64 complex,
// power-of-two lengths up to FFT_MAX_L, register-blocked radix-8 passes on a padded layout).
//
// Twiddles: a radix-8 pass that fuses the radix-2 stages s, s+1, s+2 (butterfly j of q = L >> (s+3))
// needs  exp(-2 pi i (j + m q) 2^(s+t) / L)  = w_t[j] * W8^m / W4^m / 1  with only three table values
//   w1 = T_s[j],  w2 = T_{s+1}[j],  w4 = T_{s+2}[j]          (T_s[j] = exp(-2 pi i j 2^s / L))
// and the constants W8^m.  The 3*q values per pass (56 KB for L = 8192, all passes) are staged once
// per CTA in shared memory behind the data (the full per-stage tables - 128 KB, seven gathers per
// butterfly through an L1 that the 144 KB data buffer leaves ~100 KB of - were the bottleneck:
// long-scoreboard stalls).  Forward = DIF (natural in, bit-reversed out), passes (8,8,...,tail);
// inverse = DIT (bit-reversed in, natural out), passes (head,8,...,8); the inverse pass at stage s
// uses the table of the forward pass at stage logL-3-s.  The tail/head pass has q = 1: constants only.
#pragma once
#include <algorithm>
#include <cuda_runtime.h>

namespace gphm {

constexpr int FFT_THREADS = 512;
constexpr int FFT_MAX_L = 8192;                       // 144 KB of (padded) complex doubles in shared memory
constexpr int FFT_ACC = FFT_MAX_L / FFT_THREADS;      // spectrum bins per thread
// Threads per CTA are a LAUNCH parameter (blockDim.x <= FFT_THREADS): the fused row kernels run short transforms
// with L/8 threads (one radix-8 butterfly per thread and pass; several CTAs per SM) instead of idling 3/4 of a
// 512-thread CTA at L = 1024.  Every loop below strides by fft_nt<NT>(); the per-thread register arrays (FFT_ACC
// slots) need blockDim.x >= L / 16.
// NT > 0: compile-time stride (the 512-thread launches of long transforms); NT = 0: blockDim.x.
template <int NT> __device__ __forceinline__ int fft_nt() { return NT > 0 ? NT : (int)blockDim.x; }
// Octant groups (L >= 512, L/8 threads; 512 threads at L = 8192).  After the first forward pass the eight
// octants of the data are independent sub-transforms until the last inverse pass, and the butterfly -> thread map
// of every pass in between (b = tid + k * threads, octant = b / (L/64)) keeps an octant inside one group of
// G = L/64 consecutive threads (L = 8192: octants g and g+4 in group g of 128 threads, one warp per scheduler).
// With GR those passes synchronise per group - a named barrier (ids 1..8) for G >= 64, __syncwarp below -
// instead of per CTA: 2 CTA-wide barriers per convolution instead of 9.  lG = log2(G) = logL - 6.
template <bool GR> __device__ __forceinline__ void fft_pass_sync(int tid, int lG) {
    if (!GR) __syncthreads();
    else if (lG >= 6) asm volatile("bar.sync %0, %1;" :: "r"(1 + (tid >> lG)), "r"(1 << lG) : "memory");
    else __syncwarp();
}
__host__ inline bool fft_groups_supported(int L) { return L >= 512; }
__host__ inline int fft_threads_for(int L) { return std::max(64, std::min(FFT_THREADS, L >> 3)); }

// shared-memory layout: one pad slot per 8 complex values, so that the 8 lanes of a quarter-warp
// always hit 8 different 16-byte bank groups for every stride the passes use
__device__ __forceinline__ int PADI(int i) { return i + (i >> 3); }
__host__ __device__ inline int fft_data_slots(int L) { return L + (L >> 3) + 1; }
// Twiddles of a radix-8 pass: w1 = T_s[j], w2 = T_{s+1}[j], w4 = T_{s+2}[j] from the shared-memory table.  -DGPHM_TW_SQUARE
// computes w2 = w1^2, w4 = w2^2 instead (6 FP64 operations for two 16-byte shared-memory loads, table 56 -> 19 KB at
// L = 8192): measured in round 2 on the 4096^2 step and REJECTED - gs_apply_fused 3.87 -> 3.98 ms per step (the FP64 pipe and
// the register allocation of the 128-register kernels matter as much as the shared-memory traffic), toeplitz_apply_fused
// 1.50 -> 1.46 ms.
#ifdef GPHM_TW_SQUARE
constexpr int kTwPerPass = 1;
#else
constexpr int kTwPerPass = 3;
#endif
// entries of the compact twiddle table: kTwPerPass * sum over radix-8 passes of q_p = L >> (3p+3)
__host__ __device__ inline int fft_twiddle_slots(int L) {
    int logL = 0;
    while ((1 << logL) < L) ++logL;
    int n = 0;
    for (int s = 0; logL - s >= 3; s += 3) n += kTwPerPass * (L >> (s + 3));
    return n;
}
inline size_t fft_smem_bytes(int L) { return (size_t)(fft_data_slots(L) + fft_twiddle_slots(L)) * sizeof(double2); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {      // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// Shared-memory twiddle table behind the data buffer; W is the global per-stage table built by
// twiddle_init_kernel:  W[(L - (L >> s)) + j] = T_s[j],  j < L >> (s+1).   Ends with a barrier.
__device__ __forceinline__ double2* fft_twiddles(double2* xs, int L) { return xs + fft_data_slots(L); }
template <int NT = FFT_THREADS>
__device__ __forceinline__ void fft_load_twiddles(double2* xs, int L, int logL, const double2* __restrict__ W, int tid) {
    double2* tw = fft_twiddles(xs, L);
    for (int s = 0; logL - s >= 3; s += 3) {
        const int q = L >> (s + 3);
        for (int i = tid; i < kTwPerPass * q; i += fft_nt<NT>()) {
            const int t = i / q, j = i - t * q;
            tw[i] = W[(L - (L >> (s + t))) + j];
        }
        tw += kTwPerPass * q;
    }
    __syncthreads();
}

constexpr double kRsqrt2 = 0.70710678118654752440;

__device__ __forceinline__ double2 csqr(double2 a) { return make_double2(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y); }
// (w1, w2, w4) of butterfly j of a pass with q butterflies per group; tw = the pass' table
__device__ __forceinline__ void fft_tw3(const double2* tw, int q, int j, double2& w1, double2& w2, double2& w4) {
    w1 = tw[j];
    if (kTwPerPass == 3) { w2 = tw[q + j]; w4 = tw[2 * q + j]; }
    else { w2 = csqr(w1); w4 = csqr(w2); }
}

// One radix-8 DIF butterfly: e[0..7] are the points base + m*q; output slot m holds frequency
// brev3(m) of the 8-point sub-transform, twiddled for the next pass.
__device__ __forceinline__ void bfly8_dif(double2 (&e)[8], double2 w1, double2 w2, double2 w4) {
    // stage 0: span 4, twiddle w1 * W8^m
    {
        const double2 d0 = csub(e[0], e[4]), d1 = csub(e[1], e[5]), d2 = csub(e[2], e[6]), d3 = csub(e[3], e[7]);
        e[0] = cadd(e[0], e[4]); e[1] = cadd(e[1], e[5]); e[2] = cadd(e[2], e[6]); e[3] = cadd(e[3], e[7]);
        e[4] = cmul(d0, w1);
        e[5] = cmul(make_double2((d1.x + d1.y) * kRsqrt2, (d1.y - d1.x) * kRsqrt2), w1);
        e[6] = cmul(make_double2(d2.y, -d2.x), w1);
        e[7] = cmul(make_double2((d3.y - d3.x) * kRsqrt2, -(d3.x + d3.y) * kRsqrt2), w1);
    }
    // stage 1: span 2 inside each half, twiddle w2 * W4^m
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        const double2 d0 = csub(e[h], e[h + 2]), d1 = csub(e[h + 1], e[h + 3]);
        e[h] = cadd(e[h], e[h + 2]); e[h + 1] = cadd(e[h + 1], e[h + 3]);
        e[h + 2] = cmul(d0, w2);
        e[h + 3] = cmul(make_double2(d1.y, -d1.x), w2);
    }
    // stage 2: span 1, twiddle w4
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
        const double2 d = csub(e[h], e[h + 1]);
        e[h] = cadd(e[h], e[h + 1]);
        e[h + 1] = cmul(d, w4);
    }
}

// One radix-8 inverse DIT butterfly (conjugate twiddles): stages span 1 (u1), span 2 (u2 * W4^m), span 4 (u4 * W8^m).
__device__ __forceinline__ void bfly8_dit_inv(double2 (&e)[8], double2 u1, double2 u2, double2 u4) {
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
        const double2 t = cmulc(e[h + 1], u1);
        const double2 a = e[h];
        e[h] = cadd(a, t); e[h + 1] = csub(a, t);
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        const double2 t0 = cmulc(e[h + 2], u2);
        const double2 v = cmulc(e[h + 3], u2);
        const double2 t1 = make_double2(-v.y, v.x);                 // * conj(W4) = * (+i)
        const double2 a0 = e[h], a1 = e[h + 1];
        e[h] = cadd(a0, t0); e[h + 2] = csub(a0, t0);
        e[h + 1] = cadd(a1, t1); e[h + 3] = csub(a1, t1);
    }
    {
        const double2 t0 = cmulc(e[4], u4);
        const double2 v1 = cmulc(e[5], u4), v2 = cmulc(e[6], u4), v3 = cmulc(e[7], u4);
        const double2 t1 = make_double2((v1.x - v1.y) * kRsqrt2, (v1.x + v1.y) * kRsqrt2);      // * (1+i)/sqrt2
        const double2 t2 = make_double2(-v2.y, v2.x);                                           // * i
        const double2 t3 = make_double2(-(v3.x + v3.y) * kRsqrt2, (v3.x - v3.y) * kRsqrt2);     // * (-1+i)/sqrt2
        const double2 a0 = e[0], a1 = e[1], a2 = e[2], a3 = e[3];
        e[0] = cadd(a0, t0); e[4] = csub(a0, t0);
        e[1] = cadd(a1, t1); e[5] = csub(a1, t1);
        e[2] = cadd(a2, t2); e[6] = csub(a2, t2);
        e[3] = cadd(a3, t3); e[7] = csub(a3, t3);
    }
}

// Radix-8 DIF pass over stages s..s+2; tw = this pass' compact table [w1 | w2 | w4], q entries each.
template <int NT = FFT_THREADS, bool GR = false>
__device__ __forceinline__ void dif_pass8(double2* xs, int L, int logL, int s, const double2* tw, int tid) {
    const int lq = logL - s - 3, q = 1 << lq;
    for (int b = tid; b < (L >> 3); b += fft_nt<NT>()) {
        const int j = b & (q - 1);
        const int base = ((b >> lq) << (lq + 3)) + j;
        double2 e[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) e[m] = xs[PADI(base + m * q)];
        double2 w1, w2, w4;
        fft_tw3(tw, q, j, w1, w2, w4);
        bfly8_dif(e, w1, w2, w4);
#pragma unroll
        for (int m = 0; m < 8; ++m) xs[PADI(base + m * q)] = e[m];
    }
    fft_pass_sync<GR>(tid, logL - 6);
}

template <int NT = FFT_THREADS, bool GR = false>
__device__ __forceinline__ void dit_pass8(double2* xs, int L, int logL, int s, const double2* tw, int tid) {
    const int q = 1 << s;
    for (int b = tid; b < (L >> 3); b += fft_nt<NT>()) {
        const int j = b & (q - 1);
        const int base = ((b >> s) << (s + 3)) + j;
        double2 e[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) e[m] = xs[PADI(base + m * q)];
        double2 w1, w2, w4;
        fft_tw3(tw, q, j, w1, w2, w4);
        bfly8_dit_inv(e, w4, w2, w1);                                // table of the forward pass at stage logL-3-s
#pragma unroll
        for (int m = 0; m < 8; ++m) xs[PADI(base + m * q)] = e[m];
    }
    fft_pass_sync<GR>(tid, logL - 6);
}

// Tail (forward) / head (inverse) pass of 2^K points with unit stride: every twiddle is a constant.
template <int K, int NT = FFT_THREADS>
__device__ __forceinline__ void unit_pass(double2* xs, int L, bool inverse, int tid) {
    constexpr int R = 1 << K;
    for (int b = tid; b < (L >> K); b += fft_nt<NT>()) {
        const int base = b << K;
        double2 e[R];
#pragma unroll
        for (int m = 0; m < R; ++m) e[m] = xs[PADI(base + m)];
        if (K == 1) {
            const double2 a = e[0], c = e[1];
            e[0] = cadd(a, c); e[1] = csub(a, c);
        } else if (!inverse) {   // K == 2, DIF: span 2 (twiddle W4^m), then span 1
            const double2 d0 = csub(e[0], e[2]), d1 = csub(e[1], e[3]);
            const double2 s0 = cadd(e[0], e[2]), s1 = cadd(e[1], e[3]);
            const double2 r1 = make_double2(d1.y, -d1.x);
            e[0] = cadd(s0, s1); e[1] = csub(s0, s1); e[2] = cadd(d0, r1); e[3] = csub(d0, r1);
        } else {                 // K == 2, inverse DIT: span 1, then span 2 (twiddle conj W4^m)
            const double2 a0 = cadd(e[0], e[1]), a1 = csub(e[0], e[1]);
            const double2 b0 = cadd(e[2], e[3]), v = csub(e[2], e[3]);
            const double2 b1 = make_double2(-v.y, v.x);
            e[0] = cadd(a0, b0); e[2] = csub(a0, b0); e[1] = cadd(a1, b1); e[3] = csub(a1, b1);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) xs[PADI(base + m)] = e[m];
    }
    __syncthreads();
}

// natural order in -> bit-reversed order out.  Needs fft_load_twiddles(xs, L, ...) once per CTA.
template <int NT = FFT_THREADS>
__device__ __forceinline__ void fft_dif_inplace(double2* xs, int L, int logL, const double2* __restrict__, int tid) {
    const double2* tw = fft_twiddles(xs, L);
    int s = 0;
    for (; logL - s >= 3; s += 3) { dif_pass8<NT>(xs, L, logL, s, tw, tid); tw += kTwPerPass * (L >> (s + 3)); }
    if (logL - s == 2) unit_pass<2, NT>(xs, L, false, tid);
    else if (logL - s == 1) unit_pass<1, NT>(xs, L, false, tid);
}
// bit-reversed order in -> natural order out (unscaled inverse)
template <int NT = FFT_THREADS>
__device__ __forceinline__ void fft_dit_inverse_inplace(double2* xs, int L, int logL, const double2* __restrict__, int tid) {
    const int r = logL % 3;
    if (r == 2) unit_pass<2, NT>(xs, L, true, tid);
    else if (r == 1) unit_pass<1, NT>(xs, L, true, tid);
    // inverse pass at stage s pairs with the forward pass at stage logL-3-s: walk the table backwards
    const double2* tw = fft_twiddles(xs, L) + fft_twiddle_slots(L);
    for (int s = r; s + 3 <= logL; s += 3) {
        tw -= kTwPerPass * (1 << s);
        dit_pass8<NT>(xs, L, logL, s, tw, tid);
    }
}


// ---------------------------------------------------------------------------------------------
// Fused pipeline pieces (toeplitz_fused.cu): the same transforms with fewer shared-memory sweeps.
//   pass structure  forward: np8 radix-8 passes (stages 0 .. 3 np8 - 1), then a unit-stride tail
//   of KT = logL - 3 np8 in {1,2,3} stages whose twiddles are constants; inverse: head, np8 passes.
//   dif_first          global load (zero padded) fused into the first forward pass
//   mid_fused          forward tail + pointwise functor + inverse head in registers
//   dit_last           last inverse pass fused into the global store (only indices < L/2 are live)
//   dit_last_dif_first last inverse pass + truncation to n + first forward pass of the next
//                      convolution in registers (both touch the points j + m L/8)
// A convolution costs 2 np8 + 1 sweeps instead of 2 (np8 + 1) + 3.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int fft_tail_stages(int logL) { return logL % 3 == 0 ? 3 : logL % 3; }
// kernels without a per-bin register stash (toeplitz_apply_fused) merge the last radix-8 pass into a 16-point tail
__host__ __device__ inline int fft_tail_stages_wide(int logL) { return (logL % 3 == 1 && logL >= 7) ? 4 : fft_tail_stages(logL); }

__device__ __forceinline__ void bfly8_dif_unit(double2 (&e)[8]) {      // bfly8_dif with w1 = w2 = w4 = 1
    {
        const double2 d0 = csub(e[0], e[4]), d1 = csub(e[1], e[5]), d2 = csub(e[2], e[6]), d3 = csub(e[3], e[7]);
        e[0] = cadd(e[0], e[4]); e[1] = cadd(e[1], e[5]); e[2] = cadd(e[2], e[6]); e[3] = cadd(e[3], e[7]);
        e[4] = d0;
        e[5] = make_double2((d1.x + d1.y) * kRsqrt2, (d1.y - d1.x) * kRsqrt2);
        e[6] = make_double2(d2.y, -d2.x);
        e[7] = make_double2((d3.y - d3.x) * kRsqrt2, -(d3.x + d3.y) * kRsqrt2);
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        const double2 d0 = csub(e[h], e[h + 2]), d1 = csub(e[h + 1], e[h + 3]);
        e[h] = cadd(e[h], e[h + 2]); e[h + 1] = cadd(e[h + 1], e[h + 3]);
        e[h + 2] = d0;
        e[h + 3] = make_double2(d1.y, -d1.x);
    }
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
        const double2 d = csub(e[h], e[h + 1]);
        e[h] = cadd(e[h], e[h + 1]);
        e[h + 1] = d;
    }
}

__device__ __forceinline__ void bfly8_dit_inv_unit(double2 (&e)[8]) {
#pragma unroll
    for (int h = 0; h < 8; h += 2) {
        const double2 a = e[h], t = e[h + 1];
        e[h] = cadd(a, t); e[h + 1] = csub(a, t);
    }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        const double2 t0 = e[h + 2], v = e[h + 3];
        const double2 t1 = make_double2(-v.y, v.x);
        const double2 a0 = e[h], a1 = e[h + 1];
        e[h] = cadd(a0, t0); e[h + 2] = csub(a0, t0);
        e[h + 1] = cadd(a1, t1); e[h + 3] = csub(a1, t1);
    }
    {
        const double2 t0 = e[4], v1 = e[5], v2 = e[6], v3 = e[7];
        const double2 t1 = make_double2((v1.x - v1.y) * kRsqrt2, (v1.x + v1.y) * kRsqrt2);
        const double2 t2 = make_double2(-v2.y, v2.x);
        const double2 t3 = make_double2(-(v3.x + v3.y) * kRsqrt2, (v3.x - v3.y) * kRsqrt2);
        const double2 a0 = e[0], a1 = e[1], a2 = e[2], a3 = e[3];
        e[0] = cadd(a0, t0); e[4] = csub(a0, t0);
        e[1] = cadd(a1, t1); e[5] = csub(a1, t1);
        e[2] = cadd(a2, t2); e[6] = csub(a2, t2);
        e[3] = cadd(a3, t3); e[7] = csub(a3, t3);
    }
}

// 16 contiguous points, the last four forward stages (spans 8, 4, 2, 1): span 8 with the constants W16^m, then the
// 8-point unit butterfly on both halves.  Merges the last radix-8 pass with the radix-2 tail of lengths with
// log2 L = 1 (mod 3) - one shared-memory sweep less per transform.
constexpr double kC16 = 0.92387953251128675613, kS16 = 0.38268343236508977173;     // cos(pi/8), sin(pi/8)
__device__ __forceinline__ double2 mul_w16(double2 v, int m) {       // v * exp(-2 pi i m / 16), m = 0 .. 7 (compile-time after unrolling)
    switch (m) {
        case 0: return v;
        case 1: return make_double2(kC16 * v.x + kS16 * v.y, kC16 * v.y - kS16 * v.x);
        case 2: return make_double2((v.x + v.y) * kRsqrt2, (v.y - v.x) * kRsqrt2);
        case 3: return make_double2(kS16 * v.x + kC16 * v.y, kS16 * v.y - kC16 * v.x);
        case 4: return make_double2(v.y, -v.x);
        case 5: return make_double2(kC16 * v.y - kS16 * v.x, -(kC16 * v.x + kS16 * v.y));
        case 6: return make_double2((v.y - v.x) * kRsqrt2, -(v.x + v.y) * kRsqrt2);
        default: return make_double2(kS16 * v.y - kC16 * v.x, -(kS16 * v.x + kC16 * v.y));
    }
}
__device__ __forceinline__ double2 mul_w16c(double2 v, int m) {      // v * exp(+2 pi i m / 16)
    const double2 t = mul_w16(make_double2(v.x, -v.y), m);
    return make_double2(t.x, -t.y);
}
__device__ __forceinline__ void bfly16_dif_unit(double2 (&e)[16]) {
    double2 lo[8], hi[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) { lo[m] = cadd(e[m], e[m + 8]); hi[m] = mul_w16(csub(e[m], e[m + 8]), m); }
    bfly8_dif_unit(lo);
    bfly8_dif_unit(hi);
#pragma unroll
    for (int m = 0; m < 8; ++m) { e[m] = lo[m]; e[m + 8] = hi[m]; }
}
__device__ __forceinline__ void bfly16_dit_inv_unit(double2 (&e)[16]) {
    double2 lo[8], hi[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) { lo[m] = e[m]; hi[m] = e[m + 8]; }
    bfly8_dit_inv_unit(lo);
    bfly8_dit_inv_unit(hi);
#pragma unroll
    for (int m = 0; m < 8; ++m) { const double2 t = mul_w16c(hi[m], m); e[m] = cadd(lo[m], t); e[m + 8] = csub(lo[m], t); }
}

template <int KT>
__device__ __forceinline__ void unit_fwd(double2 (&e)[1 << KT]) {
    if constexpr (KT == 4) {
        bfly16_dif_unit(e);
    } else if constexpr (KT == 1) {
        const double2 a = e[0], c = e[1];
        e[0] = cadd(a, c); e[1] = csub(a, c);
    } else if constexpr (KT == 2) {
        const double2 d0 = csub(e[0], e[2]), d1 = csub(e[1], e[3]);
        const double2 s0 = cadd(e[0], e[2]), s1 = cadd(e[1], e[3]);
        const double2 r1 = make_double2(d1.y, -d1.x);
        e[0] = cadd(s0, s1); e[1] = csub(s0, s1); e[2] = cadd(d0, r1); e[3] = csub(d0, r1);
    } else {
        bfly8_dif_unit(e);
    }
}
template <int KT>
__device__ __forceinline__ void unit_inv(double2 (&e)[1 << KT]) {
    if constexpr (KT == 4) {
        bfly16_dit_inv_unit(e);
    } else if constexpr (KT == 1) {
        const double2 a = e[0], c = e[1];
        e[0] = cadd(a, c); e[1] = csub(a, c);
    } else if constexpr (KT == 2) {
        const double2 a0 = cadd(e[0], e[1]), a1 = csub(e[0], e[1]);
        const double2 b0 = cadd(e[2], e[3]), v = csub(e[2], e[3]);
        const double2 b1 = make_double2(-v.y, v.x);
        e[0] = cadd(a0, b0); e[2] = csub(a0, b0); e[1] = cadd(a1, b1); e[3] = csub(a1, b1);
    } else {
        bfly8_dit_inv_unit(e);
    }
}

// First forward pass (stage 0, q = L/8) with the input taken from ld(index); indices >= L/2 are zero padding.
template <int NT = FFT_THREADS, class Load>
__device__ __forceinline__ void dif_first(double2* xs, int L, const double2* tw, int tid, Load ld) {
    const int q = L >> 3;
    for (int j = tid; j < q; j += fft_nt<NT>()) {
        double2 e[8];
#pragma unroll
        for (int m = 0; m < 4; ++m) e[m] = ld(j + m * q);
#pragma unroll
        for (int m = 4; m < 8; ++m) e[m] = make_double2(0.0, 0.0);
        double2 w1, w2, w4;
        fft_tw3(tw, q, j, w1, w2, w4);
        bfly8_dif(e, w1, w2, w4);
#pragma unroll
        for (int m = 0; m < 8; ++m) xs[PADI(j + m * q)] = e[m];
    }
    __syncthreads();
}

// Forward passes 1 .. np8-1 (after dif_first).
template <int NT = FFT_THREADS, bool GR = false>
__device__ __forceinline__ void dif_middle(double2* xs, int L, int logL, int np8, int tid) {
    const double2* tw = fft_twiddles(xs, L) + kTwPerPass * (L >> 3);
    for (int p = 1; p < np8; ++p) { dif_pass8<NT, GR>(xs, L, logL, 3 * p, tw, tid); tw += kTwPerPass * (L >> (3 * p + 3)); }
}
// Inverse passes at stages KT, KT+3, ..., logL-6 (all but the last one).
// With GR the caller needs a CTA-wide barrier before the last inverse pass (it mixes the octants): added here.
template <int NT = FFT_THREADS, bool GR = false>
__device__ __forceinline__ void dit_middle(double2* xs, int L, int logL, int np8, int KT, int tid) {
    const double2* tw = fft_twiddles(xs, L);
    for (int p = 0; p < np8; ++p) tw += kTwPerPass * (L >> (3 * p + 3));
    for (int p = np8 - 1; p >= 1; --p) {          // inverse stage s = logL - 3 - 3p pairs with forward pass p
        tw -= kTwPerPass * (L >> (3 * p + 3));
        dit_pass8<NT, GR>(xs, L, logL, logL - 3 - 3 * p, tw, tid);
    }
    if (GR) __syncthreads();
}

// Bin group (2^KT contiguous bins) of iteration i of a thread in the tail / head pass.  GR: the thread group
// tid >> lG works on its own octant(s): 8 >> KT iterations per octant, G bin groups of the octant per iteration.
template <int KT, int NT, bool GR>
__device__ __forceinline__ int fft_mid_group(int tid, int i, int L, int lG) {
    if (!GR) return tid + i * fft_nt<NT>();
    if constexpr (KT == 4) {
        // 16-point bin groups: an octant holds L/128 of them = half a thread group; a thread group owns 8 / #groups octants
        const int ngroups = fft_nt<NT>() >> lG, opg = 8 / ngroups, lbpo = lG - 1;
        const int l = tid & ((1 << lG) - 1), k = l >> lbpo;
        if (k >= opg) return L;                                   // idle thread (lengths <= 4096: one octant per group)
        return (((tid >> lG) + ngroups * k) << lbpo) + (l & ((1 << lbpo) - 1));
    }
    constexpr int IPO = KT >= 4 ? 1 : (8 >> KT);                  // iterations per octant
    const int octant = (tid >> lG) + (fft_nt<NT>() >> lG) * (i / IPO);
    return octant * ((L >> 3) >> KT) + (tid & ((1 << lG) - 1)) + ((i % IPO) << lG);
}

// Forward tail + functor + inverse head on groups of 2^KT contiguous (bit-reversed-order) bins.
// f(slot, p, v): slot = static register slot (0 .. 15) of this thread, p = bin position, v = spectrum value.
template <int KT, int NT = FFT_THREADS, bool GR = false, class F>
__device__ __forceinline__ void mid_fused(double2* xs, int L, int logL, int tid, F f) {
    constexpr int R = 1 << KT;
    constexpr int MAXG = FFT_MAX_L / (R * FFT_THREADS);
    static_assert(MAXG >= 1, "bin group too wide");
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
        const int g = fft_mid_group<KT, NT, GR>(tid, i, L, logL - 6);
        if (g < (L >> KT)) {
            const int base = g << KT;
            double2 e[R];
#pragma unroll
            for (int m = 0; m < R; ++m) e[m] = xs[PADI(base + m)];
            unit_fwd<KT>(e);
#pragma unroll
            for (int m = 0; m < R; ++m) e[m] = f(i * R + m, base + m, e[m]);
            unit_inv<KT>(e);
#pragma unroll
            for (int m = 0; m < R; ++m) xs[PADI(base + m)] = e[m];
        }
    }
    fft_pass_sync<GR>(tid, logL - 6);
}

// mid_fused with the pointwise operand taken from ONE global array `spec` (L complex values, bin order) and loaded one
// iteration AHEAD of its use: the tail pass does almost no arithmetic, so with the load issued where it is consumed each of
// its 8 iterations exposed a full L2 round trip (ncu: the three radix-2 tail passes of gs_apply_fused cost 21.6 % of the
// kernel, 2.4 radix-8 sweeps each).  f(slot, p, v, s): s = spec[p].
template <int KT, int NT = FFT_THREADS, bool GR = false, bool PF = true, class F>
__device__ __forceinline__ void mid_fused_spec(double2* xs, int L, int logL, int tid, const double2* __restrict__ spec, F f) {
    if constexpr (!PF) {          // load where it is used
        mid_fused<KT, NT, GR>(xs, L, logL, tid, [&](int slot, int p, double2 v) { return f(slot, p, v, spec[p]); });
        return;
    }
    constexpr int R = 1 << KT;
    constexpr int MAXG = FFT_MAX_L / (R * FFT_THREADS);
    static_assert(MAXG >= 1, "bin group too wide");
    double2 sn[R];
    int gn = fft_mid_group<KT, NT, GR>(tid, 0, L, logL - 6);
    if (gn < (L >> KT)) {
#pragma unroll
        for (int m = 0; m < R; ++m) sn[m] = spec[(gn << KT) + m];
    }
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
        const int g = gn;
        double2 sc[R];
#pragma unroll
        for (int m = 0; m < R; ++m) sc[m] = sn[m];
        if (i + 1 < MAXG) {
            gn = fft_mid_group<KT, NT, GR>(tid, i + 1, L, logL - 6);
            if (gn < (L >> KT)) {
#pragma unroll
                for (int m = 0; m < R; ++m) sn[m] = spec[(gn << KT) + m];
            }
        }
        if (g < (L >> KT)) {
            const int base = g << KT;
            double2 e[R];
#pragma unroll
            for (int m = 0; m < R; ++m) e[m] = xs[PADI(base + m)];
            unit_fwd<KT>(e);
#pragma unroll
            for (int m = 0; m < R; ++m) e[m] = f(i * R + m, base + m, e[m], sc[m]);
            unit_inv<KT>(e);
#pragma unroll
            for (int m = 0; m < R; ++m) xs[PADI(base + m)] = e[m];
        }
    }
    fft_pass_sync<GR>(tid, logL - 6);
}

// First forward pass with BOTH butterflies' global loads of a thread issued before the first butterfly (512 threads, L = 8192:
// two butterflies per thread; the kernels that call it have no live register stash at this point).
template <int NT = FFT_THREADS, class Load>
__device__ __forceinline__ void dif_first2(double2* xs, int L, const double2* tw, int tid, Load ld) {
    const int q = L >> 3, nt = fft_nt<NT>();
    for (int j0 = tid; j0 < q; j0 += 2 * nt) {
        const int j1 = j0 + nt;
        double2 a[4], b[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) a[m] = ld(j0 + m * q);
        if (j1 < q) {
#pragma unroll
            for (int m = 0; m < 4; ++m) b[m] = ld(j1 + m * q);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = h == 0 ? j0 : j1;
            if (j < q) {
                double2 e[8];
#pragma unroll
                for (int m = 0; m < 4; ++m) e[m] = h == 0 ? a[m] : b[m];
#pragma unroll
                for (int m = 4; m < 8; ++m) e[m] = make_double2(0.0, 0.0);
                double2 w1, w2, w4;
                fft_tw3(tw, q, j, w1, w2, w4);
                bfly8_dif(e, w1, w2, w4);
#pragma unroll
                for (int m = 0; m < 8; ++m) xs[PADI(j + m * q)] = e[m];
            }
        }
    }
    __syncthreads();
}

// Last inverse pass (stage logL-3, q = L/8); st(index, value, addend) for the live half (index < L/2).
// The addends pre(index) are fetched before the butterfly so that their global-memory latency
// hides behind it (the stores may alias the addend, so the compiler cannot hoist the loads itself).
template <int NT = FFT_THREADS, class Pre, class Store>
__device__ __forceinline__ void dit_last(double2* xs, int L, const double2* tw, int tid, Pre pre, Store st) {
    const int q = L >> 3;
    for (int j = tid; j < q; j += fft_nt<NT>()) {
        double2 add[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) add[m] = pre(j + m * q);
        double2 e[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) e[m] = xs[PADI(j + m * q)];
        double2 w1, w2, w4;
        fft_tw3(tw, q, j, w1, w2, w4);
        bfly8_dit_inv(e, w4, w2, w1);
#pragma unroll
        for (int m = 0; m < 4; ++m) st(j + m * q, e[m], add[m]);
    }
    __syncthreads();
}

// Last inverse pass, truncation to the first n entries, first forward pass of the next convolution.
template <int NT = FFT_THREADS>
__device__ __forceinline__ void dit_last_dif_first(double2* xs, int L, const double2* tw, int tid, int n) {
    const int q = L >> 3;
    for (int j = tid; j < q; j += fft_nt<NT>()) {
        double2 e[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) e[m] = xs[PADI(j + m * q)];
        double2 w1, w2, w4;
        fft_tw3(tw, q, j, w1, w2, w4);
        bfly8_dit_inv(e, w4, w2, w1);
#pragma unroll
        for (int m = 0; m < 4; ++m) if (j + m * q >= n) e[m] = make_double2(0.0, 0.0);
#pragma unroll
        for (int m = 4; m < 8; ++m) e[m] = make_double2(0.0, 0.0);
        bfly8_dif(e, w1, w2, w4);
#pragma unroll
        for (int m = 0; m < 8; ++m) xs[PADI(j + m * q)] = e[m];
    }
    __syncthreads();
}

}  // namespace gphm
