// Fused Gram / derivative-Gram construction (HBM-bound: algorithmic traffic = the N^2 doubles
// written per output matrix; the mixture parameters live in shared memory).
//
// Drop-in for   Kernel_matrix.get_kernel_matrix              kernel_matrix.py:21-30
//               vmap(cov_func.DD_x1_kappa) / D_x1_kappa      model_GP_solver_2d.py:107-117,
//                                                            model_GP_solver_advection.py:107-117
//               rectangular cross-Grams of preds             model_GP_solver_2d.py:198-202
// Two paths:
//  * general  - arbitrary x1 (n1), x2 (n2): every entry evaluates the Q-term mixture (N^2*Q
//               exp/sincos in FP64).  K and its derivative Gram share the transcendentals.
//  * Toeplitz - uniform collocation grids (np.linspace, model_GP_solver_2d.py:369-374): the
//               Gram is symmetric Toeplitz, so only n table entries are evaluated and the fill
//               kernel is a pure streaming write with 128-bit stores.
#include <type_traits>
#include "common.cuh"
#include "kernfun.cuh"
#include "kernels.h"

namespace gphm {

constexpr int kMaxQ = 256;

template <int KID>
__device__ __forceinline__ void load_comps(CompConst* sc, const double* __restrict__ theta, int Q) {
    for (int q = threadIdx.x + threadIdx.y * blockDim.x; q < Q; q += blockDim.x * blockDim.y)
        sc[q] = make_comp(KID, theta[q], theta[Q + q], theta[2 * Q + q]);
    __syncthreads();
}

// block (64,4): thread -> row i, column pair (j0, j0+1)
template <int KID, int ORDER>
__global__ void __launch_bounds__(256)
gram_general_kernel(const double* __restrict__ x1, int n1, const double* __restrict__ x2, int n2,
                    const double* __restrict__ theta, int Q, double jitter,
                    double* __restrict__ Kout, double* __restrict__ Dout, int ld) {
    __shared__ CompConst sc[kMaxQ];
    load_comps<KID>(sc, theta, Q);
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n1 || j0 >= n2) return;
    const bool two = (j0 + 1 < n2);
    const double xi = x1[i];
    const double df0 = xi - x2[j0];
    const double df1 = two ? xi - x2[j0 + 1] : 0.0;
    const double d0 = fabs(df0), d1 = fabs(df1);
    double k0 = 0.0, k1 = 0.0, g0 = 0.0, g1 = 0.0;
    for (int q = 0; q < Q; ++q) {
        const CompConst c = sc[q];
        double a, b;
        comp_value<KID, ORDER>(d0, c, a, b); k0 += a; g0 += b;
        comp_value<KID, ORDER>(d1, c, a, b); k1 += a; g1 += b;
    }
    if (ORDER == 1) {                       // k'(d) * sgn(x1 - x2), sgn(0) = +1
        if (df0 < 0.0) g0 = -g0;
        if (df1 < 0.0) g1 = -g1;
    }
    if (jitter != 0.0) { if (i == j0) k0 += jitter; if (i == j0 + 1) k1 += jitter; }
    const size_t o = (size_t)i * ld + j0;
    const bool vec = two && ((ld & 1) == 0);
    if (Kout) { if (vec) *reinterpret_cast<double2*>(Kout + o) = make_double2(k0, k1); else { Kout[o] = k0; if (two) Kout[o + 1] = k1; } }
    if (Dout) { if (vec) *reinterpret_cast<double2*>(Dout + o) = make_double2(g0, g1); else { Dout[o] = g0; if (two) Dout[o + 1] = g1; } }
}

// Square Gram pair on ONE grid (x1 == x2): K is symmetric and the derivative Gram symmetric (ORDER 2) or antisymmetric
// (ORDER 1), so only the 32 x 32 tiles on and above the diagonal evaluate the Q-term mixture - half the exp / sincos of the
// full kernel, which is what bounds it (N^2 Q FP64 transcendentals: 2.7 ms per 4096^2 pair against 0.04 ms of HBM writes).
// The mirrored tile leaves through a shared-memory transpose so that both stores are coalesced.
// blockIdx.x enumerates the upper-triangular tile pairs (ti <= tj); block (32, 8).
template <int KID, int ORDER>
__global__ void __launch_bounds__(256)
gram_symmetric_kernel(const double* __restrict__ x, int n, const double* __restrict__ theta, int Q, double jitter,
                      double* __restrict__ Kout, double* __restrict__ Dout, int ld, int ntiles) {
    __shared__ CompConst sc[kMaxQ];
    __shared__ double tK[32][33], tD[32][33];
    load_comps<KID>(sc, theta, Q);
    // tile pair from the linear index: row ti has (ntiles - ti) tiles
    int rem = blockIdx.x, ti = 0;
    while (rem >= ntiles - ti) { rem -= ntiles - ti; ++ti; }
    const int tj = ti + rem;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j = tj * 32 + tx;
    const double xj = j < n ? x[j] : 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int a = ty + 8 * r, i = ti * 32 + a;
        double k0 = 0.0, g0 = 0.0;
        if (i < n && j < n) {
            const double df = x[i] - xj, d = fabs(df);
            for (int q = 0; q < Q; ++q) {
                double u, v;
                comp_value<KID, ORDER>(d, sc[q], u, v);
                k0 += u; g0 += v;
            }
            if (ORDER == 1 && df < 0.0) g0 = -g0;
            if (i == j) k0 += jitter;
            const size_t o = (size_t)i * ld + j;
            Kout[o] = k0; Dout[o] = g0;
        }
        tK[a][tx] = k0; tD[a][tx] = g0;
    }
    if (ti == tj) return;                                   // diagonal tiles are complete
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {                           // mirrored tile: entry (j', i') = +-entry (i', j')
        const int a = ty + 8 * r;
        const int jj = tj * 32 + a, ii = ti * 32 + tx;
        if (jj < n && ii < n) {
            const size_t o = (size_t)jj * ld + ii;
            Kout[o] = tK[tx][a];
            Dout[o] = ORDER == 1 ? -tD[tx][a] : tD[tx][a];
        }
    }
}

template <int KID, int ORDER>
__global__ void __launch_bounds__(128)
toeplitz_table_kernel(const double* __restrict__ x, int n, const double* __restrict__ theta, int Q,
                      double* __restrict__ tabK, double* __restrict__ tabD, const int* __restrict__ skip) {
    if (skip && *skip) return;             // the look-ahead of the previous step already factored this theta (plan.cu)
    __shared__ CompConst sc[kMaxQ];
    load_comps<KID>(sc, theta, Q);
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const double d = fabs(x[m] - x[0]);
    double k0 = 0.0, g0 = 0.0;
    for (int q = 0; q < Q; ++q) {
        double a, b;
        comp_value<KID, ORDER>(d, sc[q], a, b);
        k0 += a; g0 += b;
    }
    tabK[m] = k0;
    tabD[m] = g0;
}

// Streaming fill: K[i,j] = tabK[|i-j|] (+jitter on the diagonal), D[i,j] = tabD[|i-j|] (* sign).
// One thread writes 2 consecutive columns of both matrices for ROWS consecutive rows.
template <bool ANTISYM>
__global__ void __launch_bounds__(256)
toeplitz_fill_kernel(const double* __restrict__ tabK, const double* __restrict__ tabD, int n, double jitter,
                     double dirsign, double* __restrict__ Kout, double* __restrict__ Dout, int ld) {
    constexpr int ROWS = 4;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int i0 = blockIdx.y * ROWS;
    if (j0 >= n) return;
    const bool two = (j0 + 1 < n);
    const bool vec = two && ((ld & 1) == 0);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int i = i0 + r;
        if (i >= n) break;
        const int l0 = i - j0, l1 = i - j0 - 1;
        const int a0 = l0 < 0 ? -l0 : l0, a1 = l1 < 0 ? -l1 : l1;
        double k0 = __ldg(tabK + a0), g0 = __ldg(tabD + a0);
        double k1 = two ? __ldg(tabK + a1) : 0.0, g1 = two ? __ldg(tabD + a1) : 0.0;
        if (ANTISYM) { g0 *= (l0 >= 0 ? dirsign : -dirsign); g1 *= (l1 >= 0 ? dirsign : -dirsign); }
        if (l0 == 0) k0 += jitter;
        if (l1 == 0) k1 += jitter;
        const size_t o = (size_t)i * ld + j0;
        if (vec) {
            *reinterpret_cast<double2*>(Kout + o) = make_double2(k0, k1);
            *reinterpret_cast<double2*>(Dout + o) = make_double2(g0, g1);
        } else {
            Kout[o] = k0; Dout[o] = g0;
            if (two) { Kout[o + 1] = k1; Dout[o + 1] = g1; }
        }
    }
}

// Element-wise ("vmapped") evaluation over explicit pair lists: out[p] = d^ORDER/dx1^ORDER kappa(x1[p], x2[p]).
template <int KID, int ORDER>
__global__ void __launch_bounds__(256)
kappa_pairs_kernel(const double* __restrict__ x1, const double* __restrict__ x2, size_t np,
                   const double* __restrict__ theta, int Q, double* __restrict__ out) {
    __shared__ CompConst sc[kMaxQ];
    load_comps<KID>(sc, theta, Q);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (size_t)gridDim.x * blockDim.x) {
        const double df = x1[i] - x2[i];
        const double d = fabs(df);
        double k0 = 0.0, g0 = 0.0;
        for (int q = 0; q < Q; ++q) {
            double a, b;
            comp_value<KID, ORDER>(d, sc[q], a, b);
            k0 += a; g0 += b;
        }
        if (ORDER == 1 && df < 0.0) g0 = -g0;
        out[i] = ORDER == 0 ? k0 : g0;
    }
}

int launch_kappa_pairs(int kid, int order, const double* x1, const double* x2, size_t np, const double* theta, int Q,
                       double* out, cudaStream_t st) {
    if (Q > kMaxQ || Q < 1) { set_last_error("kappa: Q=%d outside [1,%d]", Q, kMaxQ); return GPHM_EINVAL; }
    if (np == 0) return GPHM_OK;
    const int blocks = (int)((np + 255) / 256 < (size_t)kNumSMs * 8 ? (np + 255) / 256 : (size_t)kNumSMs * 8);
    LaunchScope scope(CAT_GRAM, st, 0.0, 24.0 * np);
    int rc = GPHM_DISPATCH_KID_ORDER(kid, order,
        kappa_pairs_kernel<KID, ORDER><<<blocks, 256, 0, st>>>(x1, x2, np, theta, Q, out));
    if (rc != 0) { set_last_error("kappa: bad kernel id %d / order %d", kid, order); return GPHM_EINVAL; }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_gram_general(int kid, int order, const double* x1, int n1, const double* x2, int n2,
                        const double* theta, int Q, double jitter, double* Kout, double* Dout, int ld,
                        cudaStream_t st) {
    if (Q > kMaxQ || Q < 1) { set_last_error("gram: Q=%d outside [1,%d]", Q, kMaxQ); return GPHM_EINVAL; }
    if (n1 <= 0 || n2 <= 0) return GPHM_OK;
    if (x1 == x2 && n1 == n2 && Kout && Dout && order > 0) {             // the plan's own (K, D) pair: evaluate one triangle only
        const int nt = (n1 + 31) / 32;
        LaunchScope scope(CAT_GRAM, st, 0.0, 16.0 * n1 * (double)n2);
        int rc = GPHM_DISPATCH_KID_ORDER(kid, order,
            gram_symmetric_kernel<KID, ORDER><<<nt * (nt + 1) / 2, dim3(32, 8), 0, st>>>(x1, n1, theta, Q, jitter, Kout, Dout, ld, nt));
        if (rc != 0) { set_last_error("gram: bad kernel id %d / order %d", kid, order); return GPHM_EINVAL; }
        GPHM_LAUNCH_OK();
        return GPHM_OK;
    }
    dim3 block(64, 4), grid((n2 + 127) / 128, (n1 + 3) / 4);
    LaunchScope scope(CAT_GRAM, st, 0.0, 8.0 * n1 * n2 * ((Kout ? 1 : 0) + (Dout ? 1 : 0)));
    int rc = GPHM_DISPATCH_KID_ORDER(kid, order,
        gram_general_kernel<KID, ORDER><<<grid, block, 0, st>>>(x1, n1, x2, n2, theta, Q, jitter, Kout, Dout, ld));
    if (rc != 0) { set_last_error("gram: bad kernel id %d / order %d", kid, order); return GPHM_EINVAL; }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_toeplitz_table(int kid, int order, const double* x, int n, const double* theta, int Q, double* tabK, double* tabD,
                          cudaStream_t st, const int* skip) {
    if (Q > kMaxQ || Q < 1) { set_last_error("gram: Q=%d outside [1,%d]", Q, kMaxQ); return GPHM_EINVAL; }
    if (n <= 0) return GPHM_OK;
    int rc;
    { LaunchScope scope(CAT_GRAM, st, 0.0, 24.0 * n);
    rc = GPHM_DISPATCH_KID_ORDER(kid, order,
        toeplitz_table_kernel<KID, ORDER><<<(n + 127) / 128, 128, 0, st>>>(x, n, theta, Q, tabK, tabD, skip)); }
    if (rc != 0) { set_last_error("gram: bad kernel id %d / order %d", kid, order); return GPHM_EINVAL; }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_gram_toeplitz(int kid, int order, const double* x, int n, const double* theta, int Q, double jitter,
                         double dirsign, double* tabK, double* tabD, double* Kout, double* Dout, int ld,
                         cudaStream_t st) {
    if (Q > kMaxQ || Q < 1) { set_last_error("gram: Q=%d outside [1,%d]", Q, kMaxQ); return GPHM_EINVAL; }
    if (n <= 0) return GPHM_OK;
    int rc;
    { LaunchScope scope(CAT_GRAM, st, 0.0, 24.0 * n);
    rc = GPHM_DISPATCH_KID_ORDER(kid, order,
        toeplitz_table_kernel<KID, ORDER><<<(n + 127) / 128, 128, 0, st>>>(x, n, theta, Q, tabK, tabD, nullptr)); }
    if (rc != 0) { set_last_error("gram: bad kernel id %d / order %d", kid, order); return GPHM_EINVAL; }
    GPHM_LAUNCH_OK();
    dim3 block(256), grid((n + 511) / 512, (n + 3) / 4);
    LaunchScope scope(CAT_GRAM, st, 0.0, 16.0 * n * n);
    if (order == 1) toeplitz_fill_kernel<true><<<grid, block, 0, st>>>(tabK, tabD, n, jitter, dirsign, Kout, Dout, ld);
    else toeplitz_fill_kernel<false><<<grid, block, 0, st>>>(tabK, tabD, n, jitter, dirsign, Kout, Dout, ld);
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
