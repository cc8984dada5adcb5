// Row-wise Toeplitz products on uniform grids with fused shared-memory sweeps (FP64).
//
//   toeplitz_apply_fused_kernel   Out[r] = alpha * T X[r] (+ beta * Out[r]): one circulant convolution
//       per row (derivative-Gram products D A, D^T G, Bt D^T, G D; jnp.matmul at
//       model_GP_solver_2d.py:112,119 and their reverse pass).
//   gs_apply_fused_kernel         Out[r] = alpha * K^-1 X[r] (+ beta * Add[r]): the Gohberg-Semencul
//       formula  K^-1 v = ( L(g) [L(g)^T v]_n - L(h) [L(h)^T v]_n ) / g0  (jnp.linalg.solve,
//       :104-105, and the solve-VJPs) as ONE pass over the row pair:
//           Z = F z;  Q1 = F [F^-1 (conj G . Z)]_n;  Q2 = F [F^-1 (conj H . Z)]_n;
//           out = F^-1 ( G' . Q1 + H' . Q2 )            (G' = G/(L g0), H' = -H/(L g0))
//       six transforms; Z and then G'.Q1 wait in registers (16 complex values per thread).
// Two real rows ride in one complex transform (every operator here is real).  Transforms run in
// place in shared memory with the pass structure of fft_core.cuh: the global load is fused into
// the first pass, the spectrum product into the tail/head pass, the truncation between two
// convolutions into one register butterfly, the global store into the last pass - 2 np8 + 1
// sweeps per convolution (9 for L = 8192) instead of 12, and 24 instead of 48 for K^-1.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "fft_core.cuh"

namespace gphm {

namespace {

// L2 prefetch of a contiguous range (one 128-byte line per thread and iteration): issued a whole row pair ahead,
// so that the first pass of the next pair (and the stored spectra of xcorr_pairs) find their operands in L2.
template <int NT>
__device__ __forceinline__ void prefetch_l2(const void* p, size_t bytes, int tid) {
    const char* c = static_cast<const char*>(p);
    for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)fft_nt<NT>() * 128)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(c + off));
}

struct RowPairIO {
    const double* x0; const double* x1; double* o0; double* o1; const double* a0; const double* a1;
    int n; bool two; double alpha, beta;
    __device__ __forceinline__ double2 load(int idx) const {
        if (idx >= n) return make_double2(0.0, 0.0);
        return make_double2(x0[idx], two ? x1[idx] : 0.0);
    }
    __device__ __forceinline__ double2 addend(int idx) const {
        if (beta == 0.0 || idx >= n) return make_double2(0.0, 0.0);
        return make_double2(a0[idx], two ? a1[idx] : 0.0);
    }
    template <int NT>
    __device__ __forceinline__ void prefetch(int tid) const {
        prefetch_l2<NT>(x0, sizeof(double) * n, tid);
        if (two) prefetch_l2<NT>(x1, sizeof(double) * n, tid);
        if (beta != 0.0) { prefetch_l2<NT>(a0, sizeof(double) * n, tid); if (two) prefetch_l2<NT>(a1, sizeof(double) * n, tid); }
    }
    __device__ __forceinline__ void store(int idx, double2 v, double2 add) const {
        if (idx >= n) return;
        o0[idx] = fma(beta, add.x, alpha * v.x);
        if (two) o1[idx] = fma(beta, add.y, alpha * v.y);
    }
};

__device__ __forceinline__ RowPairIO row_pair(const double* X, int ldx, double* Out, int ldo, const double* Add, int lda,
                                              int rows, int n, int pr, double alpha, double beta) {
    RowPairIO io;
    const int r0 = 2 * pr, r1 = r0 + 1;
    io.two = r1 < rows;
    io.x0 = X + (size_t)r0 * ldx; io.x1 = X + (size_t)(io.two ? r1 : r0) * ldx;
    io.o0 = Out + (size_t)r0 * ldo; io.o1 = Out + (size_t)(io.two ? r1 : r0) * ldo;
    io.a0 = Add ? Add + (size_t)r0 * lda : io.o0; io.a1 = Add ? Add + (size_t)(io.two ? r1 : r0) * lda : io.o1;
    io.n = n; io.alpha = alpha; io.beta = beta;
    return io;
}

}  // namespace

template <int KT, int NT, bool GR>
__global__ void __launch_bounds__(FFT_THREADS, 1)
toeplitz_apply_fused_kernel(const double* X, int rows, int n, int ldx, const double2* __restrict__ spec, int L,
                            int logL, const double2* __restrict__ W, double alpha, double beta, const double* Add, int lda,
                            double* Out, int ldo, double2* __restrict__ SpecOut) {
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles<NT>(xs, L, logL, W, tid);
    const int np8 = (logL - KT) / 3;
    const double2* tw0 = fft_twiddles(xs, L);
    const int npairs = (rows + 1) / 2;
    for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const RowPairIO io = row_pair(X, ldx, Out, ldo, Add, lda, rows, n, pr, alpha, beta);
        if (pr + (int)gridDim.x < npairs) row_pair(X, ldx, Out, ldo, Add, lda, rows, n, pr + gridDim.x, alpha, beta).template prefetch<NT>(tid);
        dif_first<NT>(xs, L, tw0, tid, [&](int idx) { return io.load(idx); });
        dif_middle<NT, GR>(xs, L, logL, np8, tid);
        double2* so = SpecOut ? SpecOut + (size_t)pr * L : nullptr;      // spectrum of the packed row pair, kept for the diagonal sums
        mid_fused<KT, NT, GR>(xs, L, logL, tid, [&](int, int p, double2 v) { if (so) so[p] = v; return cmul(v, spec[p]); });
        dit_middle<NT, GR>(xs, L, logL, np8, KT, tid);
        dit_last<NT>(xs, L, tw0, tid, [&](int idx) { return io.addend(idx); }, [&](int idx, double2 v, double2 add) { io.store(idx, v, add); });
    }
}

// spec: the four Gohberg-Semencul spectra of gs_prepare_kernel, L complex values each.
// ONE (L == 2n, g0ptr given): only spec[0] = conj(G)/L is read; with H = (-1)^k (conj(G) - g0) (h is g reversed and shifted,
// exp(-2 pi i n k / L) = (-1)^k when L = 2n) the other three follow from it and g0:
//     conj(H)/L = sgn (conj(a) - g0/L),   G/(L g0) = conj(a)/g0,   -H/(L g0) = -sgn (a - g0/L)/g0,   a = spec[0][p],
//     sgn = (-1)^k = +1 for bin positions p < L/2 (bit-reversed order: the lowest bit of k is the highest bit of p)
// - 384 KB instead of 512 KB of spectra through L2 per row pair, and one array that stays hot.
// VAR 0: four spectra, loads where they are used (round 1);  1: ONE + spectrum prefetch + dif_first2;  2: ONE + prefetch;
// 3: ONE only;  4: four spectra + prefetch.   (variants > 0 are instantiated for the 8192-point configuration only)
template <int KT, int NT, bool GR, int VAR>
__global__ void __launch_bounds__(FFT_THREADS, 1)
gs_apply_fused_kernel(const double* X, int rows, int n, int ldx, const double2* __restrict__ spec, int L,
                      int logL, const double2* __restrict__ W, double alpha, double beta, const double* Add, int lda,
                      double* Out, int ldo, const double* __restrict__ g0ptr) {
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles<NT>(xs, L, logL, W, tid);
    const int np8 = (logL - KT) / 3;
    const double2* tw0 = fft_twiddles(xs, L);
    const double2* __restrict__ sGt = spec;             // conj(G)/L
    const double2* __restrict__ sHt = spec + L;         // conj(H)/L
    const double2* __restrict__ sG = spec + 2 * (size_t)L;   //  G/(L g0)
    const double2* __restrict__ sH = spec + 3 * (size_t)L;   // -H/(L g0)
    constexpr bool ONE = VAR >= 1 && VAR <= 3, PF = VAR == 1 || VAR == 2 || VAR == 4, F2 = VAR == 1;
    double c0 = 0.0, ig0 = 0.0;
    if (ONE) { const double g0 = *g0ptr; c0 = g0 / (double)L; ig0 = 1.0 / g0; }
    const int half = L >> 1;
    const int npairs = (rows + 1) / 2;
    double2 stash[FFT_ACC];
    for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const RowPairIO io = row_pair(X, ldx, Out, ldo, Add, lda, rows, n, pr, alpha, beta);
        if (pr + (int)gridDim.x < npairs) row_pair(X, ldx, Out, ldo, Add, lda, rows, n, pr + gridDim.x, alpha, beta).template prefetch<NT>(tid);
        if (F2) dif_first2<NT>(xs, L, tw0, tid, [&](int idx) { return io.load(idx); });
        else dif_first<NT>(xs, L, tw0, tid, [&](int idx) { return io.load(idx); });
        dif_middle<NT, GR>(xs, L, logL, np8, tid);
        mid_fused_spec<KT, NT, GR, PF>(xs, L, logL, tid, sGt, [&](int slot, int, double2 v, double2 a) { stash[slot] = v; return cmul(v, a); });   // Z kept
        dit_middle<NT, GR>(xs, L, logL, np8, KT, tid);
        dit_last_dif_first<NT>(xs, L, tw0, tid, n);                                   // [L(g)^T v]_n
        dif_middle<NT, GR>(xs, L, logL, np8, tid);
        if (ONE) {
            mid_fused_spec<KT, NT, GR, PF>(xs, L, logL, tid, sGt, [&](int slot, int p, double2 v, double2 a) {
                const double2 z = stash[slot];
                const double2 q1 = cmulc(v, a);                                        // v * conj(a) = v G / L
                stash[slot] = make_double2(q1.x * ig0, q1.y * ig0);                    // G'.Q1 kept
                const double sg = p < half ? 1.0 : -1.0;
                return cmul(z, make_double2(sg * (a.x - c0), -sg * a.y));              // Z . conj(H)/L
            });
        } else {
            mid_fused<KT, NT, GR>(xs, L, logL, tid, [&](int slot, int p, double2 v) {
                const double2 z = stash[slot];
                stash[slot] = cmul(v, sG[p]);                                         // G'.Q1 kept
                return cmul(z, sHt[p]);
            });
        }
        dit_middle<NT, GR>(xs, L, logL, np8, KT, tid);
        dit_last_dif_first<NT>(xs, L, tw0, tid, n);                                   // [L(h)^T v]_n
        dif_middle<NT, GR>(xs, L, logL, np8, tid);
        if (ONE) {
            mid_fused_spec<KT, NT, GR, PF>(xs, L, logL, tid, sGt, [&](int slot, int p, double2 v, double2 a) {
                const double sg = p < half ? -ig0 : ig0;                               // -(-1)^k / g0
                const double2 b = cmul(v, make_double2(sg * (a.x - c0), sg * a.y));    // Q2 . (-H/(L g0))
                const double2 s1 = stash[slot];
                return make_double2(s1.x + b.x, s1.y + b.y);
            });
        } else {
            mid_fused_spec<KT, NT, GR, PF>(xs, L, logL, tid, sH, [&](int slot, int, double2 v, double2 h) {
                const double2 a = stash[slot], b = cmul(v, h);
                return make_double2(a.x + b.x, a.y + b.y);
            });
        }
        dit_middle<NT, GR>(xs, L, logL, np8, KT, tid);
        dit_last<NT>(xs, L, tw0, tid, [&](int idx) { return io.addend(idx); }, [&](int idx, double2 v, double2 add) { io.store(idx, v, add); });
    }
}

// partial[blockIdx.x][p] = weight * sum over the CTA's row pairs of conj(Zx)(p) * Zy(p), where Zx is the
// transform of the packed pair (x_2r + i x_2r+1) computed here and Zy the stored transform of the
// matching packed pair of Y (SpecOut of toeplitz_apply_fused_kernel).  conj(Zx) Zy = conj(X0) Y0 +
// conj(X1) Y1 + i (conj(X0) Y1 - conj(X1) Y0): the cross term transforms back to a purely imaginary
// sequence, so the real part of the inverse transform of the sum is exactly the sum over ROWS of the
// cross-correlations - one transform per row pair and no spectrum separation.
template <int KT, int NT, bool GR>
__global__ void __launch_bounds__(FFT_THREADS, 1)
xcorr_pairs_kernel(const double* __restrict__ X, int rows, int n, int ldx, const double2* __restrict__ SpecY, int L, int logL,
                   const double2* __restrict__ W, double weight, double2* __restrict__ partial) {
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles<NT>(xs, L, logL, W, tid);
    const int np8 = (logL - KT) / 3;
    const double2* tw0 = fft_twiddles(xs, L);
    const int npairs = (rows + 1) / 2;
    constexpr int R = 1 << KT;
    constexpr int MAXG = FFT_MAX_L / (R * FFT_THREADS);
    double2 acc[FFT_ACC];
#pragma unroll
    for (int k = 0; k < FFT_ACC; ++k) acc[k] = make_double2(0.0, 0.0);
    for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const int r0 = 2 * pr;
        const bool two = r0 + 1 < rows;
        const double* x0 = X + (size_t)r0 * ldx;
        const double* x1 = X + (size_t)(two ? r0 + 1 : r0) * ldx;
        const double2* __restrict__ sy = SpecY + (size_t)pr * L;
        prefetch_l2<NT>(sy, sizeof(double2) * L, tid);                  // needed after the transform below
        if (pr + (int)gridDim.x < npairs) {
            prefetch_l2<NT>(X + (size_t)(r0 + 2 * gridDim.x) * ldx, sizeof(double) * n, tid);
            if (r0 + 2 * (int)gridDim.x + 1 < rows) prefetch_l2<NT>(X + (size_t)(r0 + 2 * gridDim.x + 1) * ldx, sizeof(double) * n, tid);
        }
        dif_first<NT>(xs, L, tw0, tid, [&](int idx) { return idx < n ? make_double2(x0[idx], two ? x1[idx] : 0.0) : make_double2(0.0, 0.0); });
        dif_middle<NT, GR>(xs, L, logL, np8, tid);
#pragma unroll
        for (int i = 0; i < MAXG; ++i) {                       // forward tail in registers + accumulation
            const int g = fft_mid_group<KT, NT, GR>(tid, i, L, logL - 6);
            if (g < (L >> KT)) {
                const int base = g << KT;
                double2 y[R], e[R];
#pragma unroll
                for (int m = 0; m < R; ++m) y[m] = sy[base + m];
#pragma unroll
                for (int m = 0; m < R; ++m) e[m] = xs[PADI(base + m)];
                unit_fwd<KT>(e);
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    acc[i * R + m].x += e[m].x * y[m].x + e[m].y * y[m].y;
                    acc[i * R + m].y += e[m].x * y[m].y - e[m].y * y[m].x;
                }
            }
        }
        __syncthreads();
    }
    double2* out = partial + (size_t)blockIdx.x * L;
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
        const int g = fft_mid_group<KT, NT, GR>(tid, i, L, logL - 6);
        if (g < (L >> KT)) {
#pragma unroll
            for (int m = 0; m < R; ++m) out[(g << KT) + m] = make_double2(weight * acc[i * R + m].x, weight * acc[i * R + m].y);
        }
    }
}

// (KT, NT, GR) dispatch: NT = FFT_THREADS (compile-time strides) for the full-size CTA, 0 (blockDim.x) for shorter
// transforms; GR = octant-group barriers (fft_core.cuh) for L >= 512
#define GPHM_FUSED_KT(K, NT, GR, GRID, ...)                                                                  \
    do {                                                                                                     \
        if (KT == 1) K<1, NT, GR><<<GRID, nt, smem, st>>>(__VA_ARGS__);                                      \
        else if (KT == 2) K<2, NT, GR><<<GRID, nt, smem, st>>>(__VA_ARGS__);                                 \
        else K<3, NT, GR><<<GRID, nt, smem, st>>>(__VA_ARGS__);                                              \
    } while (0)
#define GPHM_FUSED_LAUNCH(K, GRID, ...)                                                                      \
    do {                                                                                                     \
        const bool gr = fft_groups_supported(L) && !getenv_flag("GPHM_FFT_NO_GROUPS");                      \
        if (nt == FFT_THREADS) {                                                                             \
            if (gr) GPHM_FUSED_KT(K, FFT_THREADS, true, GRID, __VA_ARGS__);                                  \
            else GPHM_FUSED_KT(K, FFT_THREADS, false, GRID, __VA_ARGS__);                                    \
        } else {                                                                                             \
            if (gr) GPHM_FUSED_KT(K, 0, true, GRID, __VA_ARGS__);                                            \
            else GPHM_FUSED_KT(K, 0, false, GRID, __VA_ARGS__);                                              \
        }                                                                                                    \
    } while (0)

constexpr int kGsDefaultVariant = 0;          // see gs_apply_fused_kernel; GPHM_GS_VARIANT overrides

static bool getenv_flag(const char* name) {       // read once per name would need a map: two callers, cheap enough
    const char* v = getenv(name);
    return v && v[0] && v[0] != '0';
}

static int ilog2f(int L) { int l = 0; while ((1 << l) < L) ++l; return l; }

int toeplitz_fused_init() {
    static DeviceOnce once;
    if (!once.needed()) return GPHM_OK;
    const int bytes = (int)fft_smem_bytes(FFT_MAX_L);
#define GPHM_FUSED_ATTR(K, KT, NT, GR) GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(K<KT, NT, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
#define GPHM_FUSED_ATTR3(K, NT, GR) GPHM_FUSED_ATTR(K, 1, NT, GR); GPHM_FUSED_ATTR(K, 2, NT, GR); GPHM_FUSED_ATTR(K, 3, NT, GR)
#define GPHM_FUSED_ATTR6(K) GPHM_FUSED_ATTR3(K, FFT_THREADS, false); GPHM_FUSED_ATTR3(K, 0, false); \
                            GPHM_FUSED_ATTR3(K, FFT_THREADS, true); GPHM_FUSED_ATTR3(K, 0, true)
    GPHM_FUSED_ATTR6(toeplitz_apply_fused_kernel);
    GPHM_FUSED_ATTR(toeplitz_apply_fused_kernel, 4, FFT_THREADS, false); GPHM_FUSED_ATTR(toeplitz_apply_fused_kernel, 4, 0, false);
    GPHM_FUSED_ATTR(toeplitz_apply_fused_kernel, 4, FFT_THREADS, true); GPHM_FUSED_ATTR(toeplitz_apply_fused_kernel, 4, 0, true);
    GPHM_FUSED_ATTR6(xcorr_pairs_kernel);
#define GPHM_GS_ATTR(KT, NT, GR) GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(gs_apply_fused_kernel<KT, NT, GR, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
#define GPHM_GS_ATTR3(NT, GR) GPHM_GS_ATTR(1, NT, GR); GPHM_GS_ATTR(2, NT, GR); GPHM_GS_ATTR(3, NT, GR)
    GPHM_GS_ATTR3(FFT_THREADS, false); GPHM_GS_ATTR3(0, false); GPHM_GS_ATTR3(FFT_THREADS, true); GPHM_GS_ATTR3(0, true);
#undef GPHM_GS_ATTR3
#undef GPHM_GS_ATTR
#define GPHM_GS_ATTRV(V) GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(gs_apply_fused_kernel<1, FFT_THREADS, true, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
    GPHM_GS_ATTRV(1); GPHM_GS_ATTRV(2); GPHM_GS_ATTRV(3); GPHM_GS_ATTRV(4);
#undef GPHM_GS_ATTRV
#undef GPHM_FUSED_ATTR3
#undef GPHM_FUSED_ATTR6
#undef GPHM_FUSED_ATTR
    once.done();
    return GPHM_OK;
}

bool toeplitz_fused_supported(int L) { return L >= 16 && L <= FFT_MAX_L; }

// Out[r] = alpha * T X[r] + beta * Add[r] (Add == NULL: beta * Out[r]).  Out may alias X or Add (each CTA
// reads its row pair completely before it writes it).
int launch_toeplitz_apply_fused(const double* X, int rows, int n, int ldx, const double* spec, int L, const double* W,
                                double alpha, double beta, const double* Add, int lda, double* Out, int ldo, double* SpecOut,
                                cudaStream_t st) {
    GPHM_TRY(toeplitz_fused_init());
    if (rows <= 0) return GPHM_OK;
    if (!toeplitz_fused_supported(L) || L < 2 * n) { set_last_error("toeplitz_apply_fused: L=%d does not fit n=%d", L, n); return GPHM_EINVAL; }
    // GPHM_FFT_KT4=1: 16-point tail / head (7 instead of 9 sweeps at L = 8192).  Measured in round 2 and left OFF: the fused middle
    // pass (two 16-point transforms, 16 spectrum products, 128 registers with spills) runs 1.46 -> 1.71 ms per step.
    static const bool kt4 = getenv_flag("GPHM_FFT_KT4");
    const int logL = ilog2f(L), KT = kt4 ? fft_tail_stages_wide(logL) : fft_tail_stages(logL);
    const int nt = fft_threads_for(L);
    const int grid = std::min(fft_grid() * (FFT_THREADS / nt), (rows + 1) / 2);
    const size_t smem = fft_smem_bytes(L);
    auto sp = reinterpret_cast<const double2*>(spec);
    auto w = reinterpret_cast<const double2*>(W);
    auto so = reinterpret_cast<double2*>(SpecOut);
    {
        const double pairs = (rows + 1) / 2;
        LaunchScope scope(CAT_TOEPLITZ_APPLY, st, pairs * (2.0 * fft_flops(L) + 6.0 * L),
                          (beta != 0.0 ? 24.0 : 16.0) * rows * (double)n + (SpecOut ? 16.0 * pairs * L : 0.0));
        if (KT == 4) {          // 16-point tail / head in registers: 7 instead of 9 shared-memory sweeps at L = 8192 (5 instead of 7 at 1024)
            const bool gr = fft_groups_supported(L) && !getenv_flag("GPHM_FFT_NO_GROUPS");
#define GPHM_KT4(NT, GR) toeplitz_apply_fused_kernel<4, NT, GR><<<grid, nt, smem, st>>>(X, rows, n, ldx, sp, L, logL, w, alpha, beta, Add, lda, Out, ldo, so)
            if (nt == FFT_THREADS) { if (gr) GPHM_KT4(FFT_THREADS, true); else GPHM_KT4(FFT_THREADS, false); }
            else { if (gr) GPHM_KT4(0, true); else GPHM_KT4(0, false); }
#undef GPHM_KT4
        } else {
            GPHM_FUSED_LAUNCH(toeplitz_apply_fused_kernel, grid, X, rows, n, ldx, sp, L, logL, w, alpha, beta, Add, lda, Out, ldo, so);
        }
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// Out[r] = alpha * K^-1 X[r] + beta * Add[r]  (Add may be NULL when beta == 0; Out may alias X or Add).
template <class... Args>
static void gs_dispatch(int KT, int nt, bool gr, int grid, size_t smem, cudaStream_t st, Args... args) {
#define GPHM_GS_KT(NT, GR)                                                                         \
    do {                                                                                           \
        if (KT == 1) gs_apply_fused_kernel<1, NT, GR, 0><<<grid, nt, smem, st>>>(args...);         \
        else if (KT == 2) gs_apply_fused_kernel<2, NT, GR, 0><<<grid, nt, smem, st>>>(args...);    \
        else gs_apply_fused_kernel<3, NT, GR, 0><<<grid, nt, smem, st>>>(args...);                 \
    } while (0)
    if (nt == FFT_THREADS) { if (gr) GPHM_GS_KT(FFT_THREADS, true); else GPHM_GS_KT(FFT_THREADS, false); }
    else { if (gr) GPHM_GS_KT(0, true); else GPHM_GS_KT(0, false); }
#undef GPHM_GS_KT
}

int launch_gs_apply_fused(const double* X, int rows, int n, int ldx, const double* gspec, int L, const double* W, double alpha,
                          double beta, const double* Add, int lda, double* Out, int ldo, cudaStream_t st, const double* g0ptr) {
    GPHM_TRY(toeplitz_fused_init());
    if (rows <= 0) return GPHM_OK;
    if (!toeplitz_fused_supported(L) || L < 2 * n) { set_last_error("gs_apply_fused: L=%d does not fit n=%d", L, n); return GPHM_EINVAL; }
    if (beta != 0.0 && !Add) { set_last_error("gs_apply_fused: beta without Add"); return GPHM_EINVAL; }
    const int logL = ilog2f(L), KT = fft_tail_stages(logL);
    const int nt = fft_threads_for(L);
    const int grid = std::min(fft_grid() * (FFT_THREADS / nt), (rows + 1) / 2);
    const size_t smem = fft_smem_bytes(L);
    auto sp = reinterpret_cast<const double2*>(gspec);
    auto w = reinterpret_cast<const double2*>(W);
    {
        const double pairs = (rows + 1) / 2;
        LaunchScope scope(CAT_GS_APPLY, st, pairs * (6.0 * fft_flops(L) + 18.0 * L), (beta != 0.0 ? 24.0 : 16.0) * rows * (double)n);
        const bool gr = fft_groups_supported(L) && !getenv_flag("GPHM_FFT_NO_GROUPS");
        static const int variant = [] { const char* e = getenv("GPHM_GS_VARIANT"); return e ? atoi(e) : kGsDefaultVariant; }();
        const bool hot = KT == 1 && nt == FFT_THREADS && gr;                 // the 8192-point configuration
        const bool one_ok = g0ptr && L == 2 * n;
        int v = hot ? variant : 0;
        if (!one_ok && v >= 1 && v <= 3) v = (v == 3) ? 0 : 4;
#define GPHM_GS_V(V) gs_apply_fused_kernel<1, FFT_THREADS, true, V><<<grid, nt, smem, st>>>(X, rows, n, ldx, sp, L, logL, w, alpha, beta, Add, lda, Out, ldo, g0ptr)
        if (v == 1) GPHM_GS_V(1); else if (v == 2) GPHM_GS_V(2); else if (v == 3) GPHM_GS_V(3); else if (v == 4) GPHM_GS_V(4);
        else gs_dispatch(KT, nt, gr, grid, smem, st, X, rows, n, ldx, sp, L, logL, w, alpha, beta, Add, lda, Out, ldo, g0ptr);
#undef GPHM_GS_V
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// partial[fft_grid()][L] = weight * sum over row pairs of conj(FFT(packed X pair)) * SpecY[pair]  (every CTA writes its slot)
int launch_xcorr_pairs(const double* X, int rows, int n, int ldx, const double* SpecY, int L, const double* W, double weight,
                       double* partial, cudaStream_t st) {
    GPHM_TRY(toeplitz_fused_init());
    if (!toeplitz_fused_supported(L) || L < 2 * n) { set_last_error("xcorr_pairs: L=%d does not fit n=%d", L, n); return GPHM_EINVAL; }
    const int logL = ilog2f(L), KT = fft_tail_stages(logL);
    const size_t smem = fft_smem_bytes(L);
    const int nt = fft_threads_for(L);
    auto sy = reinterpret_cast<const double2*>(SpecY);
    auto w = reinterpret_cast<const double2*>(W);
    auto pt = reinterpret_cast<double2*>(partial);
    {
        const double pairs = (rows + 1) / 2;
        LaunchScope scope(CAT_FFT, st, pairs * (fft_flops(L) + 8.0 * L), 8.0 * rows * (double)n + 16.0 * pairs * L);
        GPHM_FUSED_LAUNCH(xcorr_pairs_kernel, fft_grid(), X, rows, n, ldx, sy, L, logL, w, weight, pt);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
