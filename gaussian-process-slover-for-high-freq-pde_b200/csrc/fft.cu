// Diagonal sums of X^T Y by batched FFT cross-correlation (uniform-grid theta-gradient path).
//
// On a uniform grid the Gram matrices are Toeplitz, so the reverse pass only needs
//     s[m] = sum_{col-row = m} (X^T Y)[row][col] = sum_r sum_j X[r][j] Y[r][j+m]      (m = -(C-1)..C-1)
// of  Kbar = beta K^-1 - V^T Y  and  Dbar = c G^T Y  (and K^-1 = Linv^T Linv), never the matrices
// themselves.  s is the sum over rows r of the cross-correlation of row r of X with row r of Y.
// This replaces the 4 full Kbar/Dbar GEMMs and the Linv^T Linv product per step (9.3 N^3 of the
// 28 N^3 FLOPs; the reverse pass of jnp.linalg.solve / slogdet / matmul,
// model_GP_solver_2d.py:104-119,158-162,179) by O(N^2 log N) work that is bound by shared-memory/FP64-ALU throughput and reads each operand
// once from HBM.
//
// Per row: z = x + i y zero-padded to L >= 2C, one in-place DIF FFT in shared memory (natural in,
// bit-reversed out; register-blocked radix-8 passes, i.e. 5 shared-memory round trips for
// L = 8192 instead of 13, on a padded layout that keeps every 128-bit access conflict-free);
// the two real spectra are separated with Z(f), conj Z(L-f);
// conj(X^)(f) Y^(f) is accumulated in registers over all rows a CTA owns (bit-reversed order).
// A second kernel sums the per-CTA partial spectra in a fixed order (deterministic), runs one
// inverse DIT FFT (bit-reversed in, natural out) and emits the symmetric / antisymmetric
// diagonal sums that theta_grad_toeplitz_kernel consumes.
#include <algorithm>
#include "common.cuh"
#include "kernels.h"
#include "fft_core.cuh"

namespace gphm {

int fft_length_for(int n) {                            // smallest power of two >= 2n (0 if unsupported)
    int L = 2;
    while (L < 2 * n) L <<= 1;
    return L <= FFT_MAX_L ? L : 0;
}
int fft_grid() { return kNumSMs; }

// Per-stage compact twiddle tables: stage s (butterfly span L >> (s+1)) reads
//   W[(L - (L >> s)) + j] = exp(-2 pi i (j << s) / L),  j < L >> (s+1)
// so a warp's loads are contiguous (a single strided table cost 4x the FFT time in L1 gathers).
__global__ void twiddle_init_kernel(double2* __restrict__ W, int L, int logL) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L - 1) return;
    int s = 0, off = 0;
    while (t >= off + (L >> (s + 1))) { off += L >> (s + 1); ++s; }
    const int j = t - off;
    double sn, cs;
    sincospi(2.0 * (double)((long long)j << s) / (double)L, &sn, &cs);
    W[t] = make_double2(cs, -sn);
}

// partial[blockIdx.x][p] (+)= weight * sum over the CTA's rows of conj(X^)(f_p) Y^(f_p)
__global__ void __launch_bounds__(FFT_THREADS, 1)
xcorr_spectrum_kernel(const double* __restrict__ X, const double* __restrict__ Y, int rows, int cols, int ldx,
                      int ldy, int L, int logL, const double2* __restrict__ W, double weight, int accumulate,
                      double2* __restrict__ partial) {
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles(xs, L, logL, W, tid);
    double2 acc[FFT_ACC];
#pragma unroll
    for (int k = 0; k < FFT_ACC; ++k) acc[k] = make_double2(0.0, 0.0);

    __shared__ double redmax[2][FFT_THREADS / 32];
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const double* xr = X + (size_t)r * ldx;
        const double* yr = Y + (size_t)r * ldy;
        double mx = 0.0, my = 0.0;
        for (int j = tid; j < L; j += FFT_THREADS) {
            double2 v = make_double2(0.0, 0.0);
            if (j < cols) { v = make_double2(xr[j], yr[j]); mx = fmax(mx, fabs(v.x)); my = fmax(my, fabs(v.y)); }
            xs[PADI(j)] = v;
        }
        // z = x + i*y shares one FFT: the two spectra are separated by a difference, so y is first
        // rescaled (exactly, by a power of two) to x's magnitude - otherwise the smaller sequence
        // loses log10(max|x|/max|y|) digits (|V| ~ 1e8 against |A| ~ 1e2 in this problem).
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            my = fmax(my, __shfl_xor_sync(0xffffffffu, my, o));
        }
        if ((tid & 31) == 0) { redmax[0][tid >> 5] = mx; redmax[1][tid >> 5] = my; }
        __syncthreads();
        mx = 0.0; my = 0.0;
#pragma unroll
        for (int w = 0; w < FFT_THREADS / 32; ++w) { mx = fmax(mx, redmax[0][w]); my = fmax(my, redmax[1][w]); }
        if (!(mx > 0.0) || !(my > 0.0) || !isfinite(mx) || !isfinite(my)) {
            if (!isfinite(mx) || !isfinite(my)) acc[0].x += mx + my;      // propagate NaN/Inf
            __syncthreads();
            continue;                                                     // a zero row contributes nothing
        }
        const double sc = exp2(rint(log2(mx / my)));
        const double isc = 1.0 / sc;
        for (int j = tid; j < cols; j += FFT_THREADS) xs[PADI(j)].y *= sc;
        __syncthreads();
        fft_dif_inplace(xs, L, logL, W, tid);
#pragma unroll
        for (int k = 0; k < FFT_ACC; ++k) {
            const int p = tid + k * FFT_THREADS;
            if (p < L) {
                const unsigned f = __brev((unsigned)p) >> (32 - logL);
                const unsigned fm = (unsigned)(L - (int)f) & (unsigned)(L - 1);
                const unsigned pm = __brev(fm) >> (32 - logL);
                const double2 zf = xs[PADI(p)], zm = xs[PADI((int)pm)];
                const double2 xh = make_double2(0.5 * (zf.x + zm.x), 0.5 * (zf.y - zm.y));       // X^(f)
                const double2 yh = make_double2(0.5 * (zf.y + zm.y), -0.5 * (zf.x - zm.x));      // Y^(f)
                acc[k].x += isc * (xh.x * yh.x + xh.y * yh.y);                                    // conj(X^) Y^
                acc[k].y += isc * (xh.x * yh.y - xh.y * yh.x);
            }
        }
        __syncthreads();
    }
    double2* out = partial + (size_t)blockIdx.x * L;
#pragma unroll
    for (int k = 0; k < FFT_ACC; ++k) {
        const int p = tid + k * FFT_THREADS;
        if (p < L) {
            double2 v = make_double2(weight * acc[k].x, weight * acc[k].y);
            if (accumulate) { const double2 o = out[p]; v.x += o.x; v.y += o.y; }
            out[p] = v;
        }
    }
}

// part[0][p] <- sum_c part[c][p] in a fixed order (deterministic); blockIdx.y selects the K / D spectra
// block (32, 8): 32 bins, the parts split over 8 lanes of the block (fixed association: part c goes to lane c % 8,
// lanes are added 0..7) - 8 x more CTAs and 8 x shorter load chains than one thread per bin (it was 0.1 ms per call,
// a term of the sharded step that does not shrink with the number of ranks)
struct PartialSpectra { double2* part[4]; };           // blockIdx.y picks one (K, D of one or two axes)
__global__ void __launch_bounds__(256)
reduce_partial_spectra_kernel(PartialSpectra P, int nparts, int L) {
    __shared__ double2 red[8][33];
    double2* __restrict__ part = P.part[blockIdx.y];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int p = blockIdx.x * 32 + tx;
    double2 s = make_double2(0.0, 0.0);
    if (p < L) {
#pragma unroll 4
        for (int c = ty; c < nparts; c += 8) { const double2 v = part[(size_t)c * L + p]; s.x += v.x; s.y += v.y; }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && p < L) {
#pragma unroll
        for (int j = 1; j < 8; ++j) { s.x += red[j][tx].x; s.y += red[j][tx].y; }
        part[p] = s;
    }
}

// blockIdx.x & 1 = 0: K spectrum -> sK (symmetric sums); 1: D spectrum -> sD (symmetric or antisymmetric); blockIdx.x >> 1 = axis
struct DiagSumsArgs {
    const double2* partK[2]; const double2* partD[2]; const double2* W[2]; int n[2]; int antisym[2]; double dirsign[2];
    const double* addK[2]; double addK_scale[2]; double* sK[2]; double* sD[2];
};
__global__ void __launch_bounds__(FFT_THREADS, 1)
spectrum_to_diag_sums_kernel(DiagSumsArgs A, int nparts, int L, int logL) {
    const int ax = blockIdx.x >> 1, which = blockIdx.x & 1;
    const double2* __restrict__ partK = A.partK[ax];
    const double2* __restrict__ partD = A.partD[ax];
    const double2* __restrict__ W = A.W[ax];
    const int n = A.n[ax], antisym = A.antisym[ax];
    const double dirsign = A.dirsign[ax], addK_scale = A.addK_scale[ax];
    const double* __restrict__ addK = A.addK[ax];
    double* __restrict__ sK = A.sK[ax];
    double* __restrict__ sD = A.sD[ax];
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles(xs, L, logL, W, tid);
    const double2* part = which == 0 ? partK : partD;
    for (int p = tid; p < L; p += FFT_THREADS) {
        double2 s = make_double2(0.0, 0.0);
        for (int c = 0; c < nparts; ++c) { const double2 v = part[(size_t)c * L + p]; s.x += v.x; s.y += v.y; }
        xs[PADI(p)] = s;
    }
    __syncthreads();
    fft_dit_inverse_inplace(xs, L, logL, W, tid);
    const double inv = 1.0 / (double)L;
    double* out = which == 0 ? sK : sD;
    const bool anti = (which == 1) && antisym;
    for (int m = tid; m < n; m += FFT_THREADS) {
        const double up = xs[PADI(m)].x * inv;                         // col - row = m
        const double lo = xs[PADI((L - m) & (L - 1))].x * inv;         // row - col = m
        double v;
        if (m == 0) v = anti ? 0.0 : up;
        else v = anti ? dirsign * (lo - up) : (up + lo);
        if (which == 0 && addK) v += addK_scale * addK[m];           // directly summed K^-1 diagonals
        out[m] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Toeplitz (derivative-Gram) products on uniform grids:  Out = alpha * X T^T (+ beta * Out), i.e.
// every row x of X is replaced by T x, where T[i][j] = t(i-j) is the n x n derivative Gram
// (symmetric k''(|i-j|h), or antisymmetric k'(|i-j|h) sgn(i-j) for advection).  T is embedded in a
// circulant of size L >= 2n; its spectrum (bit-reversed order, scaled by 1/L) is built once per
// step from the n-entry Toeplitz table.  Two real rows share one complex FFT exactly
// (c is real, so conv(c, x_r + i x_s) = conv(c, x_r) + i conv(c, x_s)): forward DIF, pointwise
// product, inverse DIT, all in place in shared memory.  Replaces the four D-GEMMs of the step
// (jnp.matmul at model_GP_solver_2d.py:112,119 and their transposes in the reverse pass).
// ---------------------------------------------------------------------------------------------
// spec[p] = FFT(c)[brev(p)] / L  with  c[m] = t(m), c[L-m] = t(-m)
// one CTA per job (the derivative-Gram and the Gram table of one or two axes share a launch)
struct SpectrumArgs { const double* tab[4]; const double2* W[4]; double2* spec[4]; int n[4]; int antisym[4]; double dirsign[4]; double diag_add[4]; };
__global__ void __launch_bounds__(FFT_THREADS, 1)
toeplitz_spectrum_kernel(SpectrumArgs A, int L, int logL, const int* __restrict__ skip) {
    if (skip && *skip) return;
    const int b = blockIdx.x;
    const double* __restrict__ tab = A.tab[b];
    const double2* __restrict__ W = A.W[b];
    double2* __restrict__ spec = A.spec[b];
    const int n = A.n[b], antisym = A.antisym[b];
    const double dirsign = A.dirsign[b], diag_add = A.diag_add[b];
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles(xs, L, logL, W, tid);
    for (int j = tid; j < L; j += FFT_THREADS) {
        double v = 0.0;
        if (j < n) v = antisym ? dirsign * tab[j] : tab[j];                       // i - j = m >= 0
        else if (L - j < n) v = antisym ? -dirsign * tab[L - j] : tab[L - j];     // i - j = -(L - j)
        if (antisym && j == 0) v = 0.0;
        if (j == 0) v += diag_add;                                                 // jitter of K = k(d) + jitter I
        xs[PADI(j)] = make_double2(v, 0.0);
    }
    __syncthreads();
    fft_dif_inplace(xs, L, logL, W, tid);
    const double inv = 1.0 / (double)L;
    for (int p = tid; p < L; p += FFT_THREADS) { const double2 v = xs[PADI(p)]; spec[p] = make_double2(v.x * inv, v.y * inv); }
}

__global__ void __launch_bounds__(FFT_THREADS, 1)
toeplitz_apply_kernel(const double* __restrict__ X, int rows, int n, int ldx, const double2* __restrict__ spec, int L,
                      int logL, const double2* __restrict__ W, double alpha, double beta, double* __restrict__ Out,
                      int ldo) {
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles(xs, L, logL, W, tid);
    const int npairs = (rows + 1) / 2;
    for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
        const int r0 = 2 * pr, r1 = r0 + 1;
        const bool two = r1 < rows;
        const double* x0 = X + (size_t)r0 * ldx;
        const double* x1 = X + (size_t)(two ? r1 : r0) * ldx;
        for (int j = tid; j < L; j += FFT_THREADS)
            xs[PADI(j)] = (j < n) ? make_double2(x0[j], two ? x1[j] : 0.0) : make_double2(0.0, 0.0);
        __syncthreads();
        fft_dif_inplace(xs, L, logL, W, tid);
        for (int p = tid; p < L; p += FFT_THREADS) { const int q = PADI(p); xs[q] = cmul(xs[q], spec[p]); }
        __syncthreads();
        fft_dit_inverse_inplace(xs, L, logL, W, tid);
        double* o0 = Out + (size_t)r0 * ldo;
        double* o1 = Out + (size_t)r1 * ldo;
        for (int j = tid; j < n; j += FFT_THREADS) {
            const double2 v = xs[PADI(j)];
            o0[j] = alpha * v.x + (beta != 0.0 ? beta * o0[j] : 0.0);
            if (two) o1[j] = alpha * v.y + (beta != 0.0 ? beta * o1[j] : 0.0);
        }
        __syncthreads();
    }
}

// out[c][r] = in[r][c] (+ out[c][r] when ACC); leading dimensions ldi, ldo (row blocks of a larger field)
template <bool ACC>
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ in, int R, int C, size_t ldi, double* __restrict__ out, size_t ldo) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        if (r < R && c < C) tile[i][tx] = in[(size_t)r * ldi + c];
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (r < R && c < C) {
            double* o = out + (size_t)c * ldo + r;
            *o = ACC ? tile[tx][i] + *o : tile[tx][i];
        }
    }
}

// Exchange packing of the sharded step (dist.py r2ct / ct2r): the transpose of in (R x C), split into
// parts of wc consecutive output rows; part d goes to out + d * pstride (+ the caller's array offset):
//   out[(c / wc) * pstride + (c % wc) * R + r] = in[r][c]
__global__ void __launch_bounds__(256)
transpose_parts_kernel(const double* __restrict__ in, int R, int C, int wc, size_t pstride, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        if (r < R && c < C) tile[i][tx] = in[(size_t)r * C + c];
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (r < R && c < C) out[(size_t)(c / wc) * pstride + (size_t)(c % wc) * R + r] = tile[tx][i];
    }
}

// Exchange unpacking: recv[s][a][r][c] (P x k x rows x seg) -> out[a][r][s * seg + c]  (k x rows x P*seg).
// One CTA row per (s, a, r): contiguous runs of seg doubles on both sides.
__global__ void __launch_bounds__(256)
unpack_segments_kernel(const double* __restrict__ recv, int P, int k, int rows, int seg, double* __restrict__ out) {
    const int nrun = P * k * rows;
    for (int run = blockIdx.x; run < nrun; run += gridDim.x) {
        const int r = run % rows, t = run / rows;
        const int a = t % k, sidx = t / k;
        const double* src = recv + (size_t)run * seg;
        double* dst = out + ((size_t)a * rows + r) * ((size_t)P * seg) + (size_t)sidx * seg;
        if ((seg & 1) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            const double2* s2 = reinterpret_cast<const double2*>(src);
            double2* d2 = reinterpret_cast<double2*>(dst);
            const int n2 = seg >> 1;
            for (int c0 = threadIdx.x; c0 < n2; c0 += 4 * blockDim.x) {          // four loads in flight per thread
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int c = c0 + u * blockDim.x; if (c < n2) v[u] = __ldcs(s2 + c); }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int c = c0 + u * blockDim.x; if (c < n2) d2[c] = v[u]; }
            }
        } else {
            for (int c = threadIdx.x; c < seg; c += blockDim.x) dst[c] = src[c];
        }
    }
}

static int ilog2(int L) { int l = 0; while ((1 << l) < L) ++l; return l; }

int fft_init() {
    static DeviceOnce once;
    if (!once.needed()) return GPHM_OK;
    const int bytes = (int)fft_smem_bytes(FFT_MAX_L);
    GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(xcorr_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(spectrum_to_diag_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(toeplitz_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(toeplitz_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    once.done();
    return GPHM_OK;
}

int launch_twiddle_init(double* W, int L, cudaStream_t st) {
    { LaunchScope scope(CAT_FFT, st); twiddle_init_kernel<<<(L + 255) / 256, 256, 0, st>>>(reinterpret_cast<double2*>(W), L, ilog2(L)); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_xcorr_spectrum(const double* X, const double* Y, int rows, int cols, int ldx, int ldy, int L, const double* W,
                          double weight, bool accumulate, double* partial, cudaStream_t st) {
    GPHM_TRY(fft_init());
    if (L > FFT_MAX_L || L < 2 * cols) { set_last_error("xcorr: L=%d does not fit cols=%d", L, cols); return GPHM_EINVAL; }
    {
        LaunchScope scope(CAT_FFT, st, 0.0, 16.0 * rows * (double)cols);
        xcorr_spectrum_kernel<<<fft_grid(), FFT_THREADS, fft_smem_bytes(L), st>>>(
            X, Y, rows, cols, ldx, ldy, L, ilog2(L), reinterpret_cast<const double2*>(W), weight, accumulate ? 1 : 0,
            reinterpret_cast<double2*>(partial));
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_spectrum_to_diag_sums_multi(const DiagSumsJob* jobs, int count, int L, cudaStream_t st) {
    GPHM_TRY(fft_init());
    if (count < 1 || count > 2) { set_last_error("spectrum_to_diag_sums: %d jobs", count); return GPHM_EINVAL; }
    PartialSpectra P = {};
    DiagSumsArgs A = {};
    for (int i = 0; i < count; ++i) {
        const DiagSumsJob& j = jobs[i];
        P.part[2 * i] = reinterpret_cast<double2*>(const_cast<double*>(j.partK));
        P.part[2 * i + 1] = reinterpret_cast<double2*>(const_cast<double*>(j.partD));
        A.partK[i] = reinterpret_cast<const double2*>(j.partK); A.partD[i] = reinterpret_cast<const double2*>(j.partD);
        A.W[i] = reinterpret_cast<const double2*>(j.W); A.n[i] = j.n; A.antisym[i] = j.antisym ? 1 : 0; A.dirsign[i] = j.dirsign;
        A.addK[i] = j.addK; A.addK_scale[i] = j.addK_scale; A.sK[i] = j.sK; A.sD[i] = j.sD;
    }
    {   // sum the per-CTA partial spectra with the whole GPU first (148 x L complex values each)
        LaunchScope scope(CAT_FFT, st, 0.0, 2.0 * count * 16.0 * fft_grid() * (double)L);
        reduce_partial_spectra_kernel<<<dim3((L + 31) / 32, 2 * count), dim3(32, 8), 0, st>>>(P, fft_grid(), L);
    }
    GPHM_LAUNCH_OK();
    {
        LaunchScope scope(CAT_FFT, st);
        spectrum_to_diag_sums_kernel<<<2 * count, FFT_THREADS, fft_smem_bytes(L), st>>>(A, 1, L, ilog2(L));
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_spectrum_to_diag_sums(const double* partK, const double* partD, int L, const double* W, int n, bool antisym,
                                 double dirsign, const double* addK, double addK_scale, double* sK, double* sD,
                                 cudaStream_t st) {
    const DiagSumsJob job = {partK, partD, W, n, antisym, dirsign, addK, addK_scale, sK, sD};
    return launch_spectrum_to_diag_sums_multi(&job, 1, L, st);
}

int launch_toeplitz_spectrum_multi(const ToeplitzSpectrumJob* jobs, int count, int L, cudaStream_t st, const int* skip) {
    GPHM_TRY(fft_init());
    if (count < 1 || count > 4) { set_last_error("toeplitz_spectrum: %d jobs", count); return GPHM_EINVAL; }
    SpectrumArgs A = {};
    for (int i = 0; i < count; ++i) {
        const ToeplitzSpectrumJob& j = jobs[i];
        A.tab[i] = j.tab; A.W[i] = reinterpret_cast<const double2*>(j.W); A.spec[i] = reinterpret_cast<double2*>(j.spec);
        A.n[i] = j.n; A.antisym[i] = j.antisym ? 1 : 0; A.dirsign[i] = j.dirsign; A.diag_add[i] = j.diag_add;
    }
    {
        LaunchScope scope(CAT_FFT, st);
        toeplitz_spectrum_kernel<<<count, FFT_THREADS, fft_smem_bytes(L), st>>>(A, L, ilog2(L), skip);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_toeplitz_spectrum(const double* tab, int n, int L, const double* W, bool antisym, double dirsign, double* spec,
                             cudaStream_t st, double diag_add, const int* skip) {
    const ToeplitzSpectrumJob job = {tab, n, W, antisym, dirsign, diag_add, spec};
    return launch_toeplitz_spectrum_multi(&job, 1, L, st, skip);
}

int launch_toeplitz_apply(const double* X, int rows, int n, int ldx, const double* spec, int L, const double* W, double alpha,
                          double beta, double* Out, int ldo, cudaStream_t st) {
    GPHM_TRY(fft_init());
    if (rows <= 0) return GPHM_OK;
    if (toeplitz_fused_supported(L) && L >= 2 * n)
        return launch_toeplitz_apply_fused(X, rows, n, ldx, spec, L, W, alpha, beta, nullptr, 0, Out, ldo, nullptr, st);
    {
        LaunchScope scope(CAT_FFT, st, 0.0, 16.0 * rows * (double)n);
        const int grid = std::min(fft_grid() * 1, (rows + 1) / 2);
        toeplitz_apply_kernel<<<grid, FFT_THREADS, fft_smem_bytes(L), st>>>(
            X, rows, n, ldx, reinterpret_cast<const double2*>(spec), L, ilog2(L), reinterpret_cast<const double2*>(W), alpha, beta,
            Out, ldo);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_transpose(const double* in, int R, int C, double* out, cudaStream_t st, int ldi, int ldo, bool accumulate) {
    if (R <= 0 || C <= 0) return GPHM_OK;
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    const size_t li = ldi > 0 ? ldi : C, lo = ldo > 0 ? ldo : R;
    {
        LaunchScope scope(CAT_ELEMWISE, st);
        if (accumulate) transpose_kernel<true><<<grid, 256, 0, st>>>(in, R, C, li, out, lo);
        else transpose_kernel<false><<<grid, 256, 0, st>>>(in, R, C, li, out, lo);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_transpose_parts(const double* in, int R, int C, int wc, size_t pstride, double* out, cudaStream_t st) {
    if (R <= 0 || C <= 0) return GPHM_OK;
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    { LaunchScope scope(CAT_ELEMWISE, st); transpose_parts_kernel<<<grid, 256, 0, st>>>(in, R, C, wc, pstride, out); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_unpack_segments(const double* recv, int P, int k, int rows, int seg, double* out, cudaStream_t st) {
    const size_t total = (size_t)P * k * rows * seg;
    if (total == 0) return GPHM_OK;
    const int grid = (int)std::min<size_t>((size_t)P * k * rows, (size_t)148 * 32);
    { LaunchScope scope(CAT_ELEMWISE, st); unpack_segments_kernel<<<grid, 256, 0, st>>>(recv, P, k, rows, seg, out); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
