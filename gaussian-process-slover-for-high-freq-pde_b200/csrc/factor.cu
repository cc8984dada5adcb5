// Blocked Cholesky factorisation K = L L^T and explicit triangular inverse L^-1 (FP64).
//
// Replaces the LU-based jnp.linalg.solve / jnp.linalg.slogdet of the reference
// (model_GP_solver_2d.py:104-105,158-161; model_GP_solver_1d.py:92,136): K is SPD
// (stationary kernel Gram + jitter*I), so Cholesky gives the same K^-1 applications and
// log|K| = 2*sum(log diag L) at a third of the LU cost.  Every K^-1 application in the step is
// then two triangular GEMMs with L^-1 (same 2N^3 FLOPs as two TRSMs, but at GEMM speed and
// with full tile parallelism); forward error is cond(L)*u per factor, like substitution.
//
//   chol_factor : right-looking, NB=128.  Per block column: (1) one-CTA factorisation of the
//                 diagonal block in shared memory (4x4 sub-blocks of 32x32, warp-level register
//                 Cholesky), which also inverts it and accumulates log-det; (2) panel L[i,b] = K[i,b] * inv(L_bb)^T (GEMM); (3) trailing update
//                 K[i,k] -= L[i,b] L[k,b]^T on lower tiles only (GEMM, KM_C_LOWER).
//   trtri_lower : level-by-level merge  inv([[L11,0],[L21,L22]]) = [[X11,0],[-X22 L21 X11, X22]]
//                 with all nodes of a level batched into two GEMM launches.
#include "common.cuh"
#include "kernels.h"

namespace gphm {

constexpr int DIAG_THREADS = 256;
constexpr int DIAG_WARPS = DIAG_THREADS / 32;
constexpr int SLD = kNB + 1;   // odd pitch: column walks hit distinct banks
constexpr int SB = 32;         // sub-block edge: one warp factors a 32x32 sub-block in registers
static_assert(kNB == 4 * SB, "the diagonal-block kernel is written for 4x4 sub-blocks");

// One CTA: factor the nb x nb block at Kbb (lower triangle read), write L_bb (upper zeroed) and
// inv(L_bb) (kNB x kNB, zero padded), and the block's log-det contribution.
//
// The block is padded to 128x128 with an identity and processed as 4x4 sub-blocks of 32x32:
//   for J = 0..3:  warp 0 factors sub-block (J,J) in registers (lane = row; the pivot column is
//                  exchanged through shared memory once per step);
//                  panel: every row below solves  p L_JJ^T = s  by substitution (thread = row);
//                  all warps: rank-32 update of the trailing sub-blocks.   -> 12 block barriers.
//   inverse: block forward substitution, block row I:  X[I][:] = L_II^-1 (E_I - L[I][<I] X[<I][:]),
//            the product by all warps, the triangular solve by substitution (thread = column).
// Only substitutions and products are used (no multiplication by inner inverses), so the result
// is as accurate as an unblocked factorisation - the parity tests are sensitive to this.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
chol_diag_kernel(const double* __restrict__ Kbb, double* __restrict__ Lbb, int ld, int nb,
                 double* __restrict__ invd, double* __restrict__ logdet_part, int* __restrict__ status,
                 int pivot_base, long long sK, long long sL, long long sInv, long long sLd, long long sStatus) {
    // blockIdx.x selects the system (the two axes are factored side by side when N1 == N2)
    Kbb += blockIdx.x * sK; Lbb += blockIdx.x * sL; invd += blockIdx.x * sInv;
    logdet_part += blockIdx.x * sLd; status += blockIdx.x * sStatus;
    extern __shared__ double sm[];
    double* S = sm;                          // kNB x SLD
    double* T = S + kNB * SLD;               // SB x SLD scratch (one block row of the inverse)
    double* rdiag = T + SB * SLD;            // 1 / L[i][i]
    double* ddiag = rdiag + kNB;             // L[i][i]
    double* colbuf = ddiag + kNB;            // pivot column exchange inside warp 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < kNB * kNB; idx += DIAG_THREADS) {
        const int i = idx >> 7, j = idx & (kNB - 1);
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j < nb) v = (j <= i) ? Kbb[(size_t)i * ld + j] : 0.0;
        S[i * SLD + j] = v;
    }
    __syncthreads();

    for (int J = 0; J < 4; ++J) {
        const int o = J * SB;
        if (warp == 0) {
            // ---- Cholesky of the 32x32 sub-block, lane = row ----
            double a[SB];
#pragma unroll
            for (int k = 0; k < SB; ++k) a[k] = S[(o + lane) * SLD + o + k];
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                double ajj = __shfl_sync(0xffffffffu, a[j], j);
                if (!(ajj > 0.0)) {                       // also catches NaN; uniform across the warp
                    if (lane == 0) atomicCAS(status, 0, pivot_base + o + j + 1);
                    ajj = 1.0;
                }
                const double d = sqrt(ajj);
                const double rd = 1.0 / d;
                double l = a[j] * rd;                     // meaningful for lane > j
                if (lane == j) { l = d; ddiag[o + j] = d; rdiag[o + j] = rd; }
                a[j] = l;
                colbuf[lane] = l;
                __syncwarp();
#pragma unroll
                for (int k = j + 1; k < SB; ++k) a[k] = fma(-l, colbuf[k], a[k]);   // rows < k carry unused values
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < SB; ++k) S[(o + lane) * SLD + o + k] = (k <= lane) ? a[k] : 0.0;
        }
        __syncthreads();
        // ---- panel: row r below the sub-block solves p L_JJ^T = s by substitution (thread = row) ----
        {
            const int r = o + SB + tid;
            if (r < kNB) {
                double v[SB];
#pragma unroll
                for (int c = 0; c < SB; ++c) v[c] = S[r * SLD + o + c];
#pragma unroll
                for (int k = 0; k < SB; ++k) {
                    const double pk = v[k] * rdiag[o + k];
                    v[k] = pk;
#pragma unroll
                    for (int c = k + 1; c < SB; ++c) v[c] = fma(-pk, S[(o + c) * SLD + o + k], v[c]);
                }
#pragma unroll
                for (int c = 0; c < SB; ++c) S[r * SLD + o + c] = v[c];
            }
        }
        __syncthreads();
        // ---- trailing update: S[i][k] -= sum_c P[i][c] P[k][c] for o+32 <= k <= i ----
        for (int i = o + SB + warp; i < kNB; i += DIAG_WARPS) {
            for (int kb = o + SB; kb <= i; kb += SB) {
                const int k = kb + lane;
                double s = 0.0;
#pragma unroll 8
                for (int c = 0; c < SB; ++c) s = fma(S[i * SLD + o + c], S[k * SLD + o + c], s);
                if (k <= i) S[i * SLD + k] -= s;
            }
        }
        __syncthreads();
    }

    // ---- write L_bb; log-det partial (fixed order) ----
    for (int idx = tid; idx < nb * nb; idx += DIAG_THREADS) {
        const int i = idx / nb, j = idx - i * nb;
        Lbb[(size_t)i * ld + j] = (j <= i) ? S[i * SLD + j] : 0.0;
    }
    if (warp == 0) {
        double s = 0.0;
        for (int i = lane; i < nb; i += 32) s += log(ddiag[i]);
        s = warp_sum(s);
        if (lane == 0) *logdet_part = s;
    }
    __syncthreads();

    // ---- inverse by block forward substitution, in place in S (row I of X overwrites row I of L) ----
    for (int I = 0; I < 4; ++I) {
        const int o = I * SB, ncol = o + SB;
        // T[i][c] = delta(o+i, c) - sum_{k=c}^{o-1} L[o+i][k] X[k][c]      (X[k][c] = 0 for k < c)
        for (int i = warp; i < SB; i += DIAG_WARPS)
            for (int c = lane; c < ncol; c += 32) {
                double s = (c == o + i) ? 1.0 : 0.0;
                for (int k = c; k < o; ++k) s = fma(-S[(o + i) * SLD + k], S[k * SLD + c], s);
                T[i * SLD + c] = s;
            }
        __syncthreads();
        // column c solves L_II x = T[:, c] by substitution (thread = column)
        if (tid < ncol) {
            const int c = tid;
            double v[SB];
#pragma unroll
            for (int i = 0; i < SB; ++i) v[i] = T[i * SLD + c];
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                const double xk = v[k] * rdiag[o + k];
                v[k] = xk;
#pragma unroll
                for (int i = k + 1; i < SB; ++i) v[i] = fma(-xk, S[(o + i) * SLD + o + k], v[i]);
            }
#pragma unroll
            for (int i = 0; i < SB; ++i) T[i * SLD + c] = v[i];
        }
        __syncthreads();
        for (int idx = tid; idx < SB * ncol; idx += DIAG_THREADS) {
            const int i = idx / ncol, c = idx - i * ncol;
            S[(o + i) * SLD + c] = T[i * SLD + c];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < kNB * kNB; idx += DIAG_THREADS) {
        const int i = idx >> 7, j = idx & (kNB - 1);
        invd[idx] = (i < nb && j < nb && j <= i) ? S[i * SLD + j] : 0.0;
    }
}

__global__ void copy_diag_blocks_kernel(const double* __restrict__ invd, double* __restrict__ Linv, int n, int ld) {
    const int b = blockIdx.x, j0 = b * kNB;
    const int nb = min(kNB, n - j0);
    for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
        const int i = idx / nb, j = idx - i * nb;
        Linv[(size_t)(j0 + i) * ld + j0 + j] = invd[(size_t)b * kNB * kNB + i * kNB + j];
    }
}

constexpr size_t kDiagSmem = (size_t)(kNB * SLD + SB * SLD + 2 * kNB + SB) * sizeof(double);

int factor_init() {
    static DeviceOnce once;
    if (!once.needed()) return GPHM_OK;
    GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDiagSmem));
    once.done();
    return GPHM_OK;
}

// nsys = 1: one system.  nsys = 2: a second system of the same size at the given element offsets
// (K2 - K, L2 - L, ...) is factored side by side: every launch carries both, which halves the
// serial latency chain of the 2-D step when N1 == N2.
int chol_factor_multi(double* K, double* L, int n, int ld, double* invdiag, double* logdet_part, int* status, int nsys,
                      long long sK, long long sL, long long sInv, long long sLd, long long sStatus, cudaStream_t st) {
    GPHM_TRY(factor_init());
    const int nblk = num_blocks_nb(n);
    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * kNB, nb = std::min(kNB, n - j0);
        { LaunchScope scope(CAT_CHOL_DIAG, st);
        chol_diag_kernel<<<nsys, DIAG_THREADS, kDiagSmem, st>>>(K + (size_t)j0 * ld + j0, L + (size_t)j0 * ld + j0, ld, nb,
                                                             invdiag + (size_t)b * kNB * kNB, logdet_part + b, status, j0,
                                                             sK, sL, sInv, sLd, sStatus); }
        GPHM_LAUNCH_OK();
        const int rem = n - j0 - nb;
        if (rem <= 0) break;
        const double* Kpan = K + (size_t)(j0 + nb) * ld + j0;
        double* Lpan = L + (size_t)(j0 + nb) * ld + j0;
        // panel: L[i,b] = K[i,b] * inv(L_bb)^T
        GemmArgs g1 = gemm_args(Kpan, ld, false, invdiag + (size_t)b * kNB * kNB, kNB, true, Lpan, ld, rem, nb, nb, 1.0, 0.0, 0);
        g1.batch = nsys; g1.sA = sK; g1.sB = sInv; g1.sC = sL;
        GPHM_TRY(launch_dgemm(g1, st));
        // trailing update (lower tiles): K22 -= L[:,b] L[:,b]^T
        GemmArgs g2 = gemm_args(Lpan, ld, false, Lpan, ld, true, K + (size_t)(j0 + nb) * ld + j0 + nb, ld, rem, rem, nb, -1.0, 1.0,
                                KM_C_LOWER);
        g2.batch = nsys; g2.sA = sL; g2.sB = sL; g2.sC = sK;
        GPHM_TRY(launch_dgemm(g2, st));
    }
    return GPHM_OK;
}

int chol_factor(double* K, double* L, int n, int ld, double* invdiag, double* logdet_part, int* status,
                cudaStream_t st) {
    return chol_factor_multi(K, L, n, ld, invdiag, logdet_part, status, 1, 0, 0, 0, 0, 0, st);
}

int trtri_lower(const double* L, double* Linv, int n, int ld, const double* invdiag, double* T, cudaStream_t st) {
    const int nblk = num_blocks_nb(n);
    { LaunchScope scope(CAT_ELEMWISE, st);
    copy_diag_blocks_kernel<<<nblk, 256, 0, st>>>(invdiag, Linv, n, ld); }
    GPHM_LAUNCH_OK();
    for (long long b = kNB; b < n; b *= 2) {
        const long long node = 2 * b;
        const int nfull = (int)(n / node);
        const long long stride = node * ld + node;
        if (nfull > 0) {
            // T21 = L21 * X11   (X11 lower: k >= column)
            GemmArgs g1 = gemm_args(L + b * ld, ld, false, Linv, ld, false, T + b * ld, ld, (int)b, (int)b, (int)b,
                                    1.0, 0.0, KM_B_LOWER);
            g1.sA = g1.sB = g1.sC = stride; g1.batch = nfull;
            GPHM_TRY(launch_dgemm(g1, st));
            // X21 = -X22 * T21  (X22 lower: k <= row)
            GemmArgs g2 = gemm_args(Linv + b * ld + b, ld, false, T + b * ld, ld, false, Linv + b * ld, ld,
                                    (int)b, (int)b, (int)b, -1.0, 0.0, KM_A_LOWER);
            g2.sA = g2.sB = g2.sC = stride; g2.batch = nfull;
            GPHM_TRY(launch_dgemm(g2, st));
        }
        const long long o = (long long)nfull * node;
        if (o + b < n) {                                  // ragged last node: right child has r < b rows
            const int r = (int)(n - o - b);
            const double* L21 = L + (o + b) * ld + o;
            double* T21 = T + (o + b) * ld + o;
            GPHM_TRY(launch_dgemm(gemm_args(L21, ld, false, Linv + o * ld + o, ld, false, T21, ld, r, (int)b, (int)b,
                                            1.0, 0.0, KM_B_LOWER), st));
            GPHM_TRY(launch_dgemm(gemm_args(Linv + (o + b) * ld + o + b, ld, false, T21, ld, false,
                                            Linv + (o + b) * ld + o, ld, r, (int)b, r, -1.0, 0.0, KM_A_LOWER), st));
        }
    }
    return GPHM_OK;
}

}  // namespace gphm
