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
//                 diagonal block in shared memory, which also inverts it and accumulates
//                 log-det; (2) panel L[i,b] = K[i,b] * inv(L_bb)^T (GEMM); (3) trailing update
//                 K[i,k] -= L[i,b] L[k,b]^T on lower tiles only (GEMM, KM_C_LOWER).
//   trtri_lower : level-by-level merge  inv([[L11,0],[L21,L22]]) = [[X11,0],[-X22 L21 X11, X22]]
//                 with all nodes of a level batched into two GEMM launches.
#include "common.cuh"
#include "kernels.h"

namespace gphm {

constexpr int DIAG_THREADS = 1024;
constexpr int SLD = kNB + 1;   // odd pitch: column walks hit distinct banks

// One CTA: factor the nb x nb block at Kbb (lower triangle read), write L_bb (upper zeroed) and
// inv(L_bb) (kNB x kNB, zero padded), and the block's log-det contribution.
__global__ void __launch_bounds__(DIAG_THREADS, 1)
chol_diag_kernel(const double* __restrict__ Kbb, double* __restrict__ Lbb, int ld, int nb,
                 double* __restrict__ invd, double* __restrict__ logdet_part, int* __restrict__ status,
                 int pivot_base) {
    extern __shared__ double sm[];
    double* S = sm;                       // nb x nb block, pitch SLD
    double* col = sm + kNB * SLD;         // scaled pivot column
    double* dg = col + kNB;               // diagonal of L
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;

    for (int idx = tid; idx < nb * nb; idx += DIAG_THREADS) {
        const int i = idx / nb, j = idx - i * nb;
        S[i * SLD + j] = Kbb[(size_t)i * ld + j];
    }
    __syncthreads();

    // ---- Cholesky, right-looking rank-1 updates (2 barriers per column) ----
    for (int j = 0; j < nb; ++j) {
        double ajj = S[j * SLD + j];
        if (!(ajj > 0.0)) {               // also catches NaN
            if (tid == 0) atomicCAS(status, 0, pivot_base + j + 1);
            ajj = 1.0;
        }
        const double d = sqrt(ajj);
        const double r = 1.0 / d;
        for (int i = j + 1 + tid; i < nb; i += DIAG_THREADS) {
            const double v = S[i * SLD + j] * r;
            col[i] = v;
            S[i * SLD + j] = v;
        }
        if (tid == 0) dg[j] = d;
        __syncthreads();
        for (int i = j + 1 + ty; i < nb; i += 32) {
            const double ci = col[i];
            for (int k = j + 1 + tx; k <= i; k += 32) S[i * SLD + k] -= ci * col[k];
        }
        __syncthreads();
    }

    // ---- write L_bb; log-det partial (fixed order) ----
    for (int idx = tid; idx < nb * nb; idx += DIAG_THREADS) {
        const int i = idx / nb, j = idx - i * nb;
        Lbb[(size_t)i * ld + j] = (j < i) ? S[i * SLD + j] : (j == i ? dg[i] : 0.0);
    }
    if (tid < 32) {
        double s = 0.0;
        for (int i = tid; i < nb; i += 32) s += log(dg[i]);
        s = warp_sum(s);
        if (tid == 0) *logdet_part = s;
    }

    // ---- in-place inverse of the lower-triangular block (columns from last to first):
    //      X[j][j] = 1/L[j][j];  X[i][j] = -(sum_{k=j+1..i} X[i][k] L[k][j]) * X[j][j]
    //      8 lanes cooperate on one row. ----
    const int sub = tid & 7, rgrp = tid >> 3;           // 128 row groups
    const unsigned gmask = 0xffu << (tid & 24);         // the 8 lanes of this row group (same trip count)
    for (int j = nb - 1; j >= 0; --j) {
        for (int i = j + 1 + tid; i < nb; i += DIAG_THREADS) col[i] = S[i * SLD + j];
        __syncthreads();
        const double xjj = 1.0 / dg[j];
        for (int i = j + 1 + rgrp; i < nb; i += DIAG_THREADS / 8) {
            double s = 0.0;
            for (int k = j + 1 + sub; k <= i; k += 8) {
                const double xik = (k == i) ? 1.0 / dg[i] : S[i * SLD + k];
                s += xik * col[k];
            }
            s += __shfl_xor_sync(gmask, s, 1);
            s += __shfl_xor_sync(gmask, s, 2);
            s += __shfl_xor_sync(gmask, s, 4);
            if (sub == 0) S[i * SLD + j] = -s * xjj;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < kNB * kNB; idx += DIAG_THREADS) {
        const int i = idx / kNB, j = idx - i * kNB;
        double v = 0.0;
        if (i < nb && j < nb) v = (j < i) ? S[i * SLD + j] : (j == i ? 1.0 / dg[i] : 0.0);
        invd[idx] = v;
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

constexpr size_t kDiagSmem = (size_t)(kNB * SLD + 2 * kNB) * sizeof(double);

int factor_init() {
    static int done = -1;
    if (done >= 0) return done;
    GPHM_CUDA_OK(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDiagSmem));
    done = GPHM_OK;
    return done;
}

int chol_factor(double* K, double* L, int n, int ld, double* invdiag, double* logdet_part, int* status,
                cudaStream_t st) {
    GPHM_TRY(factor_init());
    const int nblk = num_blocks_nb(n);
    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * kNB, nb = std::min(kNB, n - j0);
        { LaunchScope scope(CAT_CHOL_DIAG, st);
        chol_diag_kernel<<<1, DIAG_THREADS, kDiagSmem, st>>>(K + (size_t)j0 * ld + j0, L + (size_t)j0 * ld + j0, ld, nb,
                                                             invdiag + (size_t)b * kNB * kNB, logdet_part + b, status, j0); }
        GPHM_LAUNCH_OK();
        const int rem = n - j0 - nb;
        if (rem <= 0) break;
        const double* Kpan = K + (size_t)(j0 + nb) * ld + j0;
        double* Lpan = L + (size_t)(j0 + nb) * ld + j0;
        // panel: L[i,b] = K[i,b] * inv(L_bb)^T
        GPHM_TRY(launch_dgemm(gemm_args(Kpan, ld, false, invdiag + (size_t)b * kNB * kNB, kNB, true, Lpan, ld,
                                        rem, nb, nb, 1.0, 0.0, 0), st));
        // trailing update (lower tiles): K22 -= L[:,b] L[:,b]^T
        GPHM_TRY(launch_dgemm(gemm_args(Lpan, ld, false, Lpan, ld, true, K + (size_t)(j0 + nb) * ld + j0 + nb, ld,
                                        rem, rem, nb, -1.0, 1.0, KM_C_LOWER), st));
    }
    return GPHM_OK;
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
