// Internal launcher declarations shared by the translation units of libgphm.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace gphm {

// ---- gram.cu -------------------------------------------------------------------------------
int launch_gram_general(int kid, int order, const double* x1, int n1, const double* x2, int n2,
                        const double* theta, int Q, double jitter, double* Kout, double* Dout, int ld,
                        cudaStream_t st);
int launch_gram_toeplitz(int kid, int order, const double* x, int n, const double* theta, int Q, double jitter,
                         double dirsign, double* tabK, double* tabD, double* Kout, double* Dout, int ld,
                         cudaStream_t st);

// skip (optional, every launcher of the uniform-grid factor stage): device flag; non-zero turns the launch into a no-op
int launch_toeplitz_table(int kid, int order, const double* x, int n, const double* theta, int Q, double* tabK, double* tabD,
                          cudaStream_t st, const int* skip = nullptr);
int launch_kappa_pairs(int kid, int order, const double* x1, const double* x2, size_t np, const double* theta, int Q,
                       double* out, cudaStream_t st);

// ---- dgemm.cu ------------------------------------------------------------------------------
// k-range / output structure flags.  "op(A) lower" means op(A)[m][k] == 0 for k > m, etc.
enum : int { KM_A_LOWER = 1, KM_A_UPPER = 2, KM_B_LOWER = 4, KM_B_UPPER = 8, KM_C_LOWER = 16 };

struct GemmArgs {
    const double* A; const double* B; double* C;   // row-major, leading dimensions lda/ldb/ldc
    int M, N, K;                                   // C[M,N] = alpha*op(A)[M,K]*op(B)[K,N] + beta*C
    int lda, ldb, ldc;
    double alpha, beta;
    int transA, transB;                            // op(X) = X^T when set (X stored K x M / N x K)
    int kmode;
    long long sA, sB, sC;                          // batch strides (elements)
    int batch;
};
inline GemmArgs gemm_args(const double* A, int lda, bool tA, const double* B, int ldb, bool tB, double* C, int ldc,
                          int M, int N, int K, double alpha = 1.0, double beta = 0.0, int kmode = 0) {
    GemmArgs g{A, B, C, M, N, K, lda, ldb, ldc, alpha, beta, tA ? 1 : 0, tB ? 1 : 0, kmode, 0, 0, 0, 1};
    return g;
}
int launch_dgemm(const GemmArgs& g, cudaStream_t st);
int dgemm_init();   // opt in to large dynamic shared memory once per process

// ---- factor.cu -----------------------------------------------------------------------------
constexpr int kNB = 128;   // diagonal block size of the blocked Cholesky / triangular inverse
inline int num_blocks_nb(int n) { return (n + kNB - 1) / kNB; }
// K (n x n, ld, lower triangle read, destroyed) -> L (lower, separate buffer), inverse diagonal
// blocks invdiag[nblk][kNB][kNB], logdet_part[nblk] = sum log diag(L_bb); status[0] = 1 + index
// of the first non-positive pivot (0 if SPD).
int chol_factor(double* K, double* L, int n, int ld, double* invdiag, double* logdet_part, int* status,
                cudaStream_t st);
int chol_factor_multi(double* K, double* L, int n, int ld, double* invdiag, double* logdet_part, int* status, int nsys,
                      long long sK, long long sL, long long sInv, long long sLd, long long sStatus, cudaStream_t st);
// Linv (n x n, ld; strictly-upper triangle must already be zero) = L^-1, T = n x n scratch.
int trtri_lower(const double* L, double* Linv, int n, int ld, const double* invdiag, double* T, cudaStream_t st);
int factor_init();

// ---- fft.cu --------------------------------------------------------------------------------
int fft_length_for(int n);          // power of two >= 2n, or 0 when it does not fit shared memory
int fft_grid();                     // CTAs (= partial spectra) used by launch_xcorr_spectrum
int launch_twiddle_init(double* W, int L, cudaStream_t st);
// partial[cta][L] (complex) (+)= weight * sum_rows conj(FFT(X row)) * FFT(Y row)
int launch_xcorr_spectrum(const double* X, const double* Y, int rows, int cols, int ldx, int ldy, int L, const double* W,
                          double weight, bool accumulate, double* partial, cudaStream_t st);
int launch_spectrum_to_diag_sums(const double* partK, const double* partD, int L, const double* W, int n, bool antisym,
                                 double dirsign, const double* addK, double addK_scale, double* sK, double* sD,
                                 cudaStream_t st);
// spec (L complex, bit-reversed order, scaled by 1/L) of the circulant embedding of the Toeplitz matrix t(i-j) = tab[|i-j|] (* sign)
// one launch for up to four Toeplitz tables of the same transform length (one CTA each)
struct ToeplitzSpectrumJob { const double* tab; int n; const double* W; bool antisym; double dirsign; double diag_add; double* spec; };
int launch_toeplitz_spectrum_multi(const ToeplitzSpectrumJob* jobs, int count, int L, cudaStream_t st, const int* skip = nullptr);
// the diagonal sums of one or two axes of the same transform length in one pair of launches
struct DiagSumsJob { const double* partK; const double* partD; const double* W; int n; bool antisym; double dirsign;
                     const double* addK; double addK_scale; double* sK; double* sD; };
int launch_spectrum_to_diag_sums_multi(const DiagSumsJob* jobs, int count, int L, cudaStream_t st);
int launch_toeplitz_spectrum(const double* tab, int n, int L, const double* W, bool antisym, double dirsign, double* spec,
                             cudaStream_t st, double diag_add = 0.0, const int* skip = nullptr);
// Out[r][:] = alpha * T X[r][:] + beta * Out[r][:]   for every row r (T n x n Toeplitz with spectrum `spec`)
int launch_toeplitz_apply(const double* X, int rows, int n, int ldx, const double* spec, int L, const double* W, double alpha,
                          double beta, double* Out, int ldo, cudaStream_t st);
// out (C x R, leading dimension ldo, default R) = in^T (in: R x C, leading dimension ldi, default C), + out when accumulate
int launch_transpose(const double* in, int R, int C, double* out, cudaStream_t st, int ldi = 0, int ldo = 0, bool accumulate = false);
int launch_transpose_parts(const double* in, int R, int C, int wc, size_t pstride, double* out, cudaStream_t st);
int launch_unpack_segments(const double* recv, int P, int k, int rows, int seg, double* out, cudaStream_t st);

// ---- toeplitz_inv.cu -----------------------------------------------------------------------
int toeplitz_inv_max_n();
// nsys SPD Toeplitz systems (first columns tabK + s*sTab, jitter added to entry 0):  g = K^-1 e_0,
// half_logdet[0] = log|K| / 2, status[0] = 1 + first step with a non-positive prediction error.
// gkap[n] doubles + prog[1] int per system: hand-over buffer between the generator and the lattice CTA.
int launch_schur_levinson(const double* tabK, long long sTab, int n, double jitter, double* g, long long sG,
                          double* half_logdet, long long sLd, int* status, long long sStatus, double* gkap, long long sKap,
                          int* prog, long long sProg, int nsys, cudaStream_t st, long long* dbg_cycles = nullptr,
                          int* guard = nullptr, int guard_bit0 = 0, double* gbnd = nullptr, long long sBnd = 0,
                          const int* skip = nullptr);
// gbnd (optional): 6 n doubles per system - the hand-over buffers of the multi-CTA recursion (prog then needs 8 ints per system)
int schur_split_factor(int n);
// guard: bit (guard_bit0 + s) is OR-ed in when system s has min_k (1 - kappa_k^2) < toeplitz_guard_min()
double toeplitz_guard_min();
// spec[4][L] complex (strides in doubles): Gohberg-Semencul circulant spectra; sKinv[n]: diagonal sums of K^-1
int launch_gs_prepare(const double* g, long long sG, int n, int L, const double* W, double* spec, long long sSpec,
                      double* sKinv, long long sS, int nsys, cudaStream_t st, const int* skip = nullptr);

// ---- toeplitz_fused.cu ---------------------------------------------------------------------
bool toeplitz_fused_supported(int L);   // L >= 16: fused-sweep kernels; smaller sizes use the plain ones in fft.cu
// SpecOut (optional): ceil(rows/2) x L complex, the transform of every packed row pair of X (for launch_xcorr_pairs)
int launch_toeplitz_apply_fused(const double* X, int rows, int n, int ldx, const double* spec, int L, const double* W,
                                double alpha, double beta, const double* Add, int lda, double* Out, int ldo, double* SpecOut,
                                cudaStream_t st);
int launch_xcorr_pairs(const double* X, int rows, int n, int ldx, const double* SpecY, int L, const double* W, double weight,
                       double* partial, cudaStream_t st);
// Out[r] = alpha * K^-1 X[r] + beta * Add[r] for every row, K^-1 through the four spectra of launch_gs_prepare
// g0ptr (optional): device pointer to g[0] = (K^-1)_00; with it and L == 2n only the first spectrum is read (the others derive from it)
int launch_gs_apply_fused(const double* X, int rows, int n, int ldx, const double* gspec, int L, const double* W, double alpha,
                          double beta, const double* Add, int lda, double* Out, int ldo, cudaStream_t st,
                          const double* g0ptr = nullptr);

// ---- ozaki.cu: FP64-accurate GEMM on tcgen05 (int8 Ozaki slices, TMA operands, TMEM accumulators) ----
int ozaki_default_slices();                                  // GPHM_OZAKI_SLICES, default 8
size_t ozaki_work_bytes(int M, int N, int K, int S);
double ozaki_error_factor(int K, int S);                     // |dC_ij| <= |alpha| * factor * max_k|A_ik| * max_k|B_kj|
int launch_ozaki_dgemm(bool transA, bool transB, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
                       int ldb, double beta, double* C, int ldc, int S, void* work, size_t work_bytes, cudaStream_t st);

// ---- peer.cu: transposing all-to-all through NVLink peer stores ----
size_t peer_flag_bytes();
int launch_a2a_transpose_peer(const double* const* in, int k, int rows, int cols, int pc, double* const* peer_bases, int P, int me,
                              unsigned long long seq, unsigned int* done, cudaStream_t st);
int launch_mg_wait_flags(const void* local_base, int P, unsigned long long seq, int* status, cudaStream_t st);

// ---- elemwise.cu ---------------------------------------------------------------------------
struct LossConsts { int dim, eq_type, n1, n2, nb, Q; double llk_weight, logdet, c1; };
constexpr int kRedBlocks = 592;         // 148 SMs x 4
int launch_residual(double* R, const double* U, const double* F, const double* A, const double* Bt, size_t n,
                    int eq_type, const double* base, const double* small, int Q, double* part, cudaStream_t st);
int launch_finalize(const LossConsts& c, const double* U, const double* bvals, const int* xind,
                    const double* part, const double* ldp1, int nblk1, const double* ldp2, int nblk2,
                    const double* small, double* eb, double* terms, double* gsmall, int* status, cudaStream_t st);
int launch_mg_finalize(const LossConsts& c, const double* sums, const double* ld, const double* small, double* terms,
                       double* gsmall, int* status, cudaStream_t st);
int launch_grad_u(const LossConsts& c, const double* base, const double* U, const double* G, const double* W, const double* S1,
                  const double* S2, const double* eb, const int* xind, const double* small, double* gU,
                  double* V1, double* V2, cudaStream_t st);
// Dbar may be NULL (then only sK is produced)
int launch_diag_sums(const double* Kbar, const double* Dbar, int n, int ld, bool antisym, double dirsign,
                     double* part, double* sK, double* sD, cudaStream_t st);
size_t diag_sums_part_doubles(int n);
struct ThetaGradJob { const double* x; int n; const double* theta; const double* sK; const double* sD; double* gtheta; };
int launch_theta_grad_toeplitz_multi(int kid, int order, const ThetaGradJob* jobs, int count, int Q, cudaStream_t st);   // one launch for both axes
int launch_theta_grad_toeplitz(int kid, int order, const double* x, int n, const double* theta, int Q,
                               const double* sK, const double* sD, double* gtheta, cudaStream_t st);
int launch_theta_grad_general(int kid, int order, const double* x, int n, const double* theta, int Q,
                              const double* Kbar, const double* Dbar, int ld, double* part, double* gtheta,
                              cudaStream_t st);
size_t theta_general_part_doubles(int n, int Q);
int launch_adam(double* p, const double* g, double* m, double* v, size_t n, const long long* count, double lr,
                cudaStream_t st);
int launch_count_inc(long long* count, cudaStream_t st);
int launch_adam_inc(double* p, const double* g, double* m, double* v, size_t n, long long* count, double lr, cudaStream_t st);
// look-ahead bookkeeping (plan.cu): out-of-place Adam on the short vector; theta compare / flag kernels
int launch_adam_out(const double* p, double* p_out, const double* g, double* m, double* v, size_t n, const long long* count,
                    double lr, cudaStream_t st);
int launch_lk_compare(const double* small, const double* lk_small, int n, int* flags, cudaStream_t st);   // flags[1] = flags[0] && equal; flags[0] = 0
int launch_lk_set(int* flags, int value, cudaStream_t st);                                                // flags[0] = value
int launch_rel_l2(const double* pred, const double* truth, size_t n, double* part, double* out, cudaStream_t st);
int launch_copy(double* dst, const double* src, size_t n, cudaStream_t st);
int launch_pair_reduce(const double* part, double* out2, cudaStream_t st);
int launch_boundary_indexed(const double* U, const int* bidx, const double* bvals, int nb, double* eb, double* out,
                            cudaStream_t st);
int launch_boundary_scatter_indexed(double* gU, const int* bidx, const double* eb, int nseg0, int nb, double llk_weight,
                                    const double* log_tau, cudaStream_t st);
int launch_lincomb(double* out, double a, const double* x, double b, const double* y, size_t n, cudaStream_t st);
int launch_grad_u_local(size_t n, int allencahn, const double* U, const double* G, const double* W, const double* S1,
                        const double* S2, double* gU, double* V1, double* V2, cudaStream_t st);
int launch_sum_scaled(const double* v, int n, double scale, double* out, cudaStream_t st);
int launch_status_merge_guard(int* status, const int* guard, cudaStream_t st);

}  // namespace gphm
