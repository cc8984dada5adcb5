/* libgphm - B200 (sm_100a) implementation of the GP-HM solver's inner loop.
 *
 * Plain C ABI: no torch / jax types in any signature.  Every pointer named `d_*` is a DEVICE
 * pointer owned by the caller (torch CUDA tensors in the Python host layer); `h_*` is a HOST
 * pointer.  All matrices are row-major FP64.  Kernels launch on the caller's `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream); nothing synchronises the host
 * except gphm_plan_create, gphm_plan_status and the *_host entry points.
 *
 * The reference (xuangu-fang/Gaussian-Process-Slover-for-High-Freq-PDE) has no FFI: its seam is
 * Python methods.  Each entry point below names the reference method it replaces (paths are
 * relative to the reference's code/ directory).
 *
 * Return value: 0 ok; <0 usage/runtime error (gphm_last_error() has the text);
 * numerical conditions (non-SPD Gram) are reported asynchronously through gphm_plan_status.
 */
#ifndef GPHM_H_
#define GPHM_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPHM_VERSION 100

#if defined(__GNUC__)
#define GPHM_API __attribute__((visibility("default")))
#else
#define GPHM_API
#endif

/* kernel ids = the four kernel classes of kernel_matrix.py */
#define GPHM_KERNEL_SE_COS        0   /* SE_Cos_1d        kernel_matrix.py:107-128 */
#define GPHM_KERNEL_MATERN52_COS  1   /* Matern52_Cos_1d  kernel_matrix.py:131-155 */
#define GPHM_KERNEL_MATERN52      2   /* Matern52_1d      kernel_matrix.py:158-176 */
#define GPHM_KERNEL_SE            3   /* SE_1d            kernel_matrix.py:179-193 */

/* equation types = eq_type branches of boundary_and_eq_gap */
#define GPHM_EQ_POISSON    0   /* model_GP_solver_2d.py:131-133, model_GP_solver_1d.py:107-110 */
#define GPHM_EQ_ALLENCAHN  1   /* model_GP_solver_2d.py:135-138, model_GP_solver_1d.py:112-115 */
#define GPHM_EQ_ADVECTION  2   /* model_GP_solver_advection.py:132-134 (first-derivative Grams) */

/* status codes */
#define GPHM_OK          0
#define GPHM_EINVAL     -1
#define GPHM_ECUDA      -2
#define GPHM_ENOMEM     -3
#define GPHM_NOT_SPD     1   /* gphm_plan_status: a Gram matrix had a non-positive pivot */
#define GPHM_NONFINITE   2   /* gphm_plan_status: the loss is not finite */
#define GPHM_ILL_CONDITIONED 3 /* gphm_plan_status: a uniform-grid axis is too ill-conditioned for the Toeplitz
                                 inverse-generator route (min_k (1 - kappa_k^2) < 3.5e-5): results are still finite, but
                                 call gphm_plan_use_cholesky to stay inside the 1e-6 parity bound */

#define GPHM_STALLED     4   /* gphm_plan_status: the two-CTA Schur/Levinson kernel timed out waiting for its producer
                                 (results are NaN); cannot happen while the cluster launch co-schedules both CTAs */

/* flags for gphm_logjoint_grad */
#define GPHM_FORWARD_ONLY  1   /* loss terms only (compute_early_stopping, loss) */

/* Problem constants the reference bakes into its jitted executable through the static `self`
 * (model_GP_solver_2d.py:40-85, model_GP_solver_1d.py:38-78, model_GP_solver_advection.py:40-85). */
typedef struct gphm_problem_desc {
    int dim;             /* 1: GP_solver_1d_single; 2: GP_solver_2d_single[_advection] */
    int kernel_id;       /* GPHM_KERNEL_* (trick_paras['kernel']) */
    int eq_type;         /* GPHM_EQ_* */
    int n1, n2;          /* collocation points per axis (n2 = 1 when dim == 1) */
    int Q;               /* mixture components (trick_paras['Q']) */
    int nb;              /* boundary points: 2*n1+2*n2 (2-D) or len(Xind) (1-D) */
    int force_general;   /* bit 0: never use the Toeplitz fast path even on uniform grids;
                            bit 1: Toeplitz Grams, but Kbar/Dbar by GEMM + direct diagonal sums (no FFT);
                            bit 2: FFT diagonal sums, but K^-1 by GEMM + direct sums;
                            bit 3: derivative-Gram products D*A, D^T*G by GEMM, not Toeplitz FFT;
                            bit 4: K^-1 by blocked Cholesky + triangular GEMMs even on uniform grids
                                   (default there: Schur/Levinson recursion + Gohberg-Semencul FFT
                                   products, no dense factorisation);
                            bit 5: no iterative-refinement step on the reverse-pass K^-1 applications of the
                                   Toeplitz inverse-generator route (measurement only: theta-gradients lose
                                   ~cond(K)/40 * 1e-13 of relative accuracy, 1.4e-6 at N = 4096);
                            bit 6: the plain contractions of the general path (D A, Bt D^T, D^T G, G D, V A^T, G A^T:
                                   jnp.matmul at model_GP_solver_2d.py:112,119 and their reverse pass) on the tensor
                                   cores - Ozaki int8 slices, tcgen05.mma kind::i8, TMA operands, TMEM accumulators
                                   (gphm_ozaki_dgemm, stated bound there); the solves stay native FP64;
                            bit 7: refinement step also on axes of <= 512 points (measurement only);
                            bit 8 / bit 9: gphm_step look-ahead of the next step's factor stage off / on (default off:
                                   bit-exact but no faster, see plan.cu) */
    double llk_weight;   /* trick_paras['llk_weight'] */
    double logdet;       /* trick_paras['logdet'] (True -> 1.0) */
    double beta;         /* advection speed, trick_paras['beta'] (ignored otherwise) */
    double jitter;       /* 1e-6 at every reference call site */
} gphm_problem_desc;

typedef struct gphm_plan gphm_plan;   /* opaque */

GPHM_API int gphm_version(void);
GPHM_API const char* gphm_last_error(void);

/* ---- accounting ----------------------------------------------------------------------------
 * gphm_launch_count: kernels launched by this library in this process so far.
 * gphm_profile_start/stop: while enabled every launch is bracketed by CUDA events on its stream;
 * stop synchronises the device and returns, per kernel family (index: 0 Gram builders, 1 DGEMM,
 * 2 serial factorisation chain (Cholesky diagonal blocks / Schur-Levinson recursion), 3 reductions
 * and element-wise, 4 Adam, 5 FFT diagonal sums and spectra, 6 Gohberg-Semencul K^-1 application,
 * 7 Toeplitz derivative-Gram products), the summed device time in ms, the FLOPs (issued for DGEMM;
 * 5 L log2 L per complex transform + the spectrum products for the FFT kernels), algorithmic bytes
 * and launch counts (arrays of GPHM_PROFILE_FAMILIES = 8).                                     */
#define GPHM_PROFILE_FAMILIES 8
GPHM_API long long gphm_launch_count(void);
GPHM_API int gphm_profile_start(void);
GPHM_API int gphm_profile_stop(double* ms, double* flops, double* bytes, long long* launches);

/* ---- Gram builders -------------------------------------------------------------------------
 * gphm_gram replaces Kernel_matrix.get_kernel_matrix (kernel_matrix.py:21-30; deriv_order 0,
 * jitter added on the diagonal when n1 == n2 and jitter != 0), vmap(D_x1_kappa)
 * (kernel_matrix.py:49-52; deriv_order 1), vmap(DD_x1_kappa) (:54-57; deriv_order 2) and the
 * rectangular cross-Grams of preds (model_GP_solver_2d.py:198-202; deriv_order 0, jitter 0).
 * d_theta = [log-w (Q) | log-ls (Q) | freq (Q)].  d_out is n1 x n2.                           */
GPHM_API int gphm_gram(int kernel_id, int deriv_order, const double* d_x1, int n1, const double* d_x2, int n2,
              const double* d_theta, int Q, double jitter, double* d_out, void* stream);

/* Element-wise ("vmapped") kappa / D_x1_kappa / DD_x1_kappa over explicit pair lists, exactly the
 * call shape of vmap(self.K_u.kappa, (0, 0, None))(X1.flatten(), X2.flatten(), paras)
 * (kernel_matrix.py:26-27): d_out[p] = d^order/dx1^order kappa(d_x1[p], d_x2[p]).              */
GPHM_API int gphm_kappa_pairs(int kernel_id, int deriv_order, const double* d_x1, const double* d_x2, size_t npairs,
                     const double* d_theta, int Q, double* d_out, void* stream);

/* ---- dense FP64 primitives (exported for parity tests and the prediction path) --------------
 * C[M,N] = alpha*op(A)*op(B) + beta*C, row-major (jnp.matmul, model_GP_solver_2d.py:112,119).  */
GPHM_API int gphm_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* d_A, int lda,
               const double* d_B, int ldb, double beta, double* d_C, int ldc, void* stream);
/* The same contraction on the 5th-generation tensor cores: both operands are cut into `slices` (2..8, 0 = default 8)
 * signed 7-bit digits per row / column (Ozaki splitting, exact in FP64), the digit products run as
 * tcgen05.mma kind::i8 with TMA-staged operands and exact int32 accumulators in TMEM, and the partial products are
 * recombined in FP64.  STATED BOUND (looser than native FP64, north_star "TF32-emulated GEMMs"):
 *     |C - C_exact|_ij <= |alpha| * gphm_ozaki_error_factor(K, slices) * max_k |op(A)_ik| * max_k |op(B)_kj|,
 * factor = 4 K (slices + 1.1) 2^(-7 slices): 5.2e-13 at K = 4096, slices = 8.  d_work: gphm_ozaki_work_bytes bytes.
 * Plans use it for jnp.matmul(K_dxx1, K1inv_U) (model_GP_solver_2d.py:112,119) and the matching reverse-mode products on
 * the general path when force_general bit 6 is set; the solves stay on native FP64.                            */
GPHM_API size_t gphm_ozaki_work_bytes(int M, int N, int K, int slices);
GPHM_API double gphm_ozaki_error_factor(int K, int slices);
GPHM_API int gphm_ozaki_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* d_A, int lda,
                     const double* d_B, int ldb, double beta, double* d_C, int ldc, int slices, void* d_work,
                     size_t work_bytes, void* stream);
/* Cholesky K = L L^T plus explicit L^-1 and log|K| (replaces the LU inside jnp.linalg.solve /
 * slogdet, model_GP_solver_2d.py:104-105,158-161).  d_K (n x n) is destroyed; d_L, d_Linv are
 * n x n lower-triangular outputs; d_logdet gets one double; d_status one int (0 = SPD).
 * d_work must hold gphm_potrf_work_bytes(n) bytes.                                            */
GPHM_API size_t gphm_potrf_work_bytes(int n);
GPHM_API int gphm_potrf_inv(double* d_K, int n, double* d_L, double* d_Linv, double* d_logdet, int* d_status,
                   void* d_work, void* stream);
/* The same solve / slogdet pair for a symmetric positive definite TOEPLITZ matrix given by its first
 * column d_t (n entries, jitter already added) - what K is on the uniform grids of every reference
 * config (np.linspace, model_GP_solver_2d.py:360-362).  No dense factorisation: Schur/Levinson
 * recursion for d_g = K^-1 e_0 (n) and log|K|, then each of the `rows` right-hand sides (rows of
 * d_B, rows x n) is solved by the Gohberg-Semencul formula as four FFT convolutions -> d_X.
 * d_sKinv (n) gets the diagonal sums of K^-1 (d > 0: both triangles) that the theta-gradient of
 * log|K| needs.  n <= 4096; d_status: 0 = SPD, -1 = SPD but below the conditioning guard of this route
 * (min_k (1 - kappa_k^2) < 3.5e-5, see GPHM_ILL_CONDITIONED), else 1 + first step with a non-positive prediction
 * error; d_work holds gphm_toeplitz_work_bytes(n, rows) bytes; d_X may not alias d_B.             */
GPHM_API size_t gphm_toeplitz_work_bytes(int n, int rows);
GPHM_API int gphm_toeplitz_solve(const double* d_t, int n, const double* d_B, int rows, double* d_X, double* d_g,
                        double* d_sKinv, double* d_logdet, int* d_status, void* d_work, void* stream);

/* ---- plan: one solver instance ---------------------------------------------------------------
 * Host inputs are copied to the device once.  h_x (n1), h_y (n2; NULL when dim == 1),
 * h_src (n1*n2, row-major), h_bvals (nb; 2-D edge order U[0,:],U[-1,:],U[:,0],U[:,-1] as in
 * model_GP_solver_2d.py:127,377-379; 1-D: the values y), h_xind (nb ints, 1-D only: Xind).
 * d_workspace may be NULL (the library cudaMallocs gphm_workspace_bytes(desc) itself).        */
GPHM_API size_t gphm_workspace_bytes(const gphm_problem_desc* desc);
GPHM_API int gphm_plan_create(const gphm_problem_desc* desc, const double* h_x, const double* h_y, const double* h_src,
                     const double* h_bvals, const int* h_xind, void* d_workspace, size_t workspace_bytes,
                     gphm_plan** out);
GPHM_API void gphm_plan_destroy(gphm_plan* plan);
/* Synchronises `stream`, returns GPHM_OK / GPHM_NOT_SPD / GPHM_NONFINITE / GPHM_ILL_CONDITIONED; *pivot = 1 + index
 * of the first bad pivot (axis 1: 1..n1, axis 2: n1+1..) or 0 (GPHM_ILL_CONDITIONED: bit mask of the flagged axes).
 * Clears the flags.                                                                            */
GPHM_API int gphm_plan_status(gphm_plan* plan, int* pivot, void* stream);
/* Moves every uniform-grid axis of the plan from the Toeplitz inverse generator (Schur/Levinson + Gohberg-Semencul
 * FFT products) to the blocked Cholesky + triangular-GEMM route (force_general bit 4) for all following calls: the
 * answer to GPHM_ILL_CONDITIONED.  Both routes replace the same jnp.linalg.solve / slogdet
 * (model_GP_solver_2d.py:104-105,158-161); the plan's workspace already holds the dense buffers.   */
GPHM_API int gphm_plan_use_cholesky(gphm_plan* plan);
GPHM_API int gphm_plan_uses_toeplitz(const gphm_plan* plan, int axis);
/* Frozen base field (n1*n2 host doubles, NULL clears it): the nonlinearity becomes nl(U + base).
 * This is what the second stage of GP_solver_1d_extra needs - u of the frozen first GP inside
 * (u + u_extra)((u + u_extra)^2 - 1), model_GP_solver_1d_extra.py:95-100.                     */
GPHM_API int gphm_plan_set_base_field(gphm_plan* plan, const double* h_base);

/* Parameter layout (the reference params pytree, model_GP_solver_2d.py:245-261, _1d.py:203-213):
 *   d_U     : n1*n2 doubles (params['U'], or params['u'] (N,1) in 1-D)
 *   d_small : 6Q+2 doubles = [log-w1|log-ls1|freq1|log-w2|log-ls2|freq2|log_tau|log_v]
 *             (kernel_paras_1 / kernel_paras_2; in 1-D the second triple is ignored, grads 0)
 * Gradients use the same layout.  d_terms receives 8 doubles:
 *   [loss, log|K1|, log|K2|, quad=<K1^-1 U, U K2^-1>, boundary_gap, eq_gap, dL/dlog_tau, dL/dlog_v] */

/* value_and_grad(loss): model_GP_solver_2d.py:145-174,179 (1-D: _1d.py:123-149,154;
 * advection: _advection.py:141-170,175).                                                       */
GPHM_API int gphm_logjoint_grad(gphm_plan* plan, const double* d_U, const double* d_small, double* d_gU,
                       double* d_gsmall, double* d_terms, int flags, void* stream);

/* optax.adam update of one leaf, in place (model_GP_solver_2d.py:180-182); d_count is a device
 * int64 holding the number of completed steps (read, not modified).                           */
GPHM_API int gphm_adam_update(double* d_p, const double* d_g, double* d_m, double* d_v, size_t n, const long long* d_count,
                     double lr, void* stream);

/* The same update for the SHORT leaves (kernel parameters, log_tau, log_v) followed by ++count in one launch - the
 * last kernel of a step (optax: the count advances once per optimizer.update, model_GP_solver_2d.py:180).   */
GPHM_API int gphm_adam_update_inc(double* d_p, const double* d_g, double* d_m, double* d_v, size_t n, long long* d_count,
                         double lr, void* stream);

/* step(): value_and_grad + Adam on every leaf, in place, then ++count
 * (model_GP_solver_2d.py:176-183).  d_terms holds the PRE-update loss terms.                   */
GPHM_API int gphm_step(gphm_plan* plan, double* d_U, double* d_small, double* d_mU, double* d_vU, double* d_msmall,
              double* d_vsmall, long long* d_count, double lr, double* d_terms, void* stream);

/* The same step with HOST buffers (functional JAX calling convention: params and opt_state come
 * from and return to host memory).  Copies in, steps, copies out, synchronises.  Pinned buffers let
 * the transfers overlap the kernels: U uploads beside the factor stage (which needs only theta), the
 * Adam moments behind the gradient kernels, and Adam(U) + the download of U and its moments start on a
 * second stream as soon as dL/dU exists, while the theta-gradient is still being computed.        */
GPHM_API int gphm_step_host(gphm_plan* plan, double* h_U, double* h_small, double* h_mU, double* h_vU, double* h_msmall,
                   double* h_vsmall, long long* h_count, double lr, double* h_terms, void* stream);

/* step() with HOST params and a DEVICE-resident opt_state: what `params, opt_state, loss =
 * self.step(params, opt_state, key)` (model_GP_solver_2d.py:176-183, 289) moves per iteration when only the
 * params are consumed on the host - optax's ScaleByAdamState (count, mu, nu) stays where the jitted step left
 * it.  The plan owns the moments and the count; reset_opt != 0 (and the first call) re-initialises them as
 * optimizer.init(params) does (:263).  h_count (may be NULL) receives the updated count.               */
GPHM_API int gphm_step_host_params(gphm_plan* plan, double* h_U, double* h_small, int reset_opt, long long* h_count,
                          double lr, double* h_terms, void* stream);

/* preds(): posterior mean on a test grid (model_GP_solver_2d.py:185-220, _1d.py:160-180).
 * d_xt (m1), d_yt (m2; ignored in 1-D) are device test coordinates; d_out is m1 x m2 (m1 x 1).
 * d_work must hold gphm_predict_work_bytes(plan, m1, m2) bytes.                                */
GPHM_API size_t gphm_predict_work_bytes(const gphm_plan* plan, int m1, int m2);
GPHM_API int gphm_predict(gphm_plan* plan, const double* d_U, const double* d_small, const double* d_xt, int m1,
                 const double* d_yt, int m2, double* d_out, void* d_work, void* stream);
/* ||pred - truth|| / ||truth|| into one device double (model_GP_solver_2d.py:297-300).
 * d_work: gphm_rel_l2_work_bytes() bytes.                                                      */
GPHM_API size_t gphm_rel_l2_work_bytes(void);
GPHM_API int gphm_rel_l2(const double* d_pred, const double* d_truth, size_t n, double* d_out, void* d_work, void* stream);

/* Multi-GPU building blocks (row/column block layouts; the exchange itself is done by the host
 * layer with NCCL between these calls).  See DESIGN.md "multi-GPU".                            */
GPHM_API int gphm_plan_factor(gphm_plan* plan, const double* d_small, int axis_mask, void* stream);
GPHM_API int gphm_apply_kinv(gphm_plan* plan, int axis, int side, const double* d_X, int rows, int cols, double* d_out,
                    double* d_tmp, void* stream);
GPHM_API const double* gphm_plan_matrix(const gphm_plan* plan, int axis, int which);   /* 0 K^-1, 1 D, 2 Linv, 3 L */
/* X K^-1 for every row of d_X (rows x n_axis) with ONE step of iterative refinement on the Toeplitz inverse-generator
 * route (out += K^-1 (X - K out)): the solve-VJPs of the reverse pass (V1, V2) need a small RESIDUAL, not only a small
 * forward error, because the theta-gradient contracts them with a ~5000-fold cancellation.  Same
 * result as gphm_apply_kinv(side = 1) on the Cholesky route.  d_tmp: rows x n scratch; d_X is left intact.
 * Replaces the transposed solves of jax's VJP of jnp.linalg.solve (model_GP_solver_2d.py:104-105 under :179).   */
GPHM_API int gphm_apply_kinv_rows_refined(gphm_plan* plan, int axis, const double* d_X, int rows, double* d_out, double* d_tmp,
                                 void* stream);

/* [log|K1|, log|K2|] of the last factorisation into two device doubles.                        */
GPHM_API int gphm_plan_logdet(gphm_plan* plan, double* d_out2, void* stream);
/* Rank-local pieces of boundary_and_eq_gap / loss / their reverse pass on a block of the grid
 * (model_GP_solver_2d.py:123-174,179), used by the sharded step:
 *  residual : d_R <- e^{log_v} (d_R + nl(U) - F) in place; d_out2 = [sum r^2, sum A*Bt] of the block
 *  boundary : d_eb[e] = U[bidx[e]] - bvals[e]; d_out1 = sum eb^2 (bidx: flat local indices)
 *  grad_u   : d_gU = W + S1 + S2 [+ G(3U^2-1)] + llk_weight e^{log_tau} E_b; d_V2 = S2 + W/2.
 *             bidx/eb as above; indices unique inside [0,nseg0) and [nseg0,nb_local)
 *  theta_grad: sum_ij Kbar dK/dtheta + Dbar dD/dtheta for `axis` (3Q doubles) from caller matrices. */
GPHM_API int gphm_mg_residual(gphm_plan* plan, double* d_R, const double* d_U, const double* d_F, const double* d_A,
                     const double* d_Bt, size_t n_local, const double* d_small, double* d_out2, void* stream);
GPHM_API int gphm_mg_boundary(const double* d_U, const int* d_bidx, const double* d_bvals, int nb_local, double* d_eb,
                     double* d_out1, void* stream);
/* Layout exchange of the sharded step as ONE kernel over NVLink peer memory (csrc/peer.cu), replacing
 * gphm_mg_pack_transposed -> ncclAllToAll -> gphm_mg_unpack_segments:  for every array X_a (rows x cols, cols = P * part_cols)
 * of this rank,  out_d[a][c][me * rows + r] = X_a[r][d * part_cols + c]  is stored straight into rank d's buffer; a sequence
 * number per source tells the consumer (a polling kernel enqueued behind it on `stream`) when all P sources have arrived.
 * Buffers: gphm_mg_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; 64-byte handle to be all-gathered by the host layer),
 * gphm_mg_peer_open on every other rank's handle.  Data of the local result: d_base + 512 bytes, layout [k][part_cols][P*rows].
 * This is the "NCCL all-gather of the column contraction over NVLink" step of north_star, done with in-kernel peer stores.  */
GPHM_API int gphm_mg_peer_alloc(size_t data_doubles, void** d_base, unsigned char* h_handle64);
GPHM_API int gphm_mg_peer_open(const unsigned char* h_handle64, void** d_peer_base);
GPHM_API int gphm_mg_peer_close(void* d_peer_base);
GPHM_API int gphm_mg_peer_free(void* d_base);
GPHM_API int gphm_mg_peer_exchange(const double* const* h_in, int k, int rows, int cols, int part_cols, void* const* h_peer_bases,
                          int P, int me, unsigned long long seq, size_t data_doubles, int* d_status, void* stream);

/* Layout exchange of the sharded step (the block transposes around the NCCL all-to-all; the reference has
 * no counterpart - its jnp.matmul / solve calls at model_GP_solver_2d.py:104-119 see whole matrices):
 *  pack_transposed : d_out[(c / part_cols) * part_stride + (c % part_cols) * rows + r] = d_in[r * cols + c]
 *                    (the transpose of a rows x cols block, cut into parts of part_cols rows, one per destination rank)
 *  unpack_segments : d_out[a][r][s * seg + c] = d_recv[s][a][r][c]   (parts x arrays x rows x seg -> arrays x rows x parts*seg) */
GPHM_API int gphm_mg_pack_transposed(const double* d_in, int rows, int cols, int part_cols, size_t part_stride, double* d_out,
                            void* stream);
GPHM_API int gphm_mg_unpack_segments(const double* d_recv, int parts, int arrays, int rows, int seg, double* d_out, void* stream);
/*  finalize : the eight loss terms [loss, logdet1, logdet2, quad, boundary_gap, eq_gap, dL/dlog_tau, dL/dlog_v]
 *             (model_GP_solver_2d.py:158-174) from the all-reduced sums [eq_gap, quad, boundary_gap] and the two
 *             log-dets; d_gsmall (may be NULL) receives the two scalar gradients at [6Q], [6Q+1].        */
GPHM_API int gphm_mg_finalize(gphm_plan* plan, const double* d_sums3, const double* d_ld2, const double* d_small,
                     double* d_terms, double* d_gsmall, void* stream);
GPHM_API int gphm_mg_grad_u(gphm_plan* plan, const double* d_U, const double* d_G, const double* d_W, const double* d_S1,
                   const double* d_S2, size_t n_local, const int* d_bidx, const double* d_eb, int nseg0, int nb_local,
                   const double* d_small, double* d_gU, double* d_V2, void* stream);
GPHM_API int gphm_mg_theta_grad(gphm_plan* plan, int axis, const double* d_Kbar, const double* d_Dbar, const double* d_small,
                       double* d_gtheta, void* stream);
/* Uniform-grid variant of the rank-local theta-gradient that never forms Kbar/Dbar: d_X, d_Y, d_G
 * are `rows` x n_axis row-major blocks whose rows are the sequences to correlate (axis 2: rows of
 * V2, Bt, G; axis 1: rows of V1^T, A^T, G^T), i.e. Kbar = beta*Linv[r0:r1]^T Linv[r0:r1] - X^T Y and
 * Dbar = cD * G^T Y.  Needs gphm_plan_uses_fft(plan, axis).  gphm_plan_factor: axis_mask bit 2
 * skips forming K^-1 (not needed on this path).                                                 */
GPHM_API int gphm_plan_uses_fft(const gphm_plan* plan, int axis);
/* 1 when the axis' K^-1 (jnp.linalg.solve / slogdet, model_GP_solver_2d.py:104-105,158-161) is applied through
 * the Toeplitz inverse generator (Schur/Levinson + Gohberg-Semencul): gphm_plan_factor is then O(n^2) and
 * local to every rank (nothing to broadcast), gphm_plan_matrix has no L / Linv / K^-1 to return, and
 * gphm_mg_theta_grad_fft takes the K^-1 diagonal sums from the generator on the rank whose
 * linv_row0 == 0 (the Linv row range only selects that rank).                                   */
GPHM_API int gphm_plan_uses_gs(const gphm_plan* plan, int axis);
GPHM_API int gphm_transpose(const double* d_in, int rows, int cols, double* d_out, void* stream);
/* Uniform-grid derivative-Gram product without the Gram matrix: every row x of d_X (rows x n_axis)
 * becomes alpha * D x (transposed = 0) or alpha * D^T x (transposed = 1), + beta * d_out row, by
 * FFT circulant convolution with the axis' Toeplitz table, rebuilt from d_small (so the rank need
 * not have factored this axis itself).
 * (D A = rows of A^T; Bt D^T = rows of Bt; D^T G = rows of G^T with transposed = 1; G D = rows of G, transposed = 1.) */
GPHM_API int gphm_mg_toeplitz_apply(gphm_plan* plan, int axis, int transposed, const double* d_X, int rows, double alpha,
                           double beta, const double* d_small, double* d_out, void* stream);
GPHM_API int gphm_mg_theta_grad_fft(gphm_plan* plan, int axis, const double* d_X, const double* d_Y, const double* d_G, int rows,
                           int linv_row0, int linv_row1, double beta, double cD, const double* d_small,
                           double* d_gtheta, void* stream);
/* Rank-local pieces of the all-FFT step (every axis on the Toeplitz inverse generator, gphm_plan_uses_gs):
 * gphm_mg_toeplitz_rows: d_out[r] = alpha * D x_r (D^T x_r when transposed) + beta * d_add[r] (d_add NULL: beta *
 *   d_out[r]) for the `rows` rows of d_X, with the derivative-Gram spectrum built by gphm_plan_factor;
 *   keep_spectrum != 0 stores the transform of every packed row pair of d_X in the plan (one store per axis).
 * gphm_mg_theta_grad_pairs: theta-gradient of  lead*beta*K^-1 - V^T Y  and  cD * G^T Y, where Y are the rows whose
 *   transforms the last keep_spectrum call stored (same row count and order), V and G `rows` x n_axis blocks;
 *   lead != 0 on exactly one rank.  gphm_mg_grad_u: d_S2 and d_V2 may be NULL (dU = W + S1).            */
GPHM_API int gphm_mg_toeplitz_rows(gphm_plan* plan, int axis, int transposed, const double* d_X, int rows, double alpha,
                          double beta, const double* d_add, double* d_out, int keep_spectrum, void* stream);
GPHM_API int gphm_mg_theta_grad_pairs(gphm_plan* plan, int axis, const double* d_V, const double* d_G, int rows, int lead,
                             double beta, double cD, const double* d_small, double* d_gtheta, void* stream);
/* Both axes of the all-FFT step in one call: axis 1 gets (d_V1, d_G1: rows1 x n1 blocks, i.e. transposed column blocks),
 * axis 2 (d_V2, d_G2: rows2 x n2); d_gtheta[0..3Q) and [3Q..6Q).  Same result as two gphm_mg_theta_grad_pairs calls; the
 * single-CTA tails (partial-spectrum reduction, inverse transform, theta contraction) of the two axes share launches.  */
GPHM_API int gphm_mg_theta_grad_pairs_both(gphm_plan* plan, const double* d_V1, const double* d_G1, int rows1, const double* d_V2,
                                  const double* d_G2, int rows2, int lead, double beta1, double beta2, double cD1, double cD2,
                                  const double* d_small, double* d_gtheta, void* stream);
/* d_out = a*d_x + b*d_y (d_y may be NULL).                                                      */
GPHM_API int gphm_lincomb(double* d_out, double a, const double* d_x, double b, const double* d_y, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPHM_H_ */
