// Plan object, per-iteration orchestration and the extern "C" boundary of libgphm.
//
// One gphm_plan = one reference solver instance (GP_solver_1d_single, GP_solver_2d_single,
// GP_solver_2d_single_advection): the constants the reference captures through the static
// `self` of its jitted methods plus every device buffer the step needs, so that nothing is
// allocated and nothing touches the host while iterating (CUDA-graph capturable).
//
// Per-iteration dataflow (reference lines in include/gphm.h):
//   per axis a:  K_a, D_a  <- Gram builders        L_a <- chol(K_a)      Linv_a <- L_a^-1
//   A  = K1^-1 U  = Linv1^T (Linv1 U)              Bt = U K2^-1 = (U Linv2^T) Linv2
//   R  = c1 * D1 A + Bt D2^T      G = e^v (R + nl(U) - F)       eqgap, quad, bgap, logdets
//   W  = K1^-1 Bt     S1 = K1^-1 (c1 D1^T G)     S2 = (G D2) K2^-1
//   dU = W + S1 + S2 [+ G (3U^2-1)] + lambda e^tau E_b
//   Kbar1 = ld/2 N2 K1^-1 - (S1 + W/2) A^T        Dbar1 = c1 G A^T
//   Kbar2 = ld/2 N1 K2^-1 - (S2 + W/2)^T Bt       Dbar2 = G^T Bt
//   dtheta_a = sum_ij Kbar_a * dK/dtheta + Dbar_a * dD/dtheta   (diagonal sums on uniform grids)
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/gphm.h"
#include "common.cuh"
#include "kernels.h"

namespace gphm {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- launch accounting / event profiling -------------------------------------------------------
namespace {
struct ProfRec { int cat; double flops, bytes; cudaEvent_t a, b; };
std::atomic<long long> g_launches{0};
bool g_prof = false;
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t pool_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace
bool profiling_enabled() { return g_prof; }
LaunchScope::LaunchScope(int c, cudaStream_t s, double flops, double bytes) : cat(c), st(s), slot(-1) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_prof) {
        ProfRec r{c, flops, bytes, pool_event(), pool_event()};
        cudaEventRecord(r.a, st);
        slot = (int)g_recs.size();
        g_recs.push_back(r);
    }
}
LaunchScope::~LaunchScope() {
    if (slot >= 0) cudaEventRecord(g_recs[slot].b, st);
}

struct Axis {
    int n = 0, nblk = 0, fftL = 0;   // fftL: FFT length of the diagonal-sum path (0 = GEMM path)
    bool toeplitz = false;
    bool gs = false;                 // K^-1 by Schur/Levinson + Gohberg-Semencul FFT products (no dense factorisation)
    double dirsign = 1.0;
    double *x = nullptr, *K = nullptr, *D = nullptr, *L = nullptr, *Linv = nullptr, *Kinv = nullptr, *Dbar = nullptr,
           *T = nullptr, *invdiag = nullptr, *ldpart = nullptr, *tabK = nullptr, *tabD = nullptr, *dspart = nullptr,
           *sK = nullptr, *sD = nullptr, *sKinv = nullptr, *tgpart = nullptr, *twid = nullptr, *specK = nullptr, *specD = nullptr, *specT = nullptr,   // specT: spectrum of the Toeplitz D
           *gsg = nullptr, *gspec = nullptr,   // gsg = K^-1 e_0; gspec: four Gohberg-Semencul circulant spectra
           *specKm = nullptr,                  // spectrum of K itself (circulant embedding): residuals of the refined K^-1 applications
           *specY = nullptr,                   // transforms of the packed row pairs of A^T (axis 1) / Bt (axis 2)
           *gskap = nullptr,                   // reflection coefficients handed from the generator CTA to the lattice CTA
           *gsbnd = nullptr;                   // boundary values handed upward between the CTAs of a split role (6 n)
    int* gsprog = nullptr;
};

}  // namespace gphm

using namespace gphm;

struct gphm_plan {
    gphm_problem_desc d;
    Axis ax[2];
    double *src = nullptr, *bvals = nullptr, *base = nullptr;   // base: optional frozen field added inside nl(.)
    bool has_base = false;
    int* xind = nullptr;
    double *A = nullptr, *Bt = nullptr, *Tf = nullptr, *R = nullptr, *W = nullptr, *P = nullptr, *S1 = nullptr,
           *S2 = nullptr, *V1 = nullptr, *V2 = nullptr, *gU = nullptr;
    double *part = nullptr, *eb = nullptr, *gsmall = nullptr, *terms = nullptr;
    double* gsS = nullptr;           // field-sized scratch of the Gohberg-Semencul K^-1 application
    int* status = nullptr;
    void* ws = nullptr;
    bool owns_ws = false;
    bool size_query = false;   // carve(): reserve the larger theta-gradient scratch per axis
    // device staging for gphm_step_host (allocated on first use)
    double* hs = nullptr;
    long long* hs_count = nullptr;
    cudaStream_t hs_stream = nullptr;     // second copy stream: the Adam moments travel while the gradient is computed
    cudaEvent_t hs_ev_u = nullptr, hs_ev_mv = nullptr, hs_ev_gu = nullptr, hs_ev_adam = nullptr;
    // hooks of gphm_step_host into the all-FFT step (null otherwise):
    cudaEvent_t u_ready = nullptr;        // waited for on the step's stream after the factor stage (U still uploading)
    static constexpr int kUChunks = 4;    // 2-D all-FFT step: U arrives in row blocks, Bt = U K2^-1 follows block by block
    cudaEvent_t hs_ev_chunk[kUChunks] = {};
    int u_chunks = 0;                     // > 0: hs_ev_chunk[c] marks the arrival of rows [c n1 / u_chunks, (c+1) n1 / u_chunks)
    int (*on_gu)(gphm_plan&, cudaStream_t) = nullptr;   // called once dL/dU is complete (before the theta-gradient)
    bool gu_hook_ran = false;
    // Look-ahead of the factor stage (gphm_step on large 2-D uniform plans): the theta-only work of step t+1 (tables, recursion,
    // spectra - 0.46 ms on 16 CTAs at 4096^2) runs on lk_stream beside dL/dU assembly + Adam(U) of step t.
    bool lk_enabled = false, defer_gu = false;
    cudaStream_t lk_stream = nullptr;
    cudaEvent_t lk_fork = nullptr, lk_join = nullptr;
    double* lk_small = nullptr;          // theta the look-ahead factored (device, 6Q+2)
    int* lk_flags = nullptr;             // [0] look-ahead result valid, [1] skip flag of the current factor stage
    // workspace of the tcgen05 Ozaki GEMM (force_general bit 6), allocated on first use
    void* oz_ws = nullptr;
    size_t oz_ws_bytes = 0;
};

namespace {

constexpr size_t kAlign = 256;

// Walks the workspace: with base == nullptr only sizes are accumulated.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base(static_cast<char*>(b)) {}
    template <typename T>
    void take(T*& p, size_t count) {
        const size_t bytes = (count * sizeof(T) + kAlign - 1) / kAlign * kAlign;
        p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += bytes;
    }
};

bool uniform_grid(const double* x, int n) {
    if (n < 3) return true;
    const double h0 = x[1] - x[0];
    double hmax = 0.0, dev = 0.0;
    for (int i = 1; i < n; ++i) {
        const double h = x[i] - x[i - 1];
        hmax = std::max(hmax, std::fabs(h));
        dev = std::max(dev, std::fabs(h - h0));
    }
    return h0 != 0.0 && dev <= 1e-9 * hmax;
}

size_t carve(gphm_plan& p, void* base) {
    Carver c(base);
    const gphm_problem_desc& d = p.d;
    const size_t nf = (size_t)d.n1 * d.n2;
    const int naxes = d.dim == 2 ? 2 : 1;
    for (int a = 0; a < naxes; ++a) {
        Axis& X = p.ax[a];
        const size_t n = X.n, nn = n * n;
        c.take(X.x, n);
        c.take(X.K, nn); c.take(X.D, nn); c.take(X.L, nn); c.take(X.Linv, nn); c.take(X.Kinv, nn);
        c.take(X.Dbar, nn); c.take(X.T, nn);
        c.take(X.invdiag, (size_t)X.nblk * kNB * kNB);
        c.take(X.ldpart, X.nblk);
        c.take(X.tabK, n); c.take(X.tabD, n);
        c.take(X.sK, n); c.take(X.sD, n); c.take(X.sKinv, n);
        const size_t ds = diag_sums_part_doubles((int)n), tg = theta_general_part_doubles((int)n, d.Q);
        const size_t Lq = (size_t)fft_length_for((int)n);
        const size_t spec = 2 * Lq * fft_grid();              // complex partial spectra, one per CTA
        if (p.size_query) { double* dummy; c.take(dummy, std::max(ds, tg)); c.take(dummy, 2 * Lq); c.take(dummy, spec); c.take(dummy, spec); c.take(dummy, 2 * Lq); }
        else if (X.toeplitz) {
            c.take(X.dspart, ds);
            if (X.fftL > 0) { c.take(X.twid, 2 * (size_t)X.fftL); c.take(X.specK, 2 * (size_t)X.fftL * fft_grid()); c.take(X.specD, 2 * (size_t)X.fftL * fft_grid()); c.take(X.specT, 2 * (size_t)X.fftL); }
        } else c.take(X.tgpart, tg);
        if (p.size_query || X.gs) {
            const size_t other = (d.dim == 2) ? (a == 0 ? (size_t)d.n2 : (size_t)d.n1) : 1;      // rows this axis' operators act on
            c.take(X.gsg, n); c.take(X.gspec, 8 * Lq); c.take(X.specKm, 2 * Lq); c.take(X.specY, 2 * Lq * ((other + 1) / 2));
            c.take(X.gskap, n); c.take(X.gsprog, 8); c.take(X.gsbnd, 6 * n);
        }
    }
    if (p.size_query || p.ax[0].gs || p.ax[1].gs) c.take(p.gsS, nf);
    c.take(p.src, nf); c.take(p.base, nf); c.take(p.bvals, d.nb); c.take(p.xind, std::max(d.nb, 1));
    c.take(p.A, nf); c.take(p.Tf, nf); c.take(p.R, nf); c.take(p.P, nf); c.take(p.S1, nf); c.take(p.V1, nf);
    c.take(p.gU, nf);
    if (d.dim == 2) { c.take(p.Bt, nf); c.take(p.W, nf); c.take(p.S2, nf); c.take(p.V2, nf); }
    c.take(p.part, 2 * (size_t)kRedBlocks);
    c.take(p.eb, std::max(d.nb, 1));
    c.take(p.gsmall, 6 * (size_t)d.Q + 2);
    c.take(p.terms, 8);
    c.take(p.status, 4);
    return c.off;
}

int check_desc(const gphm_problem_desc* d) {
    if (!d) { set_last_error("null problem descriptor"); return GPHM_EINVAL; }
    if (d->dim != 1 && d->dim != 2) { set_last_error("dim must be 1 or 2 (got %d)", d->dim); return GPHM_EINVAL; }
    if (d->kernel_id < 0 || d->kernel_id > 3) { set_last_error("Invalid Kernel id %d", d->kernel_id); return GPHM_EINVAL; }
    if (d->eq_type < 0 || d->eq_type > 2) { set_last_error("unknown equation type %d", d->eq_type); return GPHM_EINVAL; }
    if (d->eq_type == GPHM_EQ_ADVECTION && d->dim != 2) { set_last_error("advection needs dim == 2"); return GPHM_EINVAL; }
    if (d->n1 < 1 || d->n2 < 1 || (d->dim == 1 && d->n2 != 1)) { set_last_error("bad grid %d x %d", d->n1, d->n2); return GPHM_EINVAL; }
    if (d->Q < 1 || d->Q > 256) { set_last_error("Q=%d outside [1,256]", d->Q); return GPHM_EINVAL; }
    if (d->dim == 2 && d->nb != 2 * d->n1 + 2 * d->n2) { set_last_error("2-D nb must be 2*n1+2*n2"); return GPHM_EINVAL; }
    if (d->nb < 0) { set_last_error("nb < 0"); return GPHM_EINVAL; }
    return GPHM_OK;
}

void init_axes(gphm_plan& p, const double* hx, const double* hy) {
    p.ax[0].n = p.d.n1; p.ax[1].n = p.d.dim == 2 ? p.d.n2 : 0;
    const double* h[2] = {hx, hy};
    for (int a = 0; a < 2; ++a) {
        Axis& X = p.ax[a];
        X.nblk = X.n > 0 ? num_blocks_nb(X.n) : 0;
        if (X.n > 0 && h[a]) {
            X.toeplitz = !(p.d.force_general & 1) && uniform_grid(h[a], X.n);
            X.dirsign = (h[a][X.n - 1] >= h[a][0]) ? 1.0 : -1.0;
            X.fftL = (X.toeplitz && !(p.d.force_general & 2)) ? fft_length_for(X.n) : 0;
            X.gs = X.fftL > 0 && !(p.d.force_general & (4 | 16)) && X.n <= toeplitz_inv_max_n();
        } else {
            X.toeplitz = !(p.d.force_general & 1);   // size query: assume the (larger) general layout below
        }
    }
}

inline int deriv_order(const gphm_plan& p) { return p.d.eq_type == GPHM_EQ_ADVECTION ? 1 : 2; }
inline double coef_c1(const gphm_plan& p) { return p.d.eq_type == GPHM_EQ_ADVECTION ? p.d.beta : 1.0; }
inline const double* theta_of(const gphm_plan& p, const double* small, int a) { return small + (size_t)a * 3 * p.d.Q; }

LossConsts loss_consts(const gphm_plan& p) {
    LossConsts c;
    c.dim = p.d.dim; c.eq_type = p.d.eq_type; c.n1 = p.d.n1; c.n2 = p.d.n2; c.nb = p.d.nb; c.Q = p.d.Q;
    c.llk_weight = p.d.llk_weight; c.logdet = p.d.logdet; c.c1 = coef_c1(p);
    return c;
}

// A plain (non-solve) contraction of the dense path: native FP64 DMMA GEMM, or - force_general bit 6 - the Ozaki-sliced
// int8 GEMM on tcgen05 (ozaki.cu) with its stated bound.  The triangular products of the K^-1 applications never come here.
int contract(gphm_plan& p, const GemmArgs& g, cudaStream_t st) {
    if (!(p.d.force_general & 64) || g.kmode != 0 || g.batch != 1 || g.K > 65536) return launch_dgemm(g, st);
    const int S = ozaki_default_slices();
    const size_t need = ozaki_work_bytes(g.M, g.N, g.K, S);
    if (need > p.oz_ws_bytes) {
        // sized once for the largest contraction of the plan (max(n1, n2)^3): no allocation inside later steps
        const int n = std::max(std::max(p.d.n1, p.d.n2), std::max(g.M, std::max(g.N, g.K)));
        const size_t want = std::max(need, ozaki_work_bytes(n, n, n, S));
        GPHM_CUDA_OK(cudaStreamSynchronize(st));
        if (p.oz_ws) GPHM_CUDA_OK(cudaFree(p.oz_ws));
        p.oz_ws = nullptr; p.oz_ws_bytes = 0;
        if (cudaMalloc(&p.oz_ws, want) != cudaSuccess) { set_last_error("ozaki workspace: cudaMalloc(%zu) failed", want); return GPHM_ENOMEM; }
        p.oz_ws_bytes = want;
    }
    return launch_ozaki_dgemm(g.transA != 0, g.transB != 0, g.M, g.N, g.K, g.alpha, g.A, g.lda, g.B, g.ldb, g.beta, g.C, g.ldc, S,
                              p.oz_ws, p.oz_ws_bytes, st);
}

// Gram matrices of one axis.
int gram_axis(gphm_plan& p, int a, const double* small, cudaStream_t st) {
    Axis& X = p.ax[a];
    const double* th = theta_of(p, small, a);
    const int n = X.n, order = deriv_order(p);
    if (X.toeplitz)
        return launch_gram_toeplitz(p.d.kernel_id, order, X.x, n, th, p.d.Q, p.d.jitter, X.dirsign, X.tabK, X.tabD,
                                    X.K, X.D, n, st);
    return launch_gram_general(p.d.kernel_id, order, X.x, n, X.x, n, th, p.d.Q, p.d.jitter, X.K, X.D, n, st);
}

// L^-1 (+ K^-1 = Linv^T Linv; the FFT diagonal-sum path works on Linv directly and skips it)
int invert_axis(gphm_plan& p, int a, bool with_kinv, cudaStream_t st) {
    Axis& X = p.ax[a];
    const int n = X.n;
    GPHM_TRY(trtri_lower(X.L, X.Linv, n, n, X.invdiag, X.T, st));
    if (with_kinv)
        GPHM_TRY(launch_dgemm(gemm_args(X.Linv, n, true, X.Linv, n, false, X.Kinv, n, n, n, n, 1.0, 0.0,
                                        KM_A_UPPER | KM_B_LOWER), st));
    return GPHM_OK;
}

// The K^-1 applications of the REVERSE pass (V1 = K1^-1 (c1 D1^T G + Bt/2), V2 = (G D2 + A/2) K2^-1) get one step of
// iterative refinement, y += GS(b - K y).  The Gohberg-Semencul application is forward-accurate (cond(K) eps relative to
// |y|, like a Cholesky solve) but its residual b - K y is not small relative to |b|, and the theta-gradient contracts
// V A^T against dK/dtheta with a ~5000-fold cancellation against G A^T : dD/dtheta that only a small RESIDUAL keeps
// intact: without refinement the theta-leaves are off by 5e-8 at N = 1024 and 1.4e-6 at N = 4096 against an
// extended-precision reference (tools/extended_reference.py), with it they match the Cholesky route (1e-9 .. 5e-8).
// The forward applications (A, Bt) do not need it (measured: no change).  The loss grows like cond(K) N: 5e-8 at N = 1024,
// 3.2e-7 at 2048, 1.4e-6 at 4096 (same kernel and state), so axes of up to kRefineAbove = 512 points skip the step
// (the reference's N_col = 400 configs and the ensemble; its N_col = 900 configs get it).  force_general bit 5 switches it off
// everywhere, bit 7 on everywhere (measurements: tools/dump_gpu_grad.py).
constexpr int kRefineAbove = 512;
inline bool gs_refine(const gphm_plan& p) { return (p.d.force_general & 32) == 0; }
inline bool gs_refine_axis(const gphm_plan& p, const Axis& X) { return gs_refine(p) && (X.n > kRefineAbove || (p.d.force_general & 128)); }

// Uniform-grid axes a0 .. a0+count-1 without a dense factorisation: Toeplitz tables, Schur/Levinson
// recursion for g = K^-1 e_0 and log|K|, Gohberg-Semencul spectra and the diagonal sums of K^-1.
// Two axes of equal size share every launch (one CTA per axis).
int factor_gs(gphm_plan& p, int a0, int count, const double* small, cudaStream_t st, const int* skip = nullptr) {
    const int order = deriv_order(p);
    // any factor stage that is not guarded by the look-ahead comparison overwrites the buffers: a stored look-ahead dies
    if (!skip && p.lk_flags) GPHM_TRY(launch_lk_set(p.lk_flags, 0, st));
    const bool need_D = (p.d.force_general & 8) != 0;      // derivative-Gram products by GEMM want the full D
    for (int a = a0; a < a0 + count; ++a) {
        Axis& X = p.ax[a];
        const double* th = theta_of(p, small, a);
        if (need_D)
            GPHM_TRY(launch_gram_toeplitz(p.d.kernel_id, order, X.x, X.n, th, p.d.Q, p.d.jitter, X.dirsign, X.tabK, X.tabD,
                                          X.K, X.D, X.n, st));
        else
            GPHM_TRY(launch_toeplitz_table(p.d.kernel_id, order, X.x, X.n, th, p.d.Q, X.tabK, X.tabD, st, skip));
    }
    Axis& X0 = p.ax[a0];
    const bool batched = count == 2 && p.ax[a0 + 1].n == X0.n;
    for (int a = a0; a < a0 + (batched ? 1 : count); ++a) {
        Axis& X = p.ax[a];
        const int nsys = batched ? 2 : 1;
        const Axis& Y = p.ax[batched ? a + 1 : a];
        GPHM_TRY(launch_schur_levinson(X.tabK, Y.tabK - X.tabK, X.n, p.d.jitter, X.gsg, Y.gsg - X.gsg, X.ldpart,
                                       Y.ldpart - X.ldpart, p.status + a, 1, X.gskap, Y.gskap - X.gskap, X.gsprog,
                                       Y.gsprog - X.gsprog, nsys, st, nullptr, p.status + 3, a, X.gsbnd, Y.gsbnd - X.gsbnd, skip));
        GPHM_TRY(launch_gs_prepare(X.gsg, Y.gsg - X.gsg, X.n, X.fftL, X.twid, X.gspec, Y.gspec - X.gspec, X.sKinv,
                                   Y.sKinv - X.sKinv, nsys, st, skip));
    }
    // spectra of the Toeplitz derivative Gram (FFT products) and of K with the jitter (residual b - K y of the refined
    // applications): independent single-CTA transforms - axes of equal transform length share one launch
    ToeplitzSpectrumJob jobs[4];
    int nj = 0, jobL = 0;
    auto flush = [&]() -> int {
        const int rc = nj ? launch_toeplitz_spectrum_multi(jobs, nj, jobL, st, skip) : GPHM_OK;
        nj = 0;
        return rc;
    };
    for (int a = a0; a < a0 + count; ++a) {
        Axis& X = p.ax[a];
        if (nj && (X.fftL != jobL || nj > 2)) GPHM_TRY(flush());
        jobL = X.fftL;
        if (!need_D) jobs[nj++] = ToeplitzSpectrumJob{X.tabD, X.n, X.twid, order == 1, X.dirsign, 0.0, X.specT};
        if (gs_refine_axis(p, X)) jobs[nj++] = ToeplitzSpectrumJob{X.tabK, X.n, X.twid, false, 1.0, p.d.jitter, X.specKm};
    }
    return flush();
}

// Gram + Cholesky + L^-1 (+ K^-1) for one axis.
int factor_axis(gphm_plan& p, int a, const double* small, bool with_kinv, cudaStream_t st) {
    Axis& X = p.ax[a];
    if (X.gs) return factor_gs(p, a, 1, small, st);
    GPHM_TRY(gram_axis(p, a, small, st));
    GPHM_TRY(chol_factor(X.K, X.L, X.n, X.n, X.invdiag, X.ldpart, p.status + a, st));
    return invert_axis(p, a, with_kinv, st);
}

// Both axes of a 2-D problem.  Equal sizes: the two Cholesky chains share every launch.
int factor_both(gphm_plan& p, const double* small, bool kinv0, bool kinv1, cudaStream_t st) {
    Axis& X0 = p.ax[0];
    Axis& X1 = p.ax[1];
    if (X0.gs && X1.gs) return factor_gs(p, 0, 2, small, st);
    if (X0.n != X1.n || X0.gs || X1.gs) {
        GPHM_TRY(factor_axis(p, 0, small, kinv0, st));
        return factor_axis(p, 1, small, kinv1, st);
    }
    GPHM_TRY(gram_axis(p, 0, small, st));
    GPHM_TRY(gram_axis(p, 1, small, st));
    GPHM_TRY(chol_factor_multi(X0.K, X0.L, X0.n, X0.n, X0.invdiag, X0.ldpart, p.status, 2, X1.K - X0.K, X1.L - X0.L,
                               X1.invdiag - X0.invdiag, X1.ldpart - X0.ldpart, 1, st));
    GPHM_TRY(invert_axis(p, 0, kinv0, st));
    return invert_axis(p, 1, kinv1, st);
}

// rows x n matrix Xm -> out = Xm K_a^-1 (every row v -> K^-1 v) by the Gohberg-Semencul formula:
// four circulant convolutions per row.  tmp: rows x n scratch; out may not alias Xm.
int apply_kinv_rows_gs(const Axis& X, const double* Xm, int rows, double* out, double* tmp, cudaStream_t st) {
    const int n = X.n, L = X.fftL;
    const size_t sp = 2 * (size_t)L;
    if (toeplitz_fused_supported(L))
        return launch_gs_apply_fused(Xm, rows, n, n, X.gspec, L, X.twid, 1.0, 0.0, nullptr, 0, out, n, st, X.gsg);
    GPHM_TRY(launch_toeplitz_apply(Xm, rows, n, n, X.gspec, L, X.twid, 1.0, 0.0, out, n, st));            // L(g)^T v
    GPHM_TRY(launch_toeplitz_apply(Xm, rows, n, n, X.gspec + sp, L, X.twid, 1.0, 0.0, tmp, n, st));       // L(h)^T v
    GPHM_TRY(launch_toeplitz_apply(out, rows, n, n, X.gspec + 2 * sp, L, X.twid, 1.0, 0.0, out, n, st));  // L(g) . / g0
    GPHM_TRY(launch_toeplitz_apply(tmp, rows, n, n, X.gspec + 3 * sp, L, X.twid, 1.0, 1.0, out, n, st));  // - L(h) . / g0
    return GPHM_OK;
}

// out[r] += K^-1 (Xm[r] - K out[r]) for every row: one refinement step of out ~ K^-1 Xm.  tmp: rows x n scratch (may alias Xm,
// which is then destroyed).
int refine_kinv_rows_gs(const Axis& X, const double* Xm, int rows, double* out, double* tmp, cudaStream_t st) {
    const int n = X.n, L = X.fftL;
    GPHM_TRY(launch_toeplitz_apply_fused(out, rows, n, n, X.specKm, L, X.twid, -1.0, 1.0, Xm, n, tmp, n, nullptr, st));   // b - K y
    return launch_gs_apply_fused(tmp, rows, n, n, X.gspec, L, X.twid, 1.0, 1.0, out, n, out, n, st, X.gsg);               // y += K^-1 r
}

// out = K_a^-1 X (side 0, X is n x cols) or X K_a^-1 (side 1, X is rows x n); tmp has X's shape.
int apply_kinv(gphm_plan& p, int a, int side, const double* Xm, int rows, int cols, double* out, double* tmp,
               cudaStream_t st) {
    const Axis& X = p.ax[a];
    const int n = X.n;
    if (X.gs) {
        if ((side == 0 ? rows : cols) != n) { set_last_error("apply_kinv: operand %d x %d does not match n %d", rows, cols, n); return GPHM_EINVAL; }
        if (side == 1) return apply_kinv_rows_gs(X, Xm, rows, out, tmp, st);
        if ((size_t)rows * cols > (size_t)p.d.n1 * p.d.n2) { set_last_error("apply_kinv: %d x %d exceeds the plan's scratch", rows, cols); return GPHM_EINVAL; }
        GPHM_TRY(launch_transpose(Xm, n, cols, p.gsS, st));             // columns -> rows
        GPHM_TRY(apply_kinv_rows_gs(X, p.gsS, cols, tmp, out, st));     // `out` is scratch here
        return launch_transpose(tmp, cols, n, out, st);
    }
    if (side == 0) {
        if (rows != n) { set_last_error("apply_kinv: rows %d != n %d", rows, n); return GPHM_EINVAL; }
        GPHM_TRY(launch_dgemm(gemm_args(X.Linv, n, false, Xm, cols, false, tmp, cols, n, cols, n, 1.0, 0.0, KM_A_LOWER), st));
        GPHM_TRY(launch_dgemm(gemm_args(X.Linv, n, true, tmp, cols, false, out, cols, n, cols, n, 1.0, 0.0, KM_A_UPPER), st));
    } else {
        if (cols != n) { set_last_error("apply_kinv: cols %d != n %d", cols, n); return GPHM_EINVAL; }
        GPHM_TRY(launch_dgemm(gemm_args(Xm, n, false, X.Linv, n, true, tmp, n, rows, n, n, 1.0, 0.0, KM_B_UPPER), st));
        GPHM_TRY(launch_dgemm(gemm_args(tmp, n, false, X.Linv, n, false, out, n, rows, n, n, 1.0, 0.0, KM_B_LOWER), st));
    }
    return GPHM_OK;
}

// Uniform grids, every axis on the Toeplitz inverse generator: the whole iteration is FFT work.
// Four K^-1 applications instead of five: with P1 = c1 D1^T G and P2 = G D2,
//   V1 = S1 + W/2 = K1^-1 (P1 + Bt/2),   V2 = S2 + W/2 = (P2 + A/2) K2^-1,   dU = V1 + V2 + ...
// and axis-1 operands stay transposed (rows = columns of the field) from U^T to the diagonal sums.
int logjoint_grad_gs(gphm_plan& p, const double* U, const double* small, double* gU, double* gsmall, double* terms,
                     int flags, cudaStream_t st) {
    const gphm_problem_desc& d = p.d;
    const bool two = d.dim == 2, fwd_only = (flags & GPHM_FORWARD_ONLY) != 0;
    const int n1 = d.n1, n2 = d.n2, Q = d.Q;
    const size_t nf = (size_t)n1 * n2;
    const double c1 = coef_c1(p);
    const int order = deriv_order(p);
    const bool anti = order == 1;
    Axis& X1 = p.ax[0];
    Axis& X2 = p.ax[1];
    if (p.lk_enabled) {                                      // did the previous step's look-ahead factor exactly this theta?
        GPHM_TRY(launch_lk_compare(small, p.lk_small, 6 * Q + 2, p.lk_flags, st));
        GPHM_TRY(factor_gs(p, 0, two ? 2 : 1, small, st, p.lk_flags + 1));
    } else {
        GPHM_TRY(factor_gs(p, 0, two ? 2 : 1, small, st));   // includes the spectra of D1, D2; needs only theta
    }
    bool bt_done = false;
    if (p.u_ready && two && p.u_chunks > 0) {
        // gphm_step_host: everything that acts on ROWS of U runs row block by row block behind the upload -
        // Bt = U K2^-1, R = Bt D2^T (with the transforms of Bt's rows kept) and the block's share of U^T
        const int rows = n1 / p.u_chunks;
        for (int c = 0; c < p.u_chunks; ++c) {
            const size_t o = (size_t)c * rows * n2;
            GPHM_CUDA_OK(cudaStreamWaitEvent(st, p.hs_ev_chunk[c], 0));
            GPHM_TRY(launch_gs_apply_fused(U + o, rows, n2, n2, X2.gspec, X2.fftL, X2.twid, 1.0, 0.0, nullptr, 0, p.Bt + o, n2, st, X2.gsg));
            GPHM_TRY(launch_toeplitz_apply_fused(p.Bt + o, rows, n2, n2, X2.specT, X2.fftL, X2.twid, 1.0, 0.0, nullptr, n2, p.R + o, n2,
                                                 X2.specY + (size_t)(c * rows / 2) * X2.fftL * 2, st));
            GPHM_TRY(launch_transpose(U + o, rows, n2, p.Tf + (size_t)c * rows, st, n2, n1));
        }
        bt_done = true;
    }
    if (p.u_ready) GPHM_CUDA_OK(cudaStreamWaitEvent(st, p.u_ready, 0));     // gphm_step_host: U arrives meanwhile
    auto gs1 = [&](const double* Xr, double* out) {       // rows of length n1 (columns of the field)
        return launch_gs_apply_fused(Xr, n2, n1, n1, X1.gspec, X1.fftL, X1.twid, 1.0, 0.0, nullptr, 0, out, n1, st, X1.gsg);
    };
    auto gs2 = [&](const double* Xr, double* out) {       // rows of length n2
        return launch_gs_apply_fused(Xr, n1, n2, n2, X2.gspec, X2.fftL, X2.twid, 1.0, 0.0, nullptr, 0, out, n2, st, X2.gsg);
    };
    auto d1 = [&](const double* Xr, double alpha, double beta, const double* add, double* out, double* spec_out = nullptr) {
        return launch_toeplitz_apply_fused(Xr, n2, n1, n1, X1.specT, X1.fftL, X1.twid, alpha, beta, add, n1, out, n1, spec_out, st);
    };
    auto d2 = [&](const double* Xr, double alpha, double beta, const double* add, double* out, double* spec_out = nullptr) {
        return launch_toeplitz_apply_fused(Xr, n1, n2, n2, X2.specT, X2.fftL, X2.twid, alpha, beta, add, n2, out, n2, spec_out, st);
    };
    // ---- forward ----
    const double* Ut = U;                                  // 1-D: the field is one row already
    if (two) { if (!bt_done) GPHM_TRY(launch_transpose(U, n1, n2, p.Tf, st)); Ut = p.Tf; }
    double* At = p.P;
    GPHM_TRY(gs1(Ut, At));                                 // A^T = (K1^-1 U)^T
    const double* Bt = U;
    const double* A = At;
    if (two) {
        if (!bt_done) GPHM_TRY(gs2(U, p.Bt));              // Bt = U K2^-1
        Bt = p.Bt;
        GPHM_TRY(d1(At, c1, 0.0, nullptr, p.Tf, X1.specY));   // (c1 D1 A)^T; keeps the transforms of A^T's rows
        if (bt_done) {
            GPHM_TRY(launch_transpose(p.Tf, n2, n1, p.R, st, 0, 0, true));   // R = Bt D2^T is there already: same sum, same bits
        } else {
            GPHM_TRY(launch_transpose(p.Tf, n2, n1, p.R, st));
            GPHM_TRY(d2(Bt, 1.0, 1.0, nullptr, p.R, X2.specY));   // + Bt D2^T; keeps the transforms of Bt's rows
        }
        GPHM_TRY(launch_transpose(At, n2, n1, p.A, st)); A = p.A;
    } else {
        GPHM_TRY(d1(At, c1, 0.0, nullptr, p.R, X1.specY));
    }
    GPHM_TRY(launch_residual(p.R, U, p.src, A, Bt, nf, d.eq_type, p.has_base ? p.base : nullptr, small, Q, p.part, st));    // R <- G
    const LossConsts lc = loss_consts(p);
    GPHM_TRY(launch_finalize(lc, U, p.bvals, p.xind, p.part, X1.ldpart, X1.nblk, two ? X2.ldpart : nullptr,
                             two ? X2.nblk : 0, small, p.eb, terms, fwd_only ? nullptr : gsmall, p.status, st));
    if (fwd_only) return GPHM_OK;
    // ---- backward ----
    const double* G = p.R;
    const double* Gt = G;
    const double *V1t, *V2 = nullptr;
    if (two) {
        GPHM_TRY(launch_transpose(G, n1, n2, p.S1, st)); Gt = p.S1;
        GPHM_TRY(launch_transpose(Bt, n1, n2, p.Tf, st));                       // Bt^T
        GPHM_TRY(d1(Gt, anti ? -c1 : c1, 0.5, nullptr, p.Tf));                  // (c1 D1^T G + Bt/2)^T
        GPHM_TRY(gs1(p.Tf, p.W)); V1t = p.W;                                    // V1^T
        if (gs_refine_axis(p, X1)) GPHM_TRY(refine_kinv_rows_gs(X1, p.Tf, n2, p.W, p.Tf, st));
        GPHM_TRY(launch_transpose(p.W, n2, n1, p.V1, st));
        GPHM_TRY(d2(G, anti ? -1.0 : 1.0, 0.5, nullptr, p.A));                  // G D2 + A/2  (A is free after the residual)
        GPHM_TRY(gs2(p.A, p.V2)); V2 = p.V2;
        if (gs_refine_axis(p, X2)) GPHM_TRY(refine_kinv_rows_gs(X2, p.A, n1, p.V2, p.A, st));
        if (!p.defer_gu)           // gphm_step with look-ahead assembles dL/dU after the theta-gradient (grad_u_2d_gs below)
            GPHM_TRY(launch_grad_u(lc, p.has_base ? p.base : nullptr, U, G, p.V1, p.V2, nullptr, p.eb, p.xind, small, gU, nullptr,
                                   nullptr, st));
    } else {
        GPHM_TRY(d1(G, anti ? -c1 : c1, 0.5, U, p.Tf));                         // D^T g + u/2
        GPHM_TRY(gs1(p.Tf, p.V1)); V1t = p.V1;                                  // s + a/2
        if (gs_refine_axis(p, X1)) GPHM_TRY(refine_kinv_rows_gs(X1, p.Tf, n2, p.V1, p.Tf, st));
        GPHM_TRY(launch_lincomb(p.S1, 0.5, At, 0.0, nullptr, nf, st));
        GPHM_TRY(launch_grad_u(lc, p.has_base ? p.base : nullptr, U, G, p.V1, p.S1, nullptr, p.eb, p.xind, small, gU, nullptr,
                               nullptr, st));
    }
    if (p.on_gu) { GPHM_TRY(p.on_gu(p, st)); p.gu_hook_ran = true; }        // nothing below reads U or gU
    // diagonal sums of Kbar_a = ld/2 N_b K_a^-1 - V_a (.)^T and Dbar_a by row cross-correlations
    // (one transform per row PAIR against the stored transforms of A^T / Bt)
    // (the single-CTA tails - partial-spectrum reduction, inverse transform, theta contraction - of both axes share launches)
    GPHM_TRY(launch_xcorr_pairs(V1t, n2, n1, n1, X1.specY, X1.fftL, X1.twid, -1.0, X1.specK, st));
    GPHM_TRY(launch_xcorr_pairs(Gt, n2, n1, n1, X1.specY, X1.fftL, X1.twid, c1, X1.specD, st));
    DiagSumsJob dj[2];
    dj[0] = DiagSumsJob{X1.specK, X1.specD, X1.twid, n1, anti, X1.dirsign, X1.sKinv, 0.5 * d.logdet * n2, X1.sK, X1.sD};
    const bool paired = two && X1.fftL == X2.fftL;
    if (!paired) GPHM_TRY(launch_spectrum_to_diag_sums_multi(dj, 1, X1.fftL, st));
    if (two) {
        GPHM_TRY(launch_xcorr_pairs(V2, n1, n2, n2, X2.specY, X2.fftL, X2.twid, -1.0, X2.specK, st));
        GPHM_TRY(launch_xcorr_pairs(G, n1, n2, n2, X2.specY, X2.fftL, X2.twid, 1.0, X2.specD, st));
        dj[1] = DiagSumsJob{X2.specK, X2.specD, X2.twid, n2, anti, X2.dirsign, X2.sKinv, 0.5 * d.logdet * n1, X2.sK, X2.sD};
        if (paired) GPHM_TRY(launch_spectrum_to_diag_sums_multi(dj, 2, X1.fftL, st));
        else GPHM_TRY(launch_spectrum_to_diag_sums_multi(dj + 1, 1, X2.fftL, st));
    }
    ThetaGradJob tj[2];
    for (int a = 0; a < (two ? 2 : 1); ++a) {
        Axis& X = p.ax[a];
        tj[a] = ThetaGradJob{X.x, X.n, theta_of(p, small, a), X.sK, X.sD, gsmall + (size_t)a * 3 * Q};
    }
    GPHM_TRY(launch_theta_grad_toeplitz_multi(d.kernel_id, order, tj, two ? 2 : 1, Q, st));
    if (!two) GPHM_CUDA_OK(cudaMemsetAsync(gsmall + 3 * Q, 0, sizeof(double) * 3 * Q, st));
    return GPHM_OK;
}

// the deferred dL/dU assembly of logjoint_grad_gs (2-D): same launch, issued by gphm_step after the theta-gradient
int grad_u_2d_gs(gphm_plan& p, const double* U, const double* small, double* gU, cudaStream_t st) {
    return launch_grad_u(loss_consts(p), p.has_base ? p.base : nullptr, U, p.R, p.V1, p.V2, nullptr, p.eb, p.xind, small, gU, nullptr,
                         nullptr, st);
}

bool uses_gs_path(const gphm_plan& p) {
    const bool two = p.d.dim == 2;
    return p.ax[0].gs && (!two || p.ax[1].gs) && !(p.d.force_general & 8) && toeplitz_fused_supported(p.ax[0].fftL) &&
           (!two || toeplitz_fused_supported(p.ax[1].fftL));
}

int logjoint_grad(gphm_plan& p, const double* U, const double* small, double* gU, double* gsmall, double* terms,
                  int flags, cudaStream_t st) {
    const gphm_problem_desc& d = p.d;
    const bool two = d.dim == 2, fwd_only = (flags & GPHM_FORWARD_ONLY) != 0;
    if (p.ax[0].gs && (!two || p.ax[1].gs) && !(d.force_general & 8) && toeplitz_fused_supported(p.ax[0].fftL) &&
        (!two || toeplitz_fused_supported(p.ax[1].fftL)))
        return logjoint_grad_gs(p, U, small, gU, gsmall, terms, flags, st);
    const int n1 = d.n1, n2 = d.n2, Q = d.Q;
    const size_t nf = (size_t)n1 * n2;
    const double c1 = coef_c1(p);
    Axis& X1 = p.ax[0];
    Axis& X2 = p.ax[1];
    if (p.u_ready) GPHM_CUDA_OK(cudaStreamWaitEvent(st, p.u_ready, 0));

    const bool fft_kinv = (d.force_general & 4) == 0;   // then K^-1 itself is never formed on FFT axes
    const bool kinv0 = !fwd_only && !(fft_kinv && X1.fftL > 0), kinv1 = !fwd_only && !(fft_kinv && X2.fftL > 0);
    if (two) GPHM_TRY(factor_both(p, small, kinv0, kinv1, st));
    else GPHM_TRY(factor_axis(p, 0, small, kinv0, st));

    // ---- forward ----
    GPHM_TRY(apply_kinv(p, 0, 0, U, n1, n2, p.A, p.Tf, st));                                   // A = K1^-1 U
    const double* Bt = U;
    if (two) { GPHM_TRY(apply_kinv(p, 1, 1, U, n1, n2, p.Bt, p.Tf, st)); Bt = p.Bt; }           // Bt = U K2^-1
    const bool anti = deriv_order(p) == 1;                 // advection: D is antisymmetric, D^T = -D
    const bool tfft1 = X1.fftL > 0 && !(d.force_general & 8), tfft2 = two && X2.fftL > 0 && !(d.force_general & 8);
    if (tfft1) {   // uniform grid: D1 is Toeplitz -> circulant convolution of every column of A (rows of A^T)
        GPHM_TRY(launch_toeplitz_spectrum(X1.tabD, n1, X1.fftL, X1.twid, anti, X1.dirsign, X1.specT, st));
        GPHM_TRY(launch_transpose(p.A, n1, n2, p.Tf, st));
        GPHM_TRY(launch_toeplitz_apply(p.Tf, n2, n1, n1, X1.specT, X1.fftL, X1.twid, c1, 0.0, p.P, n1, st));
        GPHM_TRY(launch_transpose(p.P, n2, n1, p.R, st));
    } else {
        GPHM_TRY(contract(p, gemm_args(X1.D, n1, false, p.A, n2, false, p.R, n2, n1, n2, n1, c1, 0.0), st));   // c1 D1 A
    }
    if (two) {
        if (tfft2) {   // (Bt D2^T)[i,:] = D2 Bt[i,:]
            GPHM_TRY(launch_toeplitz_spectrum(X2.tabD, n2, X2.fftL, X2.twid, anti, X2.dirsign, X2.specT, st));
            GPHM_TRY(launch_toeplitz_apply(Bt, n1, n2, n2, X2.specT, X2.fftL, X2.twid, 1.0, 1.0, p.R, n2, st));
        } else {
            GPHM_TRY(contract(p, gemm_args(Bt, n2, false, X2.D, n2, true, p.R, n2, n1, n2, n2, 1.0, 1.0), st)); // + Bt D2^T
        }
    }
    GPHM_TRY(launch_residual(p.R, U, p.src, p.A, Bt, nf, d.eq_type, p.has_base ? p.base : nullptr, small, Q, p.part, st));    // R <- G
    const LossConsts lc = loss_consts(p);
    GPHM_TRY(launch_finalize(lc, U, p.bvals, p.xind, p.part, X1.ldpart, X1.nblk, two ? X2.ldpart : nullptr,
                             two ? X2.nblk : 0, small, p.eb, terms, fwd_only ? nullptr : gsmall, p.status, st));
    if (fwd_only) return GPHM_OK;

    // ---- backward ----
    const double* G = p.R;
    const double* W = p.A;
    if (two) { GPHM_TRY(apply_kinv(p, 0, 0, Bt, n1, n2, p.W, p.Tf, st)); W = p.W; }             // W = K1^-1 Bt
    if (tfft1) {   // c1 D1^T G: D1^T = +-D1 applied to the columns of G
        GPHM_TRY(launch_transpose(G, n1, n2, p.Tf, st));
        GPHM_TRY(launch_toeplitz_apply(p.Tf, n2, n1, n1, X1.specT, X1.fftL, X1.twid, anti ? -c1 : c1, 0.0, p.V1, n1, st));
        GPHM_TRY(launch_transpose(p.V1, n2, n1, p.P, st));
    } else {
        GPHM_TRY(contract(p, gemm_args(X1.D, n1, true, G, n2, false, p.P, n2, n1, n2, n1, c1, 0.0), st));  // c1 D1^T G
    }
    GPHM_TRY(apply_kinv(p, 0, 0, p.P, n1, n2, p.S1, p.Tf, st));                                 // S1
    if (two) {
        if (tfft2)     // (G D2)[i,:] = D2^T G[i,:]
            GPHM_TRY(launch_toeplitz_apply(G, n1, n2, n2, X2.specT, X2.fftL, X2.twid, anti ? -1.0 : 1.0, 0.0, p.P, n2, st));
        else
            GPHM_TRY(contract(p, gemm_args(G, n2, false, X2.D, n2, false, p.P, n2, n1, n2, n2, 1.0, 0.0), st)); // G D2
        GPHM_TRY(apply_kinv(p, 1, 1, p.P, n1, n2, p.S2, p.Tf, st));                             // S2
    }
    GPHM_TRY(launch_grad_u(lc, p.has_base ? p.base : nullptr, U, G, W, p.S1, two ? p.S2 : nullptr, p.eb, p.xind, small, gU, p.V1,
                           two ? p.V2 : nullptr, st));
    const int order = deriv_order(p);
    // ---- axis 1: Kbar1 = ld/2*N2*K1^-1 - V1 A^T,  Dbar1 = c1 G A^T ----
    if (X1.fftL > 0) {
        // uniform grid: only the diagonal sums are needed -> FFT cross-correlations of the columns
        // (rows after a transpose) instead of the two GEMMs and K1^-1
        double *V1t = p.P, *At = p.Tf, *Gt = p.S1;            // free scratch at this point
        GPHM_TRY(launch_transpose(p.V1, n1, n2, V1t, st));
        GPHM_TRY(launch_transpose(p.A, n1, n2, At, st));
        GPHM_TRY(launch_transpose(G, n1, n2, Gt, st));
        const bool fk = (d.force_general & 4) == 0 && !X1.gs;   // K^-1 sums as FFT autocorrelation of the rows of Linv
        if (X1.gs) {}                                           // sKinv already holds them (factor_gs)
        else if (fk) GPHM_TRY(launch_xcorr_spectrum(X1.Linv, X1.Linv, n1, n1, n1, n1, X1.fftL, X1.twid, 0.5 * d.logdet * n2, false, X1.specK, st));
        else GPHM_TRY(launch_diag_sums(X1.Kinv, nullptr, n1, n1, false, 1.0, X1.dspart, X1.sKinv, nullptr, st));
        GPHM_TRY(launch_xcorr_spectrum(V1t, At, n2, n1, n1, n1, X1.fftL, X1.twid, -1.0, fk, X1.specK, st));
        GPHM_TRY(launch_xcorr_spectrum(Gt, At, n2, n1, n1, n1, X1.fftL, X1.twid, c1, false, X1.specD, st));
        GPHM_TRY(launch_spectrum_to_diag_sums(X1.specK, X1.specD, X1.fftL, X1.twid, n1, order == 1, X1.dirsign,
                                              fk ? nullptr : X1.sKinv, 0.5 * d.logdet * n2, X1.sK, X1.sD, st));
    } else {
        GPHM_TRY(contract(p, gemm_args(p.V1, n2, false, p.A, n2, true, X1.Kinv, n1, n1, n1, n2, -1.0,
                                        0.5 * d.logdet * n2), st));
        GPHM_TRY(contract(p, gemm_args(G, n2, false, p.A, n2, true, X1.Dbar, n1, n1, n1, n2, c1, 0.0), st));
        if (X1.toeplitz) GPHM_TRY(launch_diag_sums(X1.Kinv, X1.Dbar, n1, n1, order == 1, X1.dirsign, X1.dspart, X1.sK, X1.sD, st));
    }
    // ---- axis 2: Kbar2 = ld/2*N1*K2^-1 - V2^T Bt,  Dbar2 = G^T Bt ----
    if (two) {
        if (X2.fftL > 0) {
            const bool fk = (d.force_general & 4) == 0 && !X2.gs;
            if (X2.gs) {}
            else if (fk) GPHM_TRY(launch_xcorr_spectrum(X2.Linv, X2.Linv, n2, n2, n2, n2, X2.fftL, X2.twid, 0.5 * d.logdet * n1, false, X2.specK, st));
            else GPHM_TRY(launch_diag_sums(X2.Kinv, nullptr, n2, n2, false, 1.0, X2.dspart, X2.sKinv, nullptr, st));
            GPHM_TRY(launch_xcorr_spectrum(p.V2, Bt, n1, n2, n2, n2, X2.fftL, X2.twid, -1.0, fk, X2.specK, st));
            GPHM_TRY(launch_xcorr_spectrum(G, Bt, n1, n2, n2, n2, X2.fftL, X2.twid, 1.0, false, X2.specD, st));
            GPHM_TRY(launch_spectrum_to_diag_sums(X2.specK, X2.specD, X2.fftL, X2.twid, n2, order == 1, X2.dirsign,
                                                  fk ? nullptr : X2.sKinv, 0.5 * d.logdet * n1, X2.sK, X2.sD, st));
        } else {
            GPHM_TRY(contract(p, gemm_args(p.V2, n2, true, Bt, n2, false, X2.Kinv, n2, n2, n2, n1, -1.0,
                                            0.5 * d.logdet * n1), st));
            GPHM_TRY(contract(p, gemm_args(G, n2, true, Bt, n2, false, X2.Dbar, n2, n2, n2, n1, 1.0, 0.0), st));
            if (X2.toeplitz) GPHM_TRY(launch_diag_sums(X2.Kinv, X2.Dbar, n2, n2, order == 1, X2.dirsign, X2.dspart, X2.sK, X2.sD, st));
        }
    }
    for (int a = 0; a < (two ? 2 : 1); ++a) {
        Axis& X = p.ax[a];
        const double* th = theta_of(p, small, a);
        double* gth = gsmall + (size_t)a * 3 * Q;
        if (X.toeplitz)
            GPHM_TRY(launch_theta_grad_toeplitz(d.kernel_id, order, X.x, X.n, th, Q, X.sK, X.sD, gth, st));
        else
            GPHM_TRY(launch_theta_grad_general(d.kernel_id, order, X.x, X.n, th, Q, X.Kinv, X.Dbar, X.n, X.tgpart, gth, st));
    }
    if (!two) GPHM_CUDA_OK(cudaMemsetAsync(gsmall + 3 * Q, 0, sizeof(double) * 3 * Q, st));
    return GPHM_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int gphm_version(void) { return GPHM_VERSION; }
const char* gphm_last_error(void) { return g_err; }

long long gphm_launch_count(void) { return g_launches.load(); }

int gphm_profile_start(void) {
    for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
    g_recs.clear();
    g_prof = true;
    return GPHM_OK;
}

int gphm_profile_stop(double* ms, double* flops, double* bytes, long long* launches) {
    g_prof = false;
    GPHM_CUDA_OK(cudaDeviceSynchronize());
    for (int c = 0; c < CAT_COUNT; ++c) { if (ms) ms[c] = 0; if (flops) flops[c] = 0; if (bytes) bytes[c] = 0; if (launches) launches[c] = 0; }
    for (auto& r : g_recs) {
        float t = 0.f;
        GPHM_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
        if (ms) ms[r.cat] += t;
        if (flops) flops[r.cat] += r.flops;
        if (bytes) bytes[r.cat] += r.bytes;
        if (launches) launches[r.cat] += 1;
        g_pool.push_back(r.a); g_pool.push_back(r.b);
    }
    g_recs.clear();
    return GPHM_OK;
}

int gphm_gram(int kernel_id, int deriv_order, const double* d_x1, int n1, const double* d_x2, int n2,
              const double* d_theta, int Q, double jitter, double* d_out, void* stream) {
    if (!d_x1 || !d_x2 || !d_theta || !d_out) { set_last_error("gphm_gram: null pointer"); return GPHM_EINVAL; }
    if (n1 < 0 || n2 < 0) { set_last_error("gphm_gram: negative size"); return GPHM_EINVAL; }
    if (kernel_id < 0 || kernel_id > 3) { set_last_error("Invalid Kernel id %d", kernel_id); return GPHM_EINVAL; }
    if (deriv_order < 0 || deriv_order > 2) { set_last_error("gphm_gram: deriv_order %d", deriv_order); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double jit = (n1 == n2) ? jitter : 0.0;
    if (deriv_order == 0)
        return launch_gram_general(kernel_id, 0, d_x1, n1, d_x2, n2, d_theta, Q, jit, d_out, nullptr, n2, st);
    return launch_gram_general(kernel_id, deriv_order, d_x1, n1, d_x2, n2, d_theta, Q, 0.0, nullptr, d_out, n2, st);
}

int gphm_kappa_pairs(int kernel_id, int deriv_order, const double* d_x1, const double* d_x2, size_t npairs,
                     const double* d_theta, int Q, double* d_out, void* stream) {
    if (npairs == 0) return GPHM_OK;
    if (!d_x1 || !d_x2 || !d_theta || !d_out) { set_last_error("gphm_kappa_pairs: null pointer"); return GPHM_EINVAL; }
    if (kernel_id < 0 || kernel_id > 3) { set_last_error("Invalid Kernel id %d", kernel_id); return GPHM_EINVAL; }
    if (deriv_order < 0 || deriv_order > 2) { set_last_error("gphm_kappa_pairs: deriv_order %d", deriv_order); return GPHM_EINVAL; }
    return launch_kappa_pairs(kernel_id, deriv_order, d_x1, d_x2, npairs, d_theta, Q, d_out, static_cast<cudaStream_t>(stream));
}

int gphm_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* d_A, int lda,
               const double* d_B, int ldb, double beta, double* d_C, int ldc, void* stream) {
    if (M < 0 || N < 0 || K < 0) { set_last_error("gphm_dgemm: negative size"); return GPHM_EINVAL; }
    if (M == 0 || N == 0) return GPHM_OK;
    if (!d_C || (K > 0 && (!d_A || !d_B))) { set_last_error("gphm_dgemm: null pointer"); return GPHM_EINVAL; }
    if (K == 0) { d_A = d_C; d_B = d_C; }          // C = beta*C; operands are never dereferenced
    return launch_dgemm(gemm_args(d_A, lda, transA != 0, d_B, ldb, transB != 0, d_C, ldc, M, N, K, alpha, beta, 0),
                        static_cast<cudaStream_t>(stream));
}

size_t gphm_ozaki_work_bytes(int M, int N, int K, int slices) {
    return ozaki_work_bytes(M, N, K, slices > 0 ? slices : ozaki_default_slices());
}

double gphm_ozaki_error_factor(int K, int slices) { return ozaki_error_factor(K, slices > 0 ? slices : ozaki_default_slices()); }

int gphm_ozaki_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double* d_A, int lda, const double* d_B,
                     int ldb, double beta, double* d_C, int ldc, int slices, void* d_work, size_t work_bytes, void* stream) {
    if (M < 0 || N < 0 || K <= 0) { set_last_error("gphm_ozaki_dgemm: bad size"); return GPHM_EINVAL; }
    if (M == 0 || N == 0) return GPHM_OK;
    if (!d_A || !d_B || !d_C || !d_work) { set_last_error("gphm_ozaki_dgemm: null pointer"); return GPHM_EINVAL; }
    return launch_ozaki_dgemm(transA != 0, transB != 0, M, N, K, alpha, d_A, lda, d_B, ldb, beta, d_C, ldc,
                              slices > 0 ? slices : ozaki_default_slices(), d_work, work_bytes, static_cast<cudaStream_t>(stream));
}

size_t gphm_potrf_work_bytes(int n) {
    if (n <= 0) return 0;
    Carver c(nullptr);
    double* p;
    c.take(p, (size_t)num_blocks_nb(n) * kNB * kNB); c.take(p, num_blocks_nb(n)); c.take(p, (size_t)n * n);
    return c.off;
}

int gphm_potrf_inv(double* d_K, int n, double* d_L, double* d_Linv, double* d_logdet, int* d_status, void* d_work,
                   void* stream) {
    if (n <= 0) { set_last_error("gphm_potrf_inv: n <= 0"); return GPHM_EINVAL; }
    if (!d_K || !d_L || !d_Linv || !d_logdet || !d_status || !d_work) { set_last_error("gphm_potrf_inv: null pointer"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(d_work);
    double *invdiag, *ldpart, *T;
    const int nblk = num_blocks_nb(n);
    c.take(invdiag, (size_t)nblk * kNB * kNB); c.take(ldpart, nblk); c.take(T, (size_t)n * n);
    GPHM_CUDA_OK(cudaMemsetAsync(d_L, 0, sizeof(double) * (size_t)n * n, st));
    GPHM_CUDA_OK(cudaMemsetAsync(d_Linv, 0, sizeof(double) * (size_t)n * n, st));
    GPHM_CUDA_OK(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    GPHM_TRY(chol_factor(d_K, d_L, n, n, invdiag, ldpart, d_status, st));
    GPHM_TRY(trtri_lower(d_L, d_Linv, n, n, invdiag, T, st));
    GPHM_TRY(launch_sum_scaled(ldpart, nblk, 2.0, d_logdet, st));
    return GPHM_OK;
}

size_t gphm_toeplitz_work_bytes(int n, int rows) {
    const int L = n > 0 ? fft_length_for(n) : 0;
    if (n <= 0 || rows < 0 || L == 0 || n > toeplitz_inv_max_n()) return 0;
    Carver c(nullptr);
    double* p;
    c.take(p, 2 * (size_t)L); c.take(p, 8 * (size_t)L); c.take(p, 1); c.take(p, (size_t)std::max(rows, 1) * n);
    c.take(p, (size_t)n); c.take(p, 8); c.take(p, 6 * (size_t)n);
    return c.off;
}

int gphm_toeplitz_solve(const double* d_t, int n, const double* d_B, int rows, double* d_X, double* d_g, double* d_sKinv,
                        double* d_logdet, int* d_status, void* d_work, void* stream) {
    if (!d_t || !d_g || !d_sKinv || !d_logdet || !d_status || !d_work || (rows > 0 && (!d_B || !d_X))) { set_last_error("gphm_toeplitz_solve: null pointer"); return GPHM_EINVAL; }
    const int L = n > 0 ? fft_length_for(n) : 0;
    if (n <= 0 || rows < 0 || L == 0 || n > toeplitz_inv_max_n()) { set_last_error("gphm_toeplitz_solve: n=%d outside [1,%d]", n, std::min(toeplitz_inv_max_n(), 4096)); return GPHM_EINVAL; }
    if (rows > 0 && d_B == d_X) { set_last_error("gphm_toeplitz_solve: d_X may not alias d_B"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(d_work);
    double *twid, *spec, *hld, *tmp, *gkap, *progd, *gbnd;
    c.take(twid, 2 * (size_t)L); c.take(spec, 8 * (size_t)L); c.take(hld, 1); c.take(tmp, (size_t)std::max(rows, 1) * n);
    c.take(gkap, (size_t)n); c.take(progd, 8); c.take(gbnd, 6 * (size_t)n);
    GPHM_CUDA_OK(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    GPHM_TRY(launch_twiddle_init(twid, L, st));
    int* guardp = reinterpret_cast<int*>(progd) + 8;           // progd: 8 progress ints, then the guard word
    GPHM_CUDA_OK(cudaMemsetAsync(guardp, 0, sizeof(int), st));
    GPHM_TRY(launch_schur_levinson(d_t, 0, n, 0.0, d_g, 0, hld, 0, d_status, 0, gkap, 0, reinterpret_cast<int*>(progd), 0, 1, st,
                                   getenv("GPHM_SCHUR_CYCLES") ? reinterpret_cast<long long*>(tmp) : nullptr, guardp, 0, gbnd, 0));
    GPHM_TRY(launch_status_merge_guard(d_status, guardp, st));
    GPHM_TRY(launch_gs_prepare(d_g, 0, n, L, twid, spec, 0, d_sKinv, 0, 1, st));
    GPHM_TRY(launch_sum_scaled(hld, 1, 2.0, d_logdet, st));
    if (rows > 0) {
        Axis X;
        X.n = n; X.fftL = L; X.twid = twid; X.gspec = spec; X.gsg = d_g;
        GPHM_TRY(apply_kinv_rows_gs(X, d_B, rows, d_X, tmp, st));
    }
    return GPHM_OK;
}

size_t gphm_workspace_bytes(const gphm_problem_desc* desc) {
    if (check_desc(desc) != GPHM_OK) return 0;
    gphm_plan tmp;
    tmp.d = *desc;
    tmp.size_query = true;     // uniform or not is only known at create time: size for either
    init_axes(tmp, nullptr, nullptr);
    return carve(tmp, nullptr);
}

int gphm_plan_create(const gphm_problem_desc* desc, const double* h_x, const double* h_y, const double* h_src,
                     const double* h_bvals, const int* h_xind, void* d_workspace, size_t workspace_bytes,
                     gphm_plan** out) {
    GPHM_TRY(check_desc(desc));
    if (!out || !h_x || !h_src || (desc->nb > 0 && !h_bvals)) { set_last_error("gphm_plan_create: null pointer"); return GPHM_EINVAL; }
    if (desc->dim == 2 && !h_y) { set_last_error("gphm_plan_create: h_y required for dim == 2"); return GPHM_EINVAL; }
    if (desc->dim == 1 && desc->nb > 0 && !h_xind) { set_last_error("gphm_plan_create: h_xind required for dim == 1"); return GPHM_EINVAL; }
    if (desc->dim == 1)
        for (int e = 0; e < desc->nb; ++e)
            if (h_xind[e] < 0 || h_xind[e] >= desc->n1) { set_last_error("Xind[%d]=%d out of range", e, h_xind[e]); return GPHM_EINVAL; }
    GPHM_TRY(dgemm_init());
    GPHM_TRY(factor_init());
    gphm_plan* p = new gphm_plan();
    p->d = *desc;
    init_axes(*p, h_x, h_y);
    const size_t need = carve(*p, nullptr);
    if (d_workspace) {
        if (workspace_bytes < need) {
            set_last_error("workspace too small: %zu < %zu", workspace_bytes, need);
            delete p;
            return GPHM_ENOMEM;
        }
        p->ws = d_workspace;
    } else {
        if (cudaMalloc(&p->ws, need) != cudaSuccess) {
            set_last_error("cudaMalloc(%zu) failed", need);
            delete p;
            return GPHM_ENOMEM;
        }
        p->owns_ws = true;
    }
    carve(*p, p->ws);
    auto fail = [&](const char* what) { set_last_error("gphm_plan_create: %s failed", what); gphm_plan_destroy(p); return GPHM_ECUDA; };
    const size_t nf = (size_t)desc->n1 * desc->n2;
    if (cudaMemset(p->ws, 0, need) != cudaSuccess) return fail("memset");
    if (cudaMemcpy(p->ax[0].x, h_x, sizeof(double) * desc->n1, cudaMemcpyHostToDevice) != cudaSuccess) return fail("copy x");
    if (desc->dim == 2 && cudaMemcpy(p->ax[1].x, h_y, sizeof(double) * desc->n2, cudaMemcpyHostToDevice) != cudaSuccess) return fail("copy y");
    if (cudaMemcpy(p->src, h_src, sizeof(double) * nf, cudaMemcpyHostToDevice) != cudaSuccess) return fail("copy src");
    if (desc->nb > 0 && cudaMemcpy(p->bvals, h_bvals, sizeof(double) * desc->nb, cudaMemcpyHostToDevice) != cudaSuccess) return fail("copy bvals");
    if (desc->dim == 1 && desc->nb > 0 && cudaMemcpy(p->xind, h_xind, sizeof(int) * desc->nb, cudaMemcpyHostToDevice) != cudaSuccess) return fail("copy xind");
    for (int a = 0; a < 2; ++a)
        if (p->ax[a].fftL > 0) {
            if (launch_twiddle_init(p->ax[a].twid, p->ax[a].fftL, nullptr) != GPHM_OK) return fail("twiddle init");
        }
    if (cudaDeviceSynchronize() != cudaSuccess) return fail("synchronize");
    *out = p;
    return GPHM_OK;
}

void gphm_plan_destroy(gphm_plan* plan) {
    if (!plan) return;
    if (plan->owns_ws && plan->ws) cudaFree(plan->ws);
    if (plan->oz_ws) cudaFree(plan->oz_ws);
    if (plan->lk_small) cudaFree(plan->lk_small);
    if (plan->lk_flags) cudaFree(plan->lk_flags);
    if (plan->lk_stream) cudaStreamDestroy(plan->lk_stream);
    if (plan->lk_fork) cudaEventDestroy(plan->lk_fork);
    if (plan->lk_join) cudaEventDestroy(plan->lk_join);
    if (plan->hs) cudaFree(plan->hs);
    if (plan->hs_count) cudaFree(plan->hs_count);
    if (plan->hs_stream) cudaStreamDestroy(plan->hs_stream);
    if (plan->hs_ev_u) cudaEventDestroy(plan->hs_ev_u);
    for (cudaEvent_t e : plan->hs_ev_chunk) if (e) cudaEventDestroy(e);
    if (plan->hs_ev_mv) cudaEventDestroy(plan->hs_ev_mv);
    if (plan->hs_ev_gu) cudaEventDestroy(plan->hs_ev_gu);
    if (plan->hs_ev_adam) cudaEventDestroy(plan->hs_ev_adam);
    delete plan;
}

int gphm_plan_status(gphm_plan* plan, int* pivot, void* stream) {
    if (!plan) { set_last_error("null plan"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int h[4] = {0, 0, 0, 0};
    GPHM_CUDA_OK(cudaMemcpyAsync(h, plan->status, sizeof(h), cudaMemcpyDeviceToHost, st));
    GPHM_CUDA_OK(cudaStreamSynchronize(st));
    GPHM_CUDA_OK(cudaMemsetAsync(plan->status, 0, sizeof(h), st));
    if (pivot) *pivot = h[0] ? h[0] : (h[1] ? plan->d.n1 + h[1] : 0);
    if (h[0] || h[1]) return GPHM_NOT_SPD;
    if (h[2]) return GPHM_NONFINITE;
    if (h[3] & 0xff00) { set_last_error("Schur/Levinson lattice CTA never received its coefficients (generator CTA not resident)"); return GPHM_STALLED; }
    if (h[3]) { if (pivot) *pivot = h[3]; return GPHM_ILL_CONDITIONED; }
    return GPHM_OK;
}

int gphm_plan_use_cholesky(gphm_plan* plan) {
    if (!plan) { set_last_error("null plan"); return GPHM_EINVAL; }
    plan->d.force_general |= 16;
    for (int a = 0; a < 2; ++a) plan->ax[a].gs = false;
    return GPHM_OK;
}

int gphm_plan_set_base_field(gphm_plan* plan, const double* h_base) {
    if (!plan) { set_last_error("null plan"); return GPHM_EINVAL; }
    plan->has_base = h_base != nullptr;
    if (h_base)
        GPHM_CUDA_OK(cudaMemcpy(plan->base, h_base, sizeof(double) * (size_t)plan->d.n1 * plan->d.n2, cudaMemcpyHostToDevice));
    return GPHM_OK;
}

int gphm_plan_uses_toeplitz(const gphm_plan* plan, int axis) {
    if (!plan || axis < 0 || axis > 1) return 0;
    return plan->ax[axis].n > 0 && plan->ax[axis].toeplitz ? 1 : 0;
}

int gphm_logjoint_grad(gphm_plan* plan, const double* d_U, const double* d_small, double* d_gU, double* d_gsmall,
                       double* d_terms, int flags, void* stream) {
    if (!plan || !d_U || !d_small || !d_terms) { set_last_error("gphm_logjoint_grad: null pointer"); return GPHM_EINVAL; }
    if (!(flags & GPHM_FORWARD_ONLY) && (!d_gU || !d_gsmall)) { set_last_error("gphm_logjoint_grad: null gradient buffer"); return GPHM_EINVAL; }
    return logjoint_grad(*plan, d_U, d_small, d_gU, d_gsmall, d_terms, flags, static_cast<cudaStream_t>(stream));
}

int gphm_adam_update(double* d_p, const double* d_g, double* d_m, double* d_v, size_t n, const long long* d_count,
                     double lr, void* stream) {
    if (n == 0) return GPHM_OK;
    if (!d_p || !d_g || !d_m || !d_v || !d_count) { set_last_error("gphm_adam_update: null pointer"); return GPHM_EINVAL; }
    return launch_adam(d_p, d_g, d_m, d_v, n, d_count, lr, static_cast<cudaStream_t>(stream));
}

int gphm_adam_update_inc(double* d_p, const double* d_g, double* d_m, double* d_v, size_t n, long long* d_count, double lr,
                         void* stream) {
    if (!d_p || !d_g || !d_m || !d_v || !d_count) { set_last_error("gphm_adam_update_inc: null pointer"); return GPHM_EINVAL; }
    if (n > (1u << 20)) { set_last_error("gphm_adam_update_inc: meant for the short parameter vector (n <= 2^20)"); return GPHM_EINVAL; }
    return launch_adam_inc(d_p, d_g, d_m, d_v, n, d_count, lr, static_cast<cudaStream_t>(stream));
}

int gphm_step(gphm_plan* plan, double* d_U, double* d_small, double* d_mU, double* d_vU, double* d_msmall,
              double* d_vsmall, long long* d_count, double lr, double* d_terms, void* stream) {
    if (!plan || !d_U || !d_small || !d_mU || !d_vU || !d_msmall || !d_vsmall || !d_count || !d_terms) {
        set_last_error("gphm_step: null pointer");
        return GPHM_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    gphm_plan& p = *plan;
    const size_t nf = (size_t)p.d.n1 * p.d.n2, ns = 6 * (size_t)p.d.Q + 2;
    // Look-ahead (2-D all-FFT plans with axes of >= 2048 points): the factor stage of the NEXT step needs only the updated theta,
    // which exists as soon as the theta-gradient does.  OFF by default (GPHM_LOOKAHEAD=1 or force_general bit 9 switch it on, bit 8
    // off): measured on the 4096^2 step it does not pay - 7.41 ms with it against 7.35 ms without (high-priority side stream
    // included): the 0.5 ms chain of small dependent kernels does not finish inside the 0.25 ms of dL/dU assembly + Adam(U) it runs
    // beside, and 13 more launches per step eat the rest.  Kept because it is bit-exact and tested.
    static const bool lk_env = [] { const char* e = getenv("GPHM_LOOKAHEAD"); return e && e[0] == '1'; }();
    const bool lk = (lk_env || (p.d.force_general & 512)) && !(p.d.force_general & 256) && p.d.dim == 2 && uses_gs_path(p) &&
                    std::min(p.d.n1, p.d.n2) >= 2048;
    if (lk && !p.lk_stream) {
        int prio_lo = 0, prio_hi = 0;          // highest priority: the 16 CTAs of the recursion must not queue behind the wide element-wise kernels
        GPHM_CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        GPHM_CUDA_OK(cudaStreamCreateWithPriority(&p.lk_stream, cudaStreamNonBlocking, prio_hi));
        GPHM_CUDA_OK(cudaEventCreateWithFlags(&p.lk_fork, cudaEventDisableTiming));
        GPHM_CUDA_OK(cudaEventCreateWithFlags(&p.lk_join, cudaEventDisableTiming));
        GPHM_CUDA_OK(cudaMalloc(&p.lk_small, sizeof(double) * ns));
        GPHM_CUDA_OK(cudaMalloc(&p.lk_flags, sizeof(int) * 2));
        GPHM_CUDA_OK(cudaMemsetAsync(p.lk_flags, 0, sizeof(int) * 2, st));
        GPHM_CUDA_OK(cudaMemsetAsync(p.lk_small, 0xff, sizeof(double) * ns, st));
    }
    p.lk_enabled = lk;
    if (!lk) {
        GPHM_TRY(logjoint_grad(p, d_U, d_small, p.gU, p.gsmall, d_terms, 0, st));
        GPHM_TRY(launch_adam(d_U, p.gU, d_mU, d_vU, nf, d_count, lr, st));
        GPHM_TRY(launch_adam_inc(d_small, p.gsmall, d_msmall, d_vsmall, ns, d_count, lr, st));     // Adam(small) and ++count
        return GPHM_OK;
    }
    p.defer_gu = true;
    const int rc = logjoint_grad(p, d_U, d_small, p.gU, p.gsmall, d_terms, 0, st);                 // dL/dtheta complete, dL/dU pending
    p.defer_gu = false;
    GPHM_TRY(rc);
    GPHM_TRY(launch_adam_out(d_small, p.lk_small, p.gsmall, d_msmall, d_vsmall, ns, d_count, lr, st));      // theta of step t+1 -> lk_small
    GPHM_CUDA_OK(cudaEventRecord(p.lk_fork, st));
    GPHM_CUDA_OK(cudaStreamWaitEvent(p.lk_stream, p.lk_fork, 0));
    GPHM_TRY(factor_gs(p, 0, 2, p.lk_small, p.lk_stream));                                         // invalidates, recomputes ...
    GPHM_TRY(launch_lk_set(p.lk_flags, 1, p.lk_stream));                                           // ... and marks the result valid
    GPHM_CUDA_OK(cudaEventRecord(p.lk_join, p.lk_stream));
    GPHM_TRY(grad_u_2d_gs(p, d_U, d_small, p.gU, st));                                             // still the OLD theta (log_tau) in d_small
    GPHM_TRY(launch_adam(d_U, p.gU, d_mU, d_vU, nf, d_count, lr, st));
    GPHM_TRY(launch_copy(d_small, p.lk_small, ns, st));
    GPHM_TRY(launch_count_inc(d_count, st));
    GPHM_CUDA_OK(cudaStreamWaitEvent(st, p.lk_join, 0));
    return GPHM_OK;
}

// Shared body of gphm_step_host / gphm_step_host_params.  resident: the Adam moments and the step count live in the
// plan's staging area (the caller's opt_state never leaves the device, as in the reference where optax's state is a
// pytree of device arrays); only params travel.
static int step_host_impl(gphm_plan* plan, double* h_U, double* h_small, double* h_mU, double* h_vU, double* h_msmall,
                          double* h_vsmall, long long* h_count, double lr, double* h_terms, bool resident, bool reset_opt,
                          cudaStream_t st) {
    const size_t nf = (size_t)plan->d.n1 * plan->d.n2, ns = 6 * (size_t)plan->d.Q + 2;
    const size_t nfp = (nf + 31) / 32 * 32, nsp = (ns + 31) / 32 * 32;
    bool fresh = false;
    if (!plan->hs) {
        GPHM_CUDA_OK(cudaMalloc(&plan->hs, sizeof(double) * (3 * nfp + 3 * nsp + 32)));
        GPHM_CUDA_OK(cudaMalloc(&plan->hs_count, sizeof(long long)));
        GPHM_CUDA_OK(cudaStreamCreateWithFlags(&plan->hs_stream, cudaStreamNonBlocking));
        GPHM_CUDA_OK(cudaEventCreateWithFlags(&plan->hs_ev_u, cudaEventDisableTiming));
        for (cudaEvent_t& e : plan->hs_ev_chunk) GPHM_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        GPHM_CUDA_OK(cudaEventCreateWithFlags(&plan->hs_ev_mv, cudaEventDisableTiming));
        GPHM_CUDA_OK(cudaEventCreateWithFlags(&plan->hs_ev_gu, cudaEventDisableTiming));
        GPHM_CUDA_OK(cudaEventCreateWithFlags(&plan->hs_ev_adam, cudaEventDisableTiming));
        fresh = true;
    }
    double *U = plan->hs, *mU = U + nfp, *vU = mU + nfp, *sm = vU + nfp, *msm = sm + nsp, *vsm = msm + nsp,
           *terms = vsm + nsp;
    // Schedule (PCIe carries 134 MB up before and 134 (403 with the moments) MB down after the kernels at 4096^2):
    //   st        : small params up -> factor stage (needs only theta) -> [wait U] gradient ... theta-gradient -> Adam(small) -> down
    //   hs_stream : U up -> [mU, vU up, hidden behind the gradient] -> [wait dL/dU] Adam(U) -> U [, mU, vU] down while st
    //               still computes the theta-gradient
    GPHM_CUDA_OK(cudaMemcpyAsync(sm, h_small, sizeof(double) * ns, cudaMemcpyHostToDevice, st));
    if (!resident) {
        GPHM_CUDA_OK(cudaMemcpyAsync(msm, h_msmall, sizeof(double) * ns, cudaMemcpyHostToDevice, st));
        GPHM_CUDA_OK(cudaMemcpyAsync(vsm, h_vsmall, sizeof(double) * ns, cudaMemcpyHostToDevice, st));
        GPHM_CUDA_OK(cudaMemcpyAsync(plan->hs_count, h_count, sizeof(long long), cudaMemcpyHostToDevice, st));
    } else if (reset_opt || fresh) {                 // optimizer.init(params): zero moments, count 0
        GPHM_CUDA_OK(cudaMemsetAsync(mU, 0, sizeof(double) * 2 * nfp, st));
        GPHM_CUDA_OK(cudaMemsetAsync(msm, 0, sizeof(double) * 2 * nsp, st));
        GPHM_CUDA_OK(cudaMemsetAsync(plan->hs_count, 0, sizeof(long long), st));
        GPHM_CUDA_OK(cudaEventRecord(plan->hs_ev_mv, st));
        GPHM_CUDA_OK(cudaStreamWaitEvent(plan->hs_stream, plan->hs_ev_mv, 0));
    }
    // 2-D all-FFT plans: U goes up in row blocks and the step applies K2^-1 to every block as it lands (rows are independent),
    // so only the last block's share of that kernel is exposed behind the 2.4 ms upload
    const int n1 = plan->d.n1, n2 = plan->d.n2;
    const int chunks = (plan->d.dim == 2 && uses_gs_path(*plan) && n1 >= 1024 && n1 % (2 * gphm_plan::kUChunks) == 0) ? gphm_plan::kUChunks : 0;
    if (chunks) {
        const size_t rows = (size_t)n1 / chunks;
        for (int c = 0; c < chunks; ++c) {
            GPHM_CUDA_OK(cudaMemcpyAsync(U + c * rows * n2, h_U + c * rows * n2, sizeof(double) * rows * n2, cudaMemcpyHostToDevice,
                                         plan->hs_stream));
            GPHM_CUDA_OK(cudaEventRecord(plan->hs_ev_chunk[c], plan->hs_stream));
        }
    } else {
        GPHM_CUDA_OK(cudaMemcpyAsync(U, h_U, sizeof(double) * nf, cudaMemcpyHostToDevice, plan->hs_stream));
    }
    GPHM_CUDA_OK(cudaEventRecord(plan->hs_ev_u, plan->hs_stream));
    if (!resident) {
        GPHM_CUDA_OK(cudaMemcpyAsync(mU, h_mU, sizeof(double) * nf, cudaMemcpyHostToDevice, plan->hs_stream));
        GPHM_CUDA_OK(cudaMemcpyAsync(vU, h_vU, sizeof(double) * nf, cudaMemcpyHostToDevice, plan->hs_stream));
    }
    struct Ctx { double *U, *mU, *vU, *hU, *hmU, *hvU; size_t nf; double lr; bool resident; };
    static thread_local Ctx ctx;
    ctx = Ctx{U, mU, vU, h_U, h_mU, h_vU, nf, lr, resident};
    plan->u_ready = plan->hs_ev_u;
    plan->u_chunks = chunks;
    plan->gu_hook_ran = false;
    plan->on_gu = [](gphm_plan& p, cudaStream_t s) -> int {                  // dL/dU complete on s
        GPHM_CUDA_OK(cudaEventRecord(p.hs_ev_gu, s));
        GPHM_CUDA_OK(cudaStreamWaitEvent(p.hs_stream, p.hs_ev_gu, 0));
        if (ctx.resident && p.u_chunks > 0 && ctx.nf % p.u_chunks == 0) {
            // Adam(U) and the download block by block: only the first block's update is exposed before PCIe is busy again
            const size_t blk = ctx.nf / p.u_chunks;
            for (int c = 0; c < p.u_chunks; ++c) {
                const size_t o = c * blk;
                GPHM_TRY(launch_adam(ctx.U + o, p.gU + o, ctx.mU + o, ctx.vU + o, blk, p.hs_count, ctx.lr, p.hs_stream));
                if (c == p.u_chunks - 1) GPHM_CUDA_OK(cudaEventRecord(p.hs_ev_adam, p.hs_stream));
                GPHM_CUDA_OK(cudaMemcpyAsync(ctx.hU + o, ctx.U + o, sizeof(double) * blk, cudaMemcpyDeviceToHost, p.hs_stream));
            }
            return GPHM_OK;
        }
        GPHM_TRY(launch_adam(ctx.U, p.gU, ctx.mU, ctx.vU, ctx.nf, p.hs_count, ctx.lr, p.hs_stream));
        GPHM_CUDA_OK(cudaEventRecord(p.hs_ev_adam, p.hs_stream));
        GPHM_CUDA_OK(cudaMemcpyAsync(ctx.hU, ctx.U, sizeof(double) * ctx.nf, cudaMemcpyDeviceToHost, p.hs_stream));
        if (!ctx.resident) {
            GPHM_CUDA_OK(cudaMemcpyAsync(ctx.hmU, ctx.mU, sizeof(double) * ctx.nf, cudaMemcpyDeviceToHost, p.hs_stream));
            GPHM_CUDA_OK(cudaMemcpyAsync(ctx.hvU, ctx.vU, sizeof(double) * ctx.nf, cudaMemcpyDeviceToHost, p.hs_stream));
        }
        return GPHM_OK;
    };
    const int rc = logjoint_grad(*plan, U, sm, plan->gU, plan->gsmall, terms, 0, st);
    plan->u_ready = nullptr;
    plan->u_chunks = 0;
    plan->on_gu = nullptr;
    if (rc != GPHM_OK) { cudaStreamSynchronize(plan->hs_stream); cudaStreamSynchronize(st); return rc; }
    if (!plan->gu_hook_ran) {            // dense path: no early hand-over, Adam(U) after the whole gradient
        GPHM_CUDA_OK(cudaEventRecord(plan->hs_ev_mv, plan->hs_stream));
        GPHM_CUDA_OK(cudaStreamWaitEvent(st, plan->hs_ev_mv, 0));
        GPHM_TRY(launch_adam(U, plan->gU, mU, vU, nf, plan->hs_count, lr, st));
        GPHM_CUDA_OK(cudaMemcpyAsync(h_U, U, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
        if (!resident) {
            GPHM_CUDA_OK(cudaMemcpyAsync(h_mU, mU, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
            GPHM_CUDA_OK(cudaMemcpyAsync(h_vU, vU, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
        }
    }
    GPHM_TRY(launch_adam(sm, plan->gsmall, msm, vsm, ns, plan->hs_count, lr, st));
    if (plan->gu_hook_ran) GPHM_CUDA_OK(cudaStreamWaitEvent(st, plan->hs_ev_adam, 0));      // Adam(U) has read the count
    GPHM_TRY(launch_count_inc(plan->hs_count, st));
    GPHM_CUDA_OK(cudaMemcpyAsync(h_small, sm, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
    if (!resident) {
        GPHM_CUDA_OK(cudaMemcpyAsync(h_msmall, msm, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
        GPHM_CUDA_OK(cudaMemcpyAsync(h_vsmall, vsm, sizeof(double) * ns, cudaMemcpyDeviceToHost, st));
    }
    if (h_count) GPHM_CUDA_OK(cudaMemcpyAsync(h_count, plan->hs_count, sizeof(long long), cudaMemcpyDeviceToHost, st));
    GPHM_CUDA_OK(cudaMemcpyAsync(h_terms, terms, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
    GPHM_CUDA_OK(cudaStreamSynchronize(st));
    GPHM_CUDA_OK(cudaStreamSynchronize(plan->hs_stream));
    return GPHM_OK;
}

int gphm_step_host(gphm_plan* plan, double* h_U, double* h_small, double* h_mU, double* h_vU, double* h_msmall,
                   double* h_vsmall, long long* h_count, double lr, double* h_terms, void* stream) {
    if (!plan || !h_U || !h_small || !h_mU || !h_vU || !h_msmall || !h_vsmall || !h_count || !h_terms) {
        set_last_error("gphm_step_host: null pointer");
        return GPHM_EINVAL;
    }
    return step_host_impl(plan, h_U, h_small, h_mU, h_vU, h_msmall, h_vsmall, h_count, lr, h_terms, false, false,
                          static_cast<cudaStream_t>(stream));
}

int gphm_step_host_params(gphm_plan* plan, double* h_U, double* h_small, int reset_opt, long long* h_count, double lr,
                          double* h_terms, void* stream) {
    if (!plan || !h_U || !h_small || !h_terms) { set_last_error("gphm_step_host_params: null pointer"); return GPHM_EINVAL; }
    return step_host_impl(plan, h_U, h_small, nullptr, nullptr, nullptr, nullptr, h_count, lr, h_terms, true, reset_opt != 0,
                          static_cast<cudaStream_t>(stream));
}

size_t gphm_predict_work_bytes(const gphm_plan* plan, int m1, int m2) {
    if (!plan || m1 <= 0) return 0;
    Carver c(nullptr);
    double* p;
    const size_t n1 = plan->d.n1, n2 = plan->d.n2;
    c.take(p, (size_t)m1 * n1); c.take(p, (size_t)m1 * n2); c.take(p, (size_t)m1 * n2); c.take(p, (size_t)m1 * n2);
    if (plan->d.dim == 2) c.take(p, (size_t)std::max(m2, 1) * n2);
    return c.off;
}

int gphm_predict(gphm_plan* plan, const double* d_U, const double* d_small, const double* d_xt, int m1,
                 const double* d_yt, int m2, double* d_out, void* d_work, void* stream) {
    if (!plan || !d_U || !d_small || !d_xt || !d_out || !d_work) { set_last_error("gphm_predict: null pointer"); return GPHM_EINVAL; }
    gphm_plan& p = *plan;
    const bool two = p.d.dim == 2;
    if (m1 <= 0 || (two && (m2 <= 0 || !d_yt))) { set_last_error("gphm_predict: bad test grid"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n1 = p.d.n1, n2 = p.d.n2, Q = p.d.Q;
    Carver c(d_work);
    double *Kmn1, *M1, *M1t, *M1K, *Kmn2 = nullptr;
    c.take(Kmn1, (size_t)m1 * n1); c.take(M1, (size_t)m1 * n2); c.take(M1t, (size_t)m1 * n2); c.take(M1K, (size_t)m1 * n2);
    if (two) c.take(Kmn2, (size_t)m2 * n2);
    if (two) GPHM_TRY(factor_both(p, d_small, false, false, st));
    else GPHM_TRY(factor_axis(p, 0, d_small, false, st));
    GPHM_TRY(apply_kinv(p, 0, 0, d_U, n1, n2, p.A, p.Tf, st));
    GPHM_TRY(launch_gram_general(p.d.kernel_id, 0, d_xt, m1, p.ax[0].x, n1, theta_of(p, d_small, 0), Q, 0.0, Kmn1,
                                 nullptr, n1, st));
    double* dst1 = two ? M1 : d_out;
    GPHM_TRY(launch_dgemm(gemm_args(Kmn1, n1, false, p.A, n2, false, dst1, n2, m1, n2, n1), st));
    if (two) {
        GPHM_TRY(apply_kinv(p, 1, 1, M1, m1, n2, M1K, M1t, st));
        GPHM_TRY(launch_gram_general(p.d.kernel_id, 0, d_yt, m2, p.ax[1].x, n2, theta_of(p, d_small, 1), Q, 0.0, Kmn2,
                                     nullptr, n2, st));
        GPHM_TRY(launch_dgemm(gemm_args(M1K, n2, false, Kmn2, n2, true, d_out, m2, m1, m2, n2), st));
    }
    return GPHM_OK;
}

size_t gphm_rel_l2_work_bytes(void) { return sizeof(double) * 2 * kRedBlocks; }

int gphm_rel_l2(const double* d_pred, const double* d_truth, size_t n, double* d_out, void* d_work, void* stream) {
    if (!d_pred || !d_truth || !d_out || !d_work) { set_last_error("gphm_rel_l2: null pointer"); return GPHM_EINVAL; }
    return launch_rel_l2(d_pred, d_truth, n, static_cast<double*>(d_work), d_out, static_cast<cudaStream_t>(stream));
}

int gphm_plan_factor(gphm_plan* plan, const double* d_small, int axis_mask, void* stream) {
    if (!plan || !d_small) { set_last_error("gphm_plan_factor: null pointer"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool kinv = (axis_mask & 4) == 0;
    if ((axis_mask & 3) == 3 && plan->d.dim == 2) return factor_both(*plan, d_small, kinv, kinv, st);
    if (axis_mask & 1) GPHM_TRY(factor_axis(*plan, 0, d_small, kinv, st));
    if ((axis_mask & 2) && plan->d.dim == 2) GPHM_TRY(factor_axis(*plan, 1, d_small, kinv, st));
    return GPHM_OK;
}

int gphm_plan_uses_fft(const gphm_plan* plan, int axis) {
    if (!plan || axis < 0 || axis > 1) return 0;
    return plan->ax[axis].n > 0 && plan->ax[axis].fftL > 0 ? 1 : 0;
}

int gphm_plan_uses_gs(const gphm_plan* plan, int axis) {
    if (!plan || axis < 0 || axis > 1) return 0;
    return plan->ax[axis].n > 0 && plan->ax[axis].gs ? 1 : 0;
}

int gphm_mg_toeplitz_apply(gphm_plan* plan, int axis, int transposed, const double* d_X, int rows, double alpha, double beta,
                           const double* d_small, double* d_out, void* stream) {
    if (!plan || !d_X || !d_out || !d_small) { set_last_error("gphm_mg_toeplitz_apply: null pointer"); return GPHM_EINVAL; }
    if (axis < 0 || axis > 1 || plan->ax[axis].n == 0 || plan->ax[axis].fftL == 0) { set_last_error("gphm_mg_toeplitz_apply: axis %d has no FFT path", axis); return GPHM_EINVAL; }
    Axis& X = plan->ax[axis];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool anti = deriv_order(*plan) == 1;
    // the table is rebuilt here (n*Q evaluations): this rank may not have factored this axis itself
    GPHM_TRY(launch_toeplitz_table(plan->d.kernel_id, deriv_order(*plan), X.x, X.n, theta_of(*plan, d_small, axis), plan->d.Q,
                                   X.tabK, X.tabD, st));
    GPHM_TRY(launch_toeplitz_spectrum(X.tabD, X.n, X.fftL, X.twid, anti, X.dirsign, X.specT, st));
    return launch_toeplitz_apply(d_X, rows, X.n, X.n, X.specT, X.fftL, X.twid, (anti && transposed) ? -alpha : alpha, beta,
                                 d_out, X.n, st);
}

int gphm_mg_toeplitz_rows(gphm_plan* plan, int axis, int transposed, const double* d_X, int rows, double alpha, double beta,
                          const double* d_add, double* d_out, int keep_spectrum, void* stream) {
    if (!plan || !d_X || !d_out) { set_last_error("gphm_mg_toeplitz_rows: null pointer"); return GPHM_EINVAL; }
    if (axis < 0 || axis > 1 || plan->ax[axis].n == 0 || !plan->ax[axis].gs || (plan->d.force_general & 8) ||
        !toeplitz_fused_supported(plan->ax[axis].fftL)) {
        set_last_error("gphm_mg_toeplitz_rows: axis %d is not on the Toeplitz inverse-generator path", axis);
        return GPHM_EINVAL;
    }
    Axis& X = plan->ax[axis];
    const size_t other = plan->d.dim == 2 ? (axis == 0 ? (size_t)plan->d.n2 : (size_t)plan->d.n1) : 1;
    if (rows < 0 || (keep_spectrum && (size_t)rows > other)) { set_last_error("gphm_mg_toeplitz_rows: %d rows exceed the plan's spectrum store", rows); return GPHM_EINVAL; }
    const bool anti = deriv_order(*plan) == 1;
    return launch_toeplitz_apply_fused(d_X, rows, X.n, X.n, X.specT, X.fftL, X.twid, (anti && transposed) ? -alpha : alpha, beta,
                                       d_add, X.n, d_out, X.n, keep_spectrum ? X.specY : nullptr, static_cast<cudaStream_t>(stream));
}

int gphm_mg_theta_grad_pairs(gphm_plan* plan, int axis, const double* d_V, const double* d_G, int rows, int lead, double beta,
                             double cD, const double* d_small, double* d_gtheta, void* stream) {
    if (!plan || !d_V || !d_G || !d_small || !d_gtheta) { set_last_error("gphm_mg_theta_grad_pairs: null pointer"); return GPHM_EINVAL; }
    if (axis < 0 || axis > 1 || plan->ax[axis].n == 0 || !plan->ax[axis].gs || !toeplitz_fused_supported(plan->ax[axis].fftL)) {
        set_last_error("gphm_mg_theta_grad_pairs: axis %d is not on the Toeplitz inverse-generator path", axis);
        return GPHM_EINVAL;
    }
    Axis& X = plan->ax[axis];
    const int n = X.n, order = deriv_order(*plan);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GPHM_TRY(launch_xcorr_pairs(d_V, rows, n, n, X.specY, X.fftL, X.twid, -1.0, X.specK, st));
    GPHM_TRY(launch_xcorr_pairs(d_G, rows, n, n, X.specY, X.fftL, X.twid, cD, X.specD, st));
    GPHM_TRY(launch_spectrum_to_diag_sums(X.specK, X.specD, X.fftL, X.twid, n, order == 1, X.dirsign, lead ? X.sKinv : nullptr,
                                          lead ? beta : 0.0, X.sK, X.sD, st));
    return launch_theta_grad_toeplitz(plan->d.kernel_id, order, X.x, n, theta_of(*plan, d_small, axis), plan->d.Q, X.sK, X.sD,
                                      d_gtheta, st);
}

int gphm_mg_theta_grad_pairs_both(gphm_plan* plan, const double* d_V1, const double* d_G1, int rows1, const double* d_V2,
                                  const double* d_G2, int rows2, int lead, double beta1, double beta2, double cD1, double cD2,
                                  const double* d_small, double* d_gtheta, void* stream) {
    if (!plan || !d_V1 || !d_G1 || !d_V2 || !d_G2 || !d_small || !d_gtheta) { set_last_error("gphm_mg_theta_grad_pairs_both: null pointer"); return GPHM_EINVAL; }
    for (int a = 0; a < 2; ++a)
        if (plan->ax[a].n == 0 || !plan->ax[a].gs || !toeplitz_fused_supported(plan->ax[a].fftL)) {
            set_last_error("gphm_mg_theta_grad_pairs_both: axis %d is not on the Toeplitz inverse-generator path", a);
            return GPHM_EINVAL;
        }
    Axis& X1 = plan->ax[0];
    Axis& X2 = plan->ax[1];
    if (X1.fftL != X2.fftL) {      // different transform lengths: nothing to share
        GPHM_TRY(gphm_mg_theta_grad_pairs(plan, 0, d_V1, d_G1, rows1, lead, beta1, cD1, d_small, d_gtheta, stream));
        return gphm_mg_theta_grad_pairs(plan, 1, d_V2, d_G2, rows2, lead, beta2, cD2, d_small, d_gtheta + 3 * plan->d.Q, stream);
    }
    const int order = deriv_order(*plan), Q = plan->d.Q;
    const bool anti = order == 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GPHM_TRY(launch_xcorr_pairs(d_V1, rows1, X1.n, X1.n, X1.specY, X1.fftL, X1.twid, -1.0, X1.specK, st));
    GPHM_TRY(launch_xcorr_pairs(d_G1, rows1, X1.n, X1.n, X1.specY, X1.fftL, X1.twid, cD1, X1.specD, st));
    GPHM_TRY(launch_xcorr_pairs(d_V2, rows2, X2.n, X2.n, X2.specY, X2.fftL, X2.twid, -1.0, X2.specK, st));
    GPHM_TRY(launch_xcorr_pairs(d_G2, rows2, X2.n, X2.n, X2.specY, X2.fftL, X2.twid, cD2, X2.specD, st));
    const DiagSumsJob dj[2] = {
        {X1.specK, X1.specD, X1.twid, X1.n, anti, X1.dirsign, lead ? X1.sKinv : nullptr, lead ? beta1 : 0.0, X1.sK, X1.sD},
        {X2.specK, X2.specD, X2.twid, X2.n, anti, X2.dirsign, lead ? X2.sKinv : nullptr, lead ? beta2 : 0.0, X2.sK, X2.sD}};
    GPHM_TRY(launch_spectrum_to_diag_sums_multi(dj, 2, X1.fftL, st));
    const ThetaGradJob tj[2] = {{X1.x, X1.n, theta_of(*plan, d_small, 0), X1.sK, X1.sD, d_gtheta},
                                {X2.x, X2.n, theta_of(*plan, d_small, 1), X2.sK, X2.sD, d_gtheta + 3 * Q}};
    return launch_theta_grad_toeplitz_multi(plan->d.kernel_id, order, tj, 2, Q, st);
}

int gphm_transpose(const double* d_in, int rows, int cols, double* d_out, void* stream) {
    if (rows <= 0 || cols <= 0) return GPHM_OK;
    if (!d_in || !d_out) { set_last_error("gphm_transpose: null pointer"); return GPHM_EINVAL; }
    return launch_transpose(d_in, rows, cols, d_out, static_cast<cudaStream_t>(stream));
}

int gphm_mg_pack_transposed(const double* d_in, int rows, int cols, int part_cols, size_t part_stride, double* d_out, void* stream) {
    if (rows <= 0 || cols <= 0) return GPHM_OK;
    if (!d_in || !d_out || part_cols <= 0 || cols % part_cols) { set_last_error("gphm_mg_pack_transposed: bad argument"); return GPHM_EINVAL; }
    return launch_transpose_parts(d_in, rows, cols, part_cols, part_stride, d_out, static_cast<cudaStream_t>(stream));
}

// ---- exchange buffers in peer-mapped device memory (cudaIpc) --------------------------------------------------------
int gphm_mg_peer_alloc(size_t data_doubles, void** d_base, unsigned char* h_handle64) {
    if (!d_base || !h_handle64) { set_last_error("gphm_mg_peer_alloc: null pointer"); return GPHM_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr;
    const size_t bytes = peer_flag_bytes() + sizeof(double) * data_doubles + sizeof(unsigned int) * 4;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { set_last_error("gphm_mg_peer_alloc: cudaMalloc(%zu) failed", bytes); return GPHM_ENOMEM; }
    GPHM_CUDA_OK(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    GPHM_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    memcpy(h_handle64, &h, 64);
    *d_base = p;
    return GPHM_OK;
}

int gphm_mg_peer_open(const unsigned char* h_handle64, void** d_peer_base) {
    if (!h_handle64 || !d_peer_base) { set_last_error("gphm_mg_peer_open: null pointer"); return GPHM_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    GPHM_CUDA_OK(cudaIpcOpenMemHandle(d_peer_base, h, cudaIpcMemLazyEnablePeerAccess));
    return GPHM_OK;
}

int gphm_mg_peer_close(void* d_peer_base) {
    if (d_peer_base) GPHM_CUDA_OK(cudaIpcCloseMemHandle(d_peer_base));
    return GPHM_OK;
}

int gphm_mg_peer_free(void* d_base) {
    if (d_base) GPHM_CUDA_OK(cudaFree(d_base));
    return GPHM_OK;
}

int gphm_mg_peer_exchange(const double* const* h_in, int k, int rows, int cols, int part_cols, void* const* h_peer_bases, int P,
                          int me, unsigned long long seq, size_t data_doubles, int* d_status, void* stream) {
    if (!h_in || !h_peer_bases) { set_last_error("gphm_mg_peer_exchange: null pointer"); return GPHM_EINVAL; }
    if ((size_t)k * rows * cols > data_doubles) { set_last_error("gphm_mg_peer_exchange: buffer too small"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* bases[16];
    if (P > 16) { set_last_error("gphm_mg_peer_exchange: P=%d > 16", P); return GPHM_EINVAL; }
    for (int d = 0; d < P; ++d) bases[d] = static_cast<double*>(h_peer_bases[d]);
    // the completion counter lives behind the data of this rank's own buffer
    unsigned int* done = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(h_peer_bases[me]) + peer_flag_bytes() + sizeof(double) * data_doubles);
    GPHM_TRY(launch_a2a_transpose_peer(h_in, k, rows, cols, part_cols, bases, P, me, seq, done, st));
    return launch_mg_wait_flags(h_peer_bases[me], P, seq, d_status, st);
}

int gphm_mg_unpack_segments(const double* d_recv, int parts, int arrays, int rows, int seg, double* d_out, void* stream) {
    if (parts <= 0 || arrays <= 0 || rows <= 0 || seg <= 0) return GPHM_OK;
    if (!d_recv || !d_out) { set_last_error("gphm_mg_unpack_segments: null pointer"); return GPHM_EINVAL; }
    return launch_unpack_segments(d_recv, parts, arrays, rows, seg, d_out, static_cast<cudaStream_t>(stream));
}

int gphm_mg_theta_grad_fft(gphm_plan* plan, int axis, const double* d_X, const double* d_Y, const double* d_G, int rows,
                           int linv_row0, int linv_row1, double beta, double cD, const double* d_small,
                           double* d_gtheta, void* stream) {
    if (!plan || !d_X || !d_Y || !d_G || !d_small || !d_gtheta) { set_last_error("gphm_mg_theta_grad_fft: null pointer"); return GPHM_EINVAL; }
    if (axis < 0 || axis > 1 || plan->ax[axis].n == 0 || plan->ax[axis].fftL == 0) { set_last_error("gphm_mg_theta_grad_fft: axis %d has no FFT path", axis); return GPHM_EINVAL; }
    Axis& X = plan->ax[axis];
    const int n = X.n;
    if (rows < 0 || linv_row0 < 0 || linv_row1 > n || linv_row0 > linv_row1) { set_last_error("gphm_mg_theta_grad_fft: bad row range"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int order = deriv_order(*plan);
    // rows == 0 / empty Linv range still have to define the spectra: weight 0 passes over row 0 of Linv
    const bool gs_lead = X.gs && linv_row0 == 0 && linv_row1 > 0;     // GS axes: the rank owning row 0 adds the K^-1 sums
    if (!X.gs)
        GPHM_TRY(launch_xcorr_spectrum(X.Linv + (size_t)linv_row0 * n, X.Linv + (size_t)linv_row0 * n, linv_row1 - linv_row0, n, n, n,
                                       X.fftL, X.twid, beta, false, X.specK, st));
    GPHM_TRY(launch_xcorr_spectrum(d_X, d_Y, rows, n, n, n, X.fftL, X.twid, -1.0, !X.gs, X.specK, st));
    GPHM_TRY(launch_xcorr_spectrum(d_G, d_Y, rows, n, n, n, X.fftL, X.twid, cD, false, X.specD, st));
    GPHM_TRY(launch_spectrum_to_diag_sums(X.specK, X.specD, X.fftL, X.twid, n, order == 1, X.dirsign, gs_lead ? X.sKinv : nullptr,
                                          gs_lead ? beta : 0.0, X.sK, X.sD, st));
    return launch_theta_grad_toeplitz(plan->d.kernel_id, order, X.x, n, theta_of(*plan, d_small, axis), plan->d.Q, X.sK, X.sD,
                                      d_gtheta, st);
}

int gphm_apply_kinv(gphm_plan* plan, int axis, int side, const double* d_X, int rows, int cols, double* d_out,
                    double* d_tmp, void* stream) {
    if (!plan || !d_X || !d_out || !d_tmp) { set_last_error("gphm_apply_kinv: null pointer"); return GPHM_EINVAL; }
    if (axis < 0 || axis > 1 || plan->ax[axis].n == 0 || (side != 0 && side != 1)) { set_last_error("gphm_apply_kinv: bad axis/side"); return GPHM_EINVAL; }
    if (rows <= 0 || cols <= 0) return GPHM_OK;
    return apply_kinv(*plan, axis, side, d_X, rows, cols, d_out, d_tmp, static_cast<cudaStream_t>(stream));
}

int gphm_apply_kinv_rows_refined(gphm_plan* plan, int axis, const double* d_X, int rows, double* d_out, double* d_tmp,
                                 void* stream) {
    if (!plan || !d_X || !d_out || !d_tmp) { set_last_error("gphm_apply_kinv_rows_refined: null pointer"); return GPHM_EINVAL; }
    if (axis < 0 || axis > 1 || plan->ax[axis].n == 0) { set_last_error("gphm_apply_kinv_rows_refined: bad axis"); return GPHM_EINVAL; }
    if (rows <= 0) return GPHM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Axis& X = plan->ax[axis];
    GPHM_TRY(apply_kinv(*plan, axis, 1, d_X, rows, X.n, d_out, d_tmp, st));
    if (X.gs && gs_refine_axis(*plan, X) && toeplitz_fused_supported(X.fftL)) {
        GPHM_TRY(launch_copy(d_tmp, d_X, (size_t)rows * X.n, st));             // keep the caller's right-hand side intact
        GPHM_TRY(refine_kinv_rows_gs(X, d_tmp, rows, d_out, d_tmp, st));
    }
    return GPHM_OK;
}

int gphm_plan_logdet(gphm_plan* plan, double* d_out2, void* stream) {
    if (!plan || !d_out2) { set_last_error("gphm_plan_logdet: null pointer"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GPHM_TRY(launch_sum_scaled(plan->ax[0].ldpart, plan->ax[0].nblk, 2.0, d_out2, st));
    if (plan->d.dim == 2) GPHM_TRY(launch_sum_scaled(plan->ax[1].ldpart, plan->ax[1].nblk, 2.0, d_out2 + 1, st));
    else GPHM_CUDA_OK(cudaMemsetAsync(d_out2 + 1, 0, sizeof(double), st));
    return GPHM_OK;
}

int gphm_mg_residual(gphm_plan* plan, double* d_R, const double* d_U, const double* d_F, const double* d_A,
                     const double* d_Bt, size_t n_local, const double* d_small, double* d_out2, void* stream) {
    if (!plan || !d_R || !d_U || !d_F || !d_A || !d_Bt || !d_small || !d_out2) { set_last_error("gphm_mg_residual: null pointer"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GPHM_TRY(launch_residual(d_R, d_U, d_F, d_A, d_Bt, n_local, plan->d.eq_type, nullptr, d_small, plan->d.Q, plan->part, st));
    return launch_pair_reduce(plan->part, d_out2, st);
}

int gphm_mg_boundary(const double* d_U, const int* d_bidx, const double* d_bvals, int nb_local, double* d_eb,
                     double* d_out1, void* stream) {
    if (!d_U || !d_out1 || (nb_local > 0 && (!d_bidx || !d_bvals || !d_eb))) { set_last_error("gphm_mg_boundary: null pointer"); return GPHM_EINVAL; }
    return launch_boundary_indexed(d_U, d_bidx, d_bvals, nb_local, d_eb, d_out1, static_cast<cudaStream_t>(stream));
}

int gphm_mg_finalize(gphm_plan* plan, const double* d_sums3, const double* d_ld2, const double* d_small, double* d_terms,
                     double* d_gsmall, void* stream) {
    if (!plan || !d_sums3 || !d_ld2 || !d_small || !d_terms) { set_last_error("gphm_mg_finalize: null pointer"); return GPHM_EINVAL; }
    return launch_mg_finalize(loss_consts(*plan), d_sums3, d_ld2, d_small, d_terms, d_gsmall, plan->status,
                              static_cast<cudaStream_t>(stream));
}

int gphm_mg_grad_u(gphm_plan* plan, const double* d_U, const double* d_G, const double* d_W, const double* d_S1,
                   const double* d_S2, size_t n_local, const int* d_bidx, const double* d_eb, int nseg0, int nb_local,
                   const double* d_small, double* d_gU, double* d_V2, void* stream) {
    if (!plan || !d_U || !d_G || !d_W || !d_S1 || !d_small || !d_gU) { set_last_error("gphm_mg_grad_u: null pointer"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GPHM_TRY(launch_grad_u_local(n_local, plan->d.eq_type == GPHM_EQ_ALLENCAHN, d_U, d_G, d_W, d_S1, d_S2, d_gU, nullptr, d_V2, st));
    if (nb_local > 0)
        GPHM_TRY(launch_boundary_scatter_indexed(d_gU, d_bidx, d_eb, nseg0, nb_local, plan->d.llk_weight,
                                                 d_small + 6 * plan->d.Q, st));
    return GPHM_OK;
}

int gphm_lincomb(double* d_out, double a, const double* d_x, double b, const double* d_y, size_t n, void* stream) {
    if (n == 0) return GPHM_OK;
    if (!d_out || !d_x) { set_last_error("gphm_lincomb: null pointer"); return GPHM_EINVAL; }
    return launch_lincomb(d_out, a, d_x, b, d_y, n, static_cast<cudaStream_t>(stream));
}

int gphm_mg_theta_grad(gphm_plan* plan, int axis, const double* d_Kbar, const double* d_Dbar, const double* d_small,
                       double* d_gtheta, void* stream) {
    if (!plan || !d_Kbar || !d_Dbar || !d_small || !d_gtheta) { set_last_error("gphm_mg_theta_grad: null pointer"); return GPHM_EINVAL; }
    if (axis < 0 || axis > 1 || plan->ax[axis].n == 0) { set_last_error("gphm_mg_theta_grad: bad axis"); return GPHM_EINVAL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Axis& X = plan->ax[axis];
    const int order = deriv_order(*plan);
    const double* th = theta_of(*plan, d_small, axis);
    if (X.toeplitz) {
        GPHM_TRY(launch_diag_sums(d_Kbar, d_Dbar, X.n, X.n, order == 1, X.dirsign, X.dspart, X.sK, X.sD, st));
        return launch_theta_grad_toeplitz(plan->d.kernel_id, order, X.x, X.n, th, plan->d.Q, X.sK, X.sD, d_gtheta, st);
    }
    return launch_theta_grad_general(plan->d.kernel_id, order, X.x, X.n, th, plan->d.Q, d_Kbar, d_Dbar, X.n, X.tgpart,
                                     d_gtheta, st);
}

const double* gphm_plan_matrix(const gphm_plan* plan, int axis, int which) {
    if (!plan || axis < 0 || axis > 1 || plan->ax[axis].n == 0) return nullptr;
    const Axis& X = plan->ax[axis];
    switch (which) {
        case 0: return X.Kinv;
        case 1: return X.D;
        case 2: return X.Linv;
        case 3: return X.L;
        default: return nullptr;
    }
}

}  // extern "C"
