// K^-1 on uniform grids without a dense factorisation (FP64).
//
// On a uniform grid K = k(|x_i - x_j|) + jitter*I is a symmetric positive definite Toeplitz
// matrix, so everything the step needs from it - log|K| (jnp.linalg.slogdet,
// model_GP_solver_2d.py:158-161), the K^-1 applications (jnp.linalg.solve, :104-105 and their
// reverse pass) and the diagonal sums of K^-1 for the theta-gradient - follows from the first
// column g = K^-1 e_0 alone:
//
//   schur_levinson_kernel   Schur recursion on the generator of K (reflection coefficients kappa_k
//       without inner products; backward stable for SPD Toeplitz matrices) fused with the Levinson
//       lattice  A_k = A_{k-1} + kappa_k z B_{k-1},  B_k = z B_{k-1} + kappa_k A_{k-1}  whose final
//       polynomial gives g = A_{n-1} / E_{n-1}.   log|K| = sum_k log E_k,  E_k = E_{k-1}(1 - kappa_k^2).
//       O(n^2) work, n sequential steps with one block barrier each; one CTA per axis, both axes
//       of a 2-D problem run concurrently.  Replaces the blocked Cholesky + L^-1 (N^3 FLOPs).
//   gs_prepare_kernel       Gohberg-Semencul:  K^-1 = (L(g) L(g)^T - L(h) L(h)^T) / g_0,
//       h = (0, g_{n-1}, ..., g_1), L(c) = lower-triangular Toeplitz with first column c.  Emits the
//       four circulant spectra that let launch_toeplitz_apply perform  v -> K^-1 v  for every
//       row of a matrix as four FFT convolutions (O(N^2 log N) per K^-1 application instead of
//       2 N^3), and the diagonal sums  s[d] = sum_i K^-1[i][i+d]  from two cross-correlations:
//       sum_i (L(c) L(c)^T)[i][i+d] = sum_p (n - d - p) c_p c_{p+d} = xcorr(c, u)[d],  u_p = (n - p) c_p.
//
// Accuracy (tools/toeplitz_numerics_*.py, against an extended-precision solve): at N = 4096,
// cond(K) = 5e7 the GS application is as accurate as a Cholesky solve (2.5e-10 vs 6e-10
// relative), g from Schur+lattice 1e-9, log|K| 4e-11 relative.
#include <algorithm>
#include <cmath>
#include "common.cuh"
#include "kernels.h"
#include "fft_core.cuh"

namespace gphm {

constexpr int SCHUR_EPT = 16;                   // consecutive elements per thread (8 and 4 measured: same or slower)
constexpr int SCHUR_MAX_THREADS = 256;
constexpr int SCHUR_MAX_N = SCHUR_EPT * SCHUR_MAX_THREADS;

int toeplitz_inv_max_n() { return SCHUR_MAX_N; }

// Two CTAs per system: the GENERATOR CTA runs the Schur recursion (it alone carries the serial
// dependency kappa_k -> kappa_{k+1}) and hands the reflection coefficients over through global memory
// in chunks; the LATTICE CTA consumes them (no feedback) and builds A_{n-1}.  Splitting halves the
// FP64 work on the critical path.
//
// Register layout (both roles): thread t owns positions j = 8t .. 8t+7 of
//   generator:  be[j]  = beta_{k-1}[j]      second generator row             (live for j >= k)
//               A[j]   = alpha_{k-1}[j-1]   first generator row, pre-shifted (the recursion shifts it by one
//                                           position per step; keeping it shifted puts the pair that defines
//                                           kappa_k = -beta[k] / alpha[k-1] into one thread)
//   lattice:    a[j]   = A_{k-1}[j],  B[j] = B_{k-1}[j-1]                    (non-zero for j <= k)
// Step k:  alpha_k = A + kappa be,  beta_k = be + kappa A;   A_k = a + kappa B,  B_k = B + kappa a,
// then the shifted sequence moves up by one position.  The shift costs no register moves: the loop is
// unrolled by 8 and at unrolled slot I the logical entry i of a shifted array lives in the physical
// register (i - I) mod 8, updates are in place, and only the entry that leaves the thread travels
// (warp shuffle; one value per warp through shared memory, patched after the barrier).
// kappa_{k+1} is produced by its owner (operands: physical register 0 / slot I+1) before the step's only
// barrier, except when the operand crosses a warp boundary (every 256th step: one more barrier).
// Warps whose generator entries are all dead / lattice entries all zero skip the update.
constexpr int SCHUR_CHUNK = 64;                 // reflection coefficients per hand-over

// -num/den to ~1 ulp in FIVE dependent FP64 operations after den is known (a division is seven):
//   r0 = MUFU.RCP64H(den) (2^-23), e = 1 - den r0, w = e + e^2, q0 = -num r0 (parallel to e),
//   result = q0 + q0 w = -num r0 (1 + e + e^2) = -num/den (1 - e^3).
// Straight-line (no slow path), so that the scheduler interleaves the step's independent FMAs with this
// chain - the one dependent chain of the whole recursion; every lane of a live warp evaluates it (only
// the owner's value is stored; other lanes may produce Inf/NaN harmlessly).
__device__ __forceinline__ double neg_div(double num, double den) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
    const double e = fma(-den, r0, 1.0);
    const double q0 = -num * r0;
    const double w = fma(e, e, e);
    return fma(q0, w, q0);
}
struct SchurTiming { long long t_kappa = 0, t_rest = 0, t_barrier = 0, t_last = 0; bool on = false; };

template <int I>
__device__ __forceinline__ void gen_step(int k, int n, int tid, int lane, int warp, double (&A)[SCHUR_EPT],
                                         double (&be)[SCHUR_EPT], double* kap, double (*bnd)[32], SchurTiming& tm) {
    constexpr int E = SCHUR_EPT;
    constexpr int I1 = (I + 1) % E;
    constexpr int OUT = (E - 1 - I) % E;       // physical slot of logical entry E-1 (leaves the thread)
    const double kp = kap[k];
    const int whi = warp * 32 * E + 32 * E - 1;
    const int owner = (k + 1) / E;
    const bool cross = (I1 == 0) && ((owner & 31) == 0);      // kappa_{k+1}'s alpha comes from the previous warp
    double cand = 0.0;                         // kappa_{k+1} in the owner thread
    if (whi >= k) {                            // warp still holds live generator entries
#pragma unroll
        for (int ii = 0; ii < E; ++ii) {
            const int i = (I + ii) % E;        // start with the entries that define kappa_{k+1}
            const int ph = (i - I + E) % E;
            const double al = A[ph], b = be[i];
            A[ph] = fma(kp, b, al);
            be[i] = fma(kp, al, b);
            if (ii == 1 && I1 != 0) cand = neg_div(be[I1], A[0]);
        }
    }
    const double out = A[OUT];
    const double up = __shfl_up_sync(0xffffffffu, out, 1);
    if (lane == 31) bnd[k & 1][warp] = out;
    A[OUT] = up;                               // becomes logical entry 0 of the next step; lane 0 is patched below
    if (I1 == 0 && !cross) cand = neg_div(be[0], up);
    if (!cross && tid == owner && k + 1 < n) kap[k + 1] = cand;
    long long t1 = 0;
    if (tm.on && tid == owner) {               // owner thread: barrier release -> kappa stored
        t1 = clock64();
        tm.t_kappa += t1 - tm.t_last;
    }
    __syncthreads();
    if (tm.on) {
        const long long t2 = clock64();
        if (tid == owner) tm.t_barrier += t2 - t1;
        tm.t_last = t2;
    }
    if (lane == 0) A[OUT] = warp > 0 ? bnd[k & 1][warp - 1] : 0.0;
    if (cross) {                                               // uniform in k
        if (tid == owner && k + 1 < n) kap[k + 1] = neg_div(be[0], A[OUT]);
        __syncthreads();
    }
}

template <int I>
__device__ __forceinline__ void lat_step(int k, int lane, int warp, double (&a)[SCHUR_EPT], double (&B)[SCHUR_EPT],
                                         const double* kap, double (*bnd)[32]) {
    constexpr int E = SCHUR_EPT;
    constexpr int OUT = (E - 1 - I) % E;
    const double kp = kap[k];
    if (warp * 32 * E <= k + 1) {              // warp holds non-zero lattice entries
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int ph = (i - I + E) % E;
            const double bs = B[ph], av = a[i];
            B[ph] = fma(kp, av, bs);
            a[i] = fma(kp, bs, av);
        }
    }
    const double out = B[OUT];
    const double up = __shfl_up_sync(0xffffffffu, out, 1);
    if (lane == 31) bnd[k & 1][warp] = out;
    B[OUT] = up;
    __syncthreads();
    if (lane == 0) B[OUT] = warp > 0 ? bnd[k & 1][warp - 1] : 0.0;
}

template <int I>
__device__ __forceinline__ void gen_steps(int kb, int n, int tid, int lane, int warp, double (&A)[SCHUR_EPT],
                                          double (&be)[SCHUR_EPT], double* kap, double (*bnd)[32], SchurTiming& tm) {
    if constexpr (I < SCHUR_EPT) {
        if (kb + I < n) gen_step<I>(kb + I, n, tid, lane, warp, A, be, kap, bnd, tm);
        gen_steps<I + 1>(kb, n, tid, lane, warp, A, be, kap, bnd, tm);
    }
}
template <int I>
__device__ __forceinline__ void lat_steps(int kb, int n, int lane, int warp, double (&a)[SCHUR_EPT], double (&B)[SCHUR_EPT],
                                          const double* kap, double (*bnd)[32]) {
    if constexpr (I < SCHUR_EPT) {
        if (kb + I < n) lat_step<I>(kb + I, lane, warp, a, B, kap, bnd);
        lat_steps<I + 1>(kb, n, lane, warp, a, B, kap, bnd);
    }
}

// blockIdx.x = 2 * system + role (0: generator, 1: lattice).  gkap[n] / prog[1] per system: hand-over buffer and
// the number of valid coefficients in it (zeroed by the launcher).
__global__ void __launch_bounds__(SCHUR_MAX_THREADS, 1)
schur_levinson_kernel(const double* __restrict__ tab, long long sTab, int n, double jitter, double* __restrict__ g,
                      long long sG, double* __restrict__ half_logdet, long long sLd, int* __restrict__ status,
                      long long sStatus, double* gkap, long long sKap, int* prog, long long sProg, long long* dbg) {
    const int sys = blockIdx.x >> 1, role = blockIdx.x & 1;
    const long long t_start = clock64();
    tab += sys * sTab; g += sys * sG; half_logdet += sys * sLd; status += sys * sStatus; gkap += sys * sKap; prog += sys * sProg;
    __shared__ double kap[SCHUR_MAX_N];
    __shared__ double bnd[2][32];
    __shared__ double red[34];
    __shared__ int bad, s_have;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j0 = tid * SCHUR_EPT;
    const double r0 = tab[0] + jitter;
    if (role == 0) {
        // ---- generator: Schur recursion ----
        double A[SCHUR_EPT], be[SCHUR_EPT];
        // State "before step 0": unshifted alpha_0 = r, beta_0 = (0, r_1, ...).  Step 0 runs with kappa_0 = 0:
        // it changes no value and performs the first shift, so that every step is identical.
#pragma unroll
        for (int i = 0; i < SCHUR_EPT; ++i) {
            const int j = j0 + i;
            const double rj = (j < n) ? (j == 0 ? r0 : tab[j]) : 0.0;
            A[i] = rj;
            be[i] = (j == 0) ? 0.0 : rj;
        }
        if (tid == 0) { bad = 0x7fffffff; kap[0] = 0.0; }
        __syncthreads();
        int published = 0;
        SchurTiming tm;
        tm.on = dbg != nullptr;
        tm.t_last = clock64();
        for (int kb = 0; kb < n; kb += SCHUR_EPT) {           // the position inside a thread is static per unrolled slot
            gen_steps<0>(kb, n, tid, lane, warp, A, be, kap, bnd, tm);
            const int valid = min(kb + SCHUR_EPT + 1, n);     // kappa_0 .. kappa_{valid-1} are final
            if (valid - published >= SCHUR_CHUNK || valid == n) {
                if (warp == 0) {
                    for (int i = published + lane; i < valid; i += 32) gkap[i] = kap[i];
                    __syncwarp();
                    if (lane == 0) { __threadfence(); *reinterpret_cast<volatile int*>(prog) = valid; }
                }
                published = valid;
            }
        }
        // log|K| = n log r0 + sum_k (n - k) log(1 - kappa_k^2); first |kappa| >= 1 <=> first non-positive prediction error
        double lsum = 0.0;
#pragma unroll
        for (int i = 0; i < SCHUR_EPT; ++i) {
            const int k = j0 + i;
            if (k >= 1 && k < n) {
                const double kp = kap[k];
                if (!(fabs(kp) < 1.0)) atomicMin(&bad, k);     // also catches NaN
                lsum += (double)(n - k) * log1p(-kp * kp);
            }
        }
        if (!(r0 > 0.0) && tid == 0) atomicMin(&bad, 0);
        const double ltot = block_sum(lsum, red);
        if (tm.on) {
            atomicAdd(reinterpret_cast<unsigned long long*>(dbg) + 8, (unsigned long long)tm.t_kappa);
            atomicAdd(reinterpret_cast<unsigned long long*>(dbg) + 9, (unsigned long long)tm.t_barrier);
        }
        if (tid == 0) {
            half_logdet[0] = 0.5 * ((double)n * log(r0) + ltot);
            if (bad != 0x7fffffff) status[0] = bad + 1;        // like a Cholesky pivot index
            if (dbg) dbg[blockIdx.x] = clock64() - t_start;
        }
        return;
    }
    // ---- lattice: A_k = A_{k-1} + kappa_k z B_{k-1},  B_k = z B_{k-1} + kappa_k A_{k-1} ----
    double a[SCHUR_EPT], B[SCHUR_EPT];
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) { a[i] = (j0 + i == 0) ? 1.0 : 0.0; B[i] = a[i]; }
    int have = 0;
    bool dead = false;
    for (int kb = 0; kb < n; kb += SCHUR_EPT) {
        const int need = min(kb + SCHUR_EPT, n);
        if (need > have) {                                     // uniform
            if (tid == 0) {
                int v = 0;
                long long spins = 0;
                while ((v = *reinterpret_cast<volatile int*>(prog)) < need && spins < (1ll << 24)) { __nanosleep(100); ++spins; }
                s_have = v;
            }
            __syncthreads();
            const int now = s_have;
            if (now < need) { dead = true; break; }            // the producer never arrived (cannot happen when both CTAs run)
            for (int i = have + tid; i < now; i += blockDim.x) kap[i] = __ldcg(gkap + i);
            __syncthreads();
            have = now;
        }
        lat_steps<0>(kb, n, lane, warp, a, B, kap, bnd);
    }
    // g = A_{n-1} / E_{n-1},  E_{n-1} = r0 * prod_k (1 - kappa_k^2)   (a does not shift: entry i is a[i])
    double prod = 1.0;
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) {
        const int k = j0 + i;
        if (k >= 1 && k < n) { const double kp = kap[k]; prod *= (1.0 - kp) * (1.0 + kp); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, o);
    __syncthreads();
    if (lane == 0) red[warp] = prod;
    __syncthreads();
    double E = r0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) E *= red[w];
    const double invE = dead ? __longlong_as_double(0x7ff8000000000000ll) : 1.0 / E;
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) { const int j = j0 + i; if (j < n) g[j] = a[i] * invE; }
    if (dbg && tid == 0) dbg[blockIdx.x] = clock64() - t_start;
}

// One CTA per axis.  spec[0..3][L] (bit-reversed order, scaled like launch_toeplitz_spectrum):
//   0: conj(G)/L   (v -> L(g)^T v)        2:  G / (L g0)   (v -> L(g) v / g0)
//   1: conj(H)/L   (v -> L(h)^T v)        3: -H / (L g0)   (v -> -L(h) v / g0)
// sKinv[d] = sum over |i-j| = d of K^-1[i][j]  (both triangles for d > 0).
__global__ void __launch_bounds__(FFT_THREADS, 1)
gs_prepare_kernel(const double* __restrict__ g, long long sG, int n, int L, int logL, const double2* __restrict__ W,
                  double2* __restrict__ spec, long long sSpec, double* __restrict__ sKinv, long long sS) {
    g += blockIdx.x * sG; spec += blockIdx.x * sSpec; sKinv += blockIdx.x * sS;
    extern __shared__ double2 xs[];
    const int tid = threadIdx.x;
    fft_load_twiddles(xs, L, logL, W, tid);
    const double g0 = g[0];
    const double wsc = exp2(-ceil(log2((double)n)));       // keeps u_p = (n-p) c_p at c's magnitude inside the shared FFT
    const double invL = 1.0 / (double)L, ig0 = 1.0 / g0;
    double2 acc[FFT_ACC];
#pragma unroll
    for (int k = 0; k < FFT_ACC; ++k) acc[k] = make_double2(0.0, 0.0);
    for (int pass = 0; pass < 2; ++pass) {
        for (int j = tid; j < L; j += FFT_THREADS) {
            double c = 0.0;
            if (j < n) c = pass == 0 ? g[j] : (j == 0 ? 0.0 : g[n - j]);
            xs[PADI(j)] = make_double2(c, (double)(n - j) * wsc * c);       // z = c + i u
        }
        __syncthreads();
        fft_dif_inplace(xs, L, logL, W, tid);
        double2* s_t = spec + (size_t)pass * L;            // L(c)^T
        double2* s_l = spec + (size_t)(2 + pass) * L;      // +-L(c)/g0
        const double sl = pass == 0 ? invL * ig0 : -invL * ig0;
#pragma unroll
        for (int k = 0; k < FFT_ACC; ++k) {
            const int p = tid + k * FFT_THREADS;
            if (p < L) {
                const unsigned f = __brev((unsigned)p) >> (32 - logL);
                const unsigned fm = (unsigned)(L - (int)f) & (unsigned)(L - 1);
                const unsigned pm = __brev(fm) >> (32 - logL);
                const double2 zf = xs[PADI(p)], zm = xs[PADI((int)pm)];
                const double2 ch = make_double2(0.5 * (zf.x + zm.x), 0.5 * (zf.y - zm.y));       // C^(f)
                const double2 uh = make_double2(0.5 * (zf.y + zm.y), -0.5 * (zf.x - zm.x));      // U^(f) * wsc
                s_t[p] = make_double2(ch.x * invL, -ch.y * invL);
                s_l[p] = make_double2(ch.x * sl, ch.y * sl);
                const double sg = pass == 0 ? 1.0 : -1.0;
                acc[k].x += sg * (ch.x * uh.x + ch.y * uh.y);                                     // conj(C^) U^
                acc[k].y += sg * (ch.x * uh.y - ch.y * uh.x);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < FFT_ACC; ++k) {
        const int p = tid + k * FFT_THREADS;
        if (p < L) xs[PADI(p)] = acc[k];
    }
    __syncthreads();
    fft_dit_inverse_inplace(xs, L, logL, W, tid);
    const double sc = invL * ig0 / wsc;
    for (int d = tid; d < n; d += FFT_THREADS) {
        const double v = xs[PADI(d)].x * sc;
        sKinv[d] = d == 0 ? v : 2.0 * v;
    }
}

static int ilog2i(int L) { int l = 0; while ((1 << l) < L) ++l; return l; }

int toeplitz_inv_init() {
    static int done = -1;
    if (done >= 0) return done;
    GPHM_CUDA_OK(cudaFuncSetAttribute(gs_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)fft_smem_bytes(FFT_MAX_L)));
    done = GPHM_OK;
    return done;
}

int launch_schur_levinson(const double* tabK, long long sTab, int n, double jitter, double* g, long long sG,
                          double* half_logdet, long long sLd, int* status, long long sStatus, double* gkap, long long sKap,
                          int* prog, long long sProg, int nsys, cudaStream_t st, long long* dbg) {
    if (n < 1 || n > SCHUR_MAX_N) { set_last_error("schur: n=%d outside [1,%d]", n, SCHUR_MAX_N); return GPHM_EINVAL; }
    if (nsys < 1 || nsys > 2) { set_last_error("schur: nsys=%d", nsys); return GPHM_EINVAL; }
    const int threads = std::min(SCHUR_MAX_THREADS, ((n + SCHUR_EPT - 1) / SCHUR_EPT + 31) / 32 * 32);
    for (int s = 0; s < nsys; ++s) GPHM_CUDA_OK(cudaMemsetAsync(prog + s * sProg, 0, sizeof(int), st));
    {
        LaunchScope scope(CAT_CHOL_DIAG, st, 8.0 * (double)n * n * nsys);
        schur_levinson_kernel<<<2 * nsys, threads, 0, st>>>(tabK, sTab, n, jitter, g, sG, half_logdet, sLd, status, sStatus,
                                                              gkap, sKap, prog, sProg, dbg);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_gs_prepare(const double* g, long long sG, int n, int L, const double* W, double* spec, long long sSpec,
                      double* sKinv, long long sS, int nsys, cudaStream_t st) {
    GPHM_TRY(toeplitz_inv_init());
    if (L > FFT_MAX_L || L < 2 * n) { set_last_error("gs_prepare: L=%d does not fit n=%d", L, n); return GPHM_EINVAL; }
    {
        LaunchScope scope(CAT_FFT, st);
        gs_prepare_kernel<<<nsys, FFT_THREADS, fft_smem_bytes(L), st>>>(
            g, sG, n, L, ilog2i(L), reinterpret_cast<const double2*>(W), reinterpret_cast<double2*>(spec), sSpec / 2, sKinv, sS);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
