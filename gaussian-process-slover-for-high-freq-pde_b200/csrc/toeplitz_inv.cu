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
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "fft_core.cuh"
#include <cooperative_groups.h>

namespace gphm {

constexpr int SCHUR_EPT = 8;                    // consecutive positions per thread = steps per batch
constexpr int SCHUR_MAX_THREADS = 512;
constexpr int SCHUR_MAX_N = SCHUR_EPT * SCHUR_MAX_THREADS;
constexpr int SCHUR_WCHUNK = 32 * SCHUR_EPT;    // positions per warp
constexpr int SCHUR_BPW = 32;                   // batches led by one warp (= its lanes)
constexpr int SCHUR_RING = 32;                  // boundary values kept per warp (4 batches)
constexpr int SCHUR_PUBLISH = 8;                // led batches per hand-over to the lattice CTA (64 coefficients)

constexpr int kSchurDefaultSplit = 4;           // CTAs per role of the recursion (schur_levinson_split_kernel); GPHM_SCHUR_SPLIT overrides
                                                // measured at n = 4096: 1 CTA 0.494 ms, 2 CTAs 0.384 ms, 4 CTAs 0.364 ms (same bits out)

int toeplitz_inv_max_n() { return SCHUR_MAX_N; }

// min_k (1 - kappa_k^2) below which the Toeplitz inverse-generator route is declared ill-conditioned.  Measured error of
// K^-1 v on that route ~ 7e-13 / min(1 - kappa^2) (tools/cond_guard_study.py): 3.5e-5 keeps it below 2e-8, a factor 50
// under the 1e-6 parity bound.  GPHM_GS_GUARD_MIN overrides (0 disables the guard).
double toeplitz_guard_min() {
    static const double v = [] { const char* e = getenv("GPHM_GS_GUARD_MIN"); return e ? atof(e) : 3.5e-5; }();
    return v;
}

// Two CTAs per system: the GENERATOR CTA runs the Schur recursion (it alone carries the serial
// dependency kappa_j -> kappa_{j+1}) and hands the reflection coefficients over through global memory;
// the LATTICE CTA consumes them (no feedback) and builds A_{n-1}.
//
// Register layout (both roles): thread t owns positions 8t .. 8t+7 of
//   generator:  be[p]  = beta_{j-1}[p]      second generator row             (live for p >= j)
//               A[p]   = alpha_{j-1}[p-1]   first generator row, pre-shifted (the recursion shifts it by one
//                                           position per step; keeping it shifted puts the pair that defines
//                                           kappa_j = -beta_{j-1}[j] / alpha_{j-1}[j-1] into one thread)
//   lattice:    a[p]   = A_{j-1}[p],  B[p] = B_{j-1}[p-1]                    (non-zero for p <= j)
// Step j:  alpha_j = A + kappa_j be,  beta_j = be + kappa_j A;   A_j = a + kappa_j B,  B_j = B + kappa_j a,
// then the shifted sequence moves up by one position.  The shift costs no register moves: steps are
// unrolled by 8 and at unrolled slot I the logical entry i of a shifted array lives in the physical
// register (i - I) mod 8, updates are in place, and only the entry that leaves the thread travels
// (warp shuffle; across warps through a ring in shared memory).
//
// Schedule.  All information flows upward (to higher positions): a warp needs the coefficients and the
// boundary values of the warp below it, nothing else.  Steps are grouped in batches of 8 (batch m =
// steps 8m .. 8m+7, whose coefficients all come from thread m's own eight positions) and time in
// periods separated by ONE block barrier: warp w processes batch m in period m + w.  The warp that
// owns thread m LEADS batch m: it needs nothing from the other warps, so inside the batch the
// coefficient chain (five dependent FP64 operations per step, see neg_div) runs back to back in the
// owner lane's registers, with no barrier, shared-memory round trip or shuffle on it; the coefficient
// reaches the warp's other lanes by shuffle and the other warps through shared memory, which they read
// one or more periods later.  (The earlier one-barrier-per-step kernel spent 405 cycles per step:
// 114 in store -> barrier -> load, ~200 in the owner warp's in-order issue, tools/ubench/lat.cu.)
// The schedule is a closed form - no flags, no spinning inside the CTA.

// -num/den to ~1 ulp in FIVE dependent FP64 operations after den is known (a division is seven):
//   r0 = MUFU.RCP64H(den) (2^-23), e = 1 - den r0, w = e + e^2, q0 = -num r0 (parallel to e),
//   result = q0 + q0 w = -num r0 (1 + e + e^2) = -num/den (1 - e^3).
// Straight-line (no slow path); every lane evaluates it, only the owner's value is used (other lanes may
// produce Inf/NaN harmlessly).
__device__ __forceinline__ double neg_div(double num, double den) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
    const double e = fma(-den, r0, 1.0);
    const double q0 = -num * r0;
    const double w = fma(e, e, e);
    return fma(q0, w, q0);
}

// One step of a warp on a pair of (shifted, static) arrays: S <- S + kp * T,  T <- T + kp * S_old, then S moves up.
// Generator: S = A (alpha), T = be.   Lattice: S = B, T = a.   ring: boundary values of every warp.
// Guarded form (branches): only the last, partial batch of an n that is not a multiple of 8 runs it.
template <int I>
__device__ __forceinline__ void pair_step(int j, double kp, int lane, int warp, double (&S)[SCHUR_EPT], double (&T)[SCHUR_EPT],
                                          double (*ring)[SCHUR_RING]) {
    constexpr int E = SCHUR_EPT;
    constexpr int OUT = (E - 1 - I) % E;       // physical slot of logical entry E-1 (leaves the thread)
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int ph = (i - I + E) % E;
        const double sv = S[ph], tv = T[i];
        S[ph] = fma(kp, tv, sv);
        T[i] = fma(kp, sv, tv);
    }
    const double out = S[OUT];
    double up = __shfl_up_sync(0xffffffffu, out, 1);
    if (lane == 31) ring[warp][j & (SCHUR_RING - 1)] = out;
    if (lane == 0) up = warp > 0 ? ring[warp - 1][j & (SCHUR_RING - 1)] : 0.0;     // written at least one period ago
    S[OUT] = up;                               // logical entry 0 of the next step
}

template <int I>
__device__ __forceinline__ void bulk_tail(int j0, int n, int lane, int warp, double (&S)[SCHUR_EPT], double (&T)[SCHUR_EPT],
                                          const double* kap, double (*ring)[SCHUR_RING]) {
    if constexpr (I < SCHUR_EPT) {
        if (j0 + I < n) pair_step<I>(j0 + I, kap[j0 + I], lane, warp, S, T, ring);
        bulk_tail<I + 1>(j0, n, lane, warp, S, T, kap, ring);
    }
}

template <int I>
__device__ __forceinline__ void lead_tail(int j0, int n, int lane, int warp, int ol, double cand, double (&A)[SCHUR_EPT],
                                          double (&be)[SCHUR_EPT], double* kap, double* gkap, double (*ring)[SCHUR_RING]) {
    if constexpr (I < SCHUR_EPT) {
        const int j = j0 + I;
        if (j < n) {
            if (lane == ol) { kap[j] = cand; gkap[j] = cand; }
            double next = 0.0;
            if constexpr (I + 1 < SCHUR_EPT) next = neg_div(fma(cand, A[1], be[I + 1]), fma(cand, be[I], A[0]));
            const double kp = __shfl_sync(0xffffffffu, cand, ol);
            pair_step<I>(j, kp, lane, warp, A, be, ring);
            lead_tail<I + 1>(j0, n, lane, warp, ol, next, A, be, kap, gkap, ring);
        }
    }
}

// A FULL batch (eight steps j0 .. j0+7) of one warp as straight-line code - no branch, so that the eight
// steps form one basic block and the scheduler overlaps the shuffles, the shared-memory accesses and, in
// the leading warp, the coefficient chain of step j+1 with the sixteen FMAs of step j (as separate
// blocks they ran back to back in the in-order issue: ~430 cycles per step, 3.4x the FP64-pipe time).
//   rin[8]   the eight values that enter lane 0 during the batch (the warp below wrote them a period ago;
//            don't-care in a leading warp - they only reach dead positions)
//   rout     this warp's eight ring slots for the values that leave lane 31
//   LEAD     lane ol owns positions j0 .. j0+7: it produces kappa_{j+1} = -beta_j[j+1] / alpha_j[j] from its
//            own registers one step ahead (logical entries I, I+1 = physical 0, 1) and the coefficient
//            reaches the other lanes by shuffle - off the chain;   else kp[8] holds the coefficients.
template <int I, bool LEAD>
__device__ __forceinline__ void batch_full(double cand, int ol, bool own, bool lane0, bool lane31, const double (&kp8)[SCHUR_EPT],
                                           const double (&rin)[SCHUR_EPT], double* __restrict__ rout, double* __restrict__ kout,
                                           double* __restrict__ gkout, double (&S)[SCHUR_EPT], double (&T)[SCHUR_EPT]) {
    constexpr int E = SCHUR_EPT;
    if constexpr (I < E) {
        constexpr int OUT = (E - 1 - I) % E;
        double kp, next = 0.0;
        if constexpr (LEAD) {
            if (own) { kout[I] = cand; gkout[I] = cand; }
            if constexpr (I + 1 < E) next = neg_div(fma(cand, S[1], T[I + 1]), fma(cand, T[I], S[0]));
            kp = __shfl_sync(0xffffffffu, cand, ol);
        } else {
            kp = kp8[I];
        }
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int ph = (i - I + E) % E;
            const double sv = S[ph], tv = T[i];
            S[ph] = fma(kp, tv, sv);
            T[i] = fma(kp, sv, tv);
        }
        const double out = S[OUT];
        const double up = __shfl_up_sync(0xffffffffu, out, 1);
        if (lane31) rout[I] = out;
        S[OUT] = lane0 ? rin[I] : up;
        batch_full<I + 1, LEAD>(next, ol, own, lane0, lane31, kp8, rin, rout, kout, gkout, S, T);
    }
}

__device__ __forceinline__ void load8(double (&v)[SCHUR_EPT], const double* p) {       // 64-byte aligned
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; i += 2) {
        const double2 t = *reinterpret_cast<const double2*>(p + i);
        v[i] = t.x; v[i + 1] = t.y;
    }
}

// blockIdx.x = 2 * system + role (0: generator, 1: lattice).  gkap[n] / prog[1] per system: hand-over buffer and
// the number of valid coefficients in it (zeroed by the launcher).
__global__ void __launch_bounds__(SCHUR_MAX_THREADS, 1)
schur_levinson_kernel(const double* __restrict__ tab, long long sTab, int n, double jitter, double* __restrict__ g,
                      long long sG, double* __restrict__ half_logdet, long long sLd, int* __restrict__ status,
                      long long sStatus, double* gkap, long long sKap, int* prog, long long sProg, int* guard, int guard_bit0,
                      double guard_min, long long* dbg, const int* __restrict__ skip) {
    if (skip && *skip) return;
    const int sys = blockIdx.x >> 1, role = blockIdx.x & 1;
    const long long t_start = clock64();
    tab += sys * sTab; g += sys * sG; half_logdet += sys * sLd; status += sys * sStatus; gkap += sys * sKap; prog += sys * sProg;
    __shared__ __align__(16) double kap[SCHUR_MAX_N];
    __shared__ __align__(16) double ring[SCHUR_MAX_THREADS / 32][SCHUR_RING];
    __shared__ double red[34];
    __shared__ int bad, s_have;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int j0t = tid * SCHUR_EPT;
    const int nb = (n + SCHUR_EPT - 1) / SCHUR_EPT;            // batches
    const double r0 = tab[0] + jitter;
    if (role == 0) {
        // ---- generator: Schur recursion ----
        double A[SCHUR_EPT], be[SCHUR_EPT];
        // State "before step 0": unshifted alpha_0 = r, beta_0 = (0, r_1, ...).  Step 0 runs with kappa_0 = 0:
        // it changes no value and performs the first shift, so that every step is identical.
#pragma unroll
        for (int i = 0; i < SCHUR_EPT; ++i) {
            const int p = j0t + i;
            const double rp = (p < n) ? (p == 0 ? r0 : tab[p]) : 0.0;
            A[i] = rp;
            be[i] = (p == 0) ? 0.0 : rp;
        }
        if (tid == 0) bad = 0x7fffffff;
        __syncthreads();
        int led = 0;                                               // batches led so far (all threads agree)
        long long c_lead = 0, c_bulk = 0, c_wait = 0, n_lead = 0, n_bulk = 0;   // cycle counters (dbg only)
        for (int T = 0; T < nb + nwarps - 1; ++T) {
            const int m = T - warp, j0 = m * SCHUR_EPT;
            const bool live = m >= 0 && m < nb && m < SCHUR_BPW * (warp + 1);   // else: not started yet / every position of the warp is dead
            const bool lead = live && m / SCHUR_BPW == warp;
            const long long t0 = dbg ? clock64() : 0;
            long long t1 = t0;
            // votes: the branch conditions are warp-uniform, and the compiler has to know it (no divergence handling around the shuffles)
            if (__all_sync(0xffffffffu, live)) {
                const double cand = j0 == 0 ? 0.0 : neg_div(be[0], A[0]);          // kappa_{8m} (leading warp): slot 0, logical 0 = physical 0
                if (__all_sync(0xffffffffu, j0 + SCHUR_EPT <= n)) {
                    const int rs = j0 & (SCHUR_RING - 1);
                    double rin[SCHUR_EPT], kp8[SCHUR_EPT];
                    load8(rin, &ring[warp > 0 ? warp - 1 : 0][rs]);
                    if (__all_sync(0xffffffffu, lead)) {
                        batch_full<0, true>(cand, m % SCHUR_BPW, lane == m % SCHUR_BPW, lane == 0, lane == 31, kp8, rin, &ring[warp][rs],
                                            kap + j0, gkap + j0, A, be);
                        if (dbg) { t1 = clock64(); c_lead += t1 - t0; ++n_lead; }
                    } else {
                        load8(kp8, kap + j0);
                        batch_full<0, false>(0.0, 0, false, lane == 0, lane == 31, kp8, rin, &ring[warp][rs], nullptr, nullptr, A, be);
                        if (dbg) { t1 = clock64(); c_bulk += t1 - t0; ++n_bulk; }
                    }
                } else if (__all_sync(0xffffffffu, lead)) {
                    lead_tail<0>(j0, n, lane, warp, m % SCHUR_BPW, cand, A, be, kap, gkap, ring);
                } else {
                    bulk_tail<0>(j0, n, lane, warp, A, be, kap, ring);
                }
            }
            __syncthreads();
            if (dbg) c_wait += clock64() - t1;
            // batch `led` is led in period led + led / 32; hand over every SCHUR_PUBLISH led batches and at the end
            if (led < nb && led + led / SCHUR_BPW == T) {
                ++led;
                if ((led % SCHUR_PUBLISH == 0 || led == nb) && tid == blockDim.x - 1) {
                    __threadfence();                               // cumulative: the owners' gkap stores were ordered by the barrier
                    *reinterpret_cast<volatile int*>(prog) = min(led * SCHUR_EPT, n);
                }
            }
        }
        // log|K| = n log r0 + sum_k (n - k) log(1 - kappa_k^2); first |kappa| >= 1 <=> first non-positive prediction error
        // Conditioning guard (the Gohberg-Semencul formula divides a difference of two triangular products by g0 and
        // loses ~7e-13 / min_k(1 - kappa_k^2) of relative accuracy, tools/cond_guard_study.py): flag the system when
        // that margin falls below guard_min so that the host moves the plan to the Cholesky path.
        double lsum = 0.0, gmin = 1.0;
#pragma unroll
        for (int i = 0; i < SCHUR_EPT; ++i) {
            const int k = j0t + i;
            if (k >= 1 && k < n) {
                const double kp = kap[k];
                if (!(fabs(kp) < 1.0)) atomicMin(&bad, k);         // also catches NaN
                lsum += (double)(n - k) * log1p(-kp * kp);
                gmin = fmin(gmin, (1.0 - kp) * (1.0 + kp));
            }
        }
        if (!(r0 > 0.0) && tid == 0) atomicMin(&bad, 0);
        const double ltot = block_sum(lsum, red);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmin = fmin(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        if (lane == 0 && gmin < guard_min && guard) atomicOr(guard, 1 << (guard_bit0 + sys));
        if (dbg && lane == 0) {
            unsigned long long* d = reinterpret_cast<unsigned long long*>(dbg);
            atomicAdd(d + 8, (unsigned long long)c_lead); atomicAdd(d + 9, (unsigned long long)n_lead);
            atomicAdd(d + 10, (unsigned long long)c_bulk); atomicAdd(d + 11, (unsigned long long)n_bulk);
            atomicAdd(d + 12, (unsigned long long)c_wait);
        }
        if (tid == 0) {
            half_logdet[0] = 0.5 * ((double)n * log(r0) + ltot);
            if (bad != 0x7fffffff) status[0] = bad + 1;            // like a Cholesky pivot index
            if (dbg) dbg[blockIdx.x] = clock64() - t_start;
        }
        return;
    }
    // ---- lattice: A_j = A_{j-1} + kappa_j z B_{j-1},  B_j = z B_{j-1} + kappa_j A_{j-1};  warp 0 always leads ----
    double a[SCHUR_EPT], B[SCHUR_EPT];
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) { a[i] = (j0t + i == 0) ? 1.0 : 0.0; B[i] = a[i]; }
    int have = 0;
    bool dead = false;
    for (int T = 0; T < nb + nwarps - 1; ++T) {
        const int need = min((T + 1) * SCHUR_EPT, n);              // warp 0 processes batch T now; the others older ones
        if (need > have) {                                         // uniform
            if (tid == 0) {
                int v = 0;
                long long spins = 0;
                while ((v = *reinterpret_cast<volatile int*>(prog)) < need && spins < (1ll << 24)) { __nanosleep(100); ++spins; }
                s_have = v;
            }
            __syncthreads();
            const int now = s_have;
            if (now < need) { dead = true; break; }                // the producer never arrived (cannot happen when both CTAs run)
            for (int i = have + tid; i < now; i += blockDim.x) kap[i] = __ldcg(gkap + i);
            __syncthreads();
            have = now;
        }
        const int m = T - warp, j0 = m * SCHUR_EPT;
        // positions of this warp become non-zero at step 256 w - 1: earlier batches would only move zeros
        const bool live = m >= 0 && m < nb && (m + 1) * SCHUR_EPT >= warp * SCHUR_WCHUNK;
        if (__all_sync(0xffffffffu, live)) {
            if (__all_sync(0xffffffffu, j0 + SCHUR_EPT <= n)) {
                const int rs = j0 & (SCHUR_RING - 1);
                double rin[SCHUR_EPT], kp8[SCHUR_EPT];
                load8(rin, &ring[warp > 0 ? warp - 1 : 0][rs]);
                load8(kp8, kap + j0);
                if (warp == 0) {
#pragma unroll
                    for (int i = 0; i < SCHUR_EPT; ++i) rin[i] = 0.0;                  // B[-1] = 0
                }
                batch_full<0, false>(0.0, 0, false, lane == 0, lane == 31, kp8, rin, &ring[warp][rs], nullptr, nullptr, B, a);
            } else {
                bulk_tail<0>(j0, n, lane, warp, B, a, kap, ring);
            }
        }
        __syncthreads();
    }
    // g = A_{n-1} / E_{n-1},  E_{n-1} = r0 * prod_k (1 - kappa_k^2)   (a does not shift: entry i is a[i])
    double prod = 1.0;
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) {
        const int k = j0t + i;
        if (k >= 1 && k < n) { const double kp = kap[k]; prod *= (1.0 - kp) * (1.0 + kp); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, o);
    __syncthreads();
    if (lane == 0) red[warp] = prod;
    __syncthreads();
    double E = r0;
    for (int w = 0; w < nwarps; ++w) E *= red[w];
    const double invE = dead ? __longlong_as_double(0x7ff8000000000000ll) : 1.0 / E;
    if (dead && tid == 0 && guard) atomicOr(guard, 1 << (8 + guard_bit0 + sys));      // distinct status: GPHM_STALLED
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) { const int p = j0t + i; if (p < n) g[p] = a[i] * invE; }
    if (dbg && tid == 0) dbg[blockIdx.x] = clock64() - t_start;
}

// ---------------------------------------------------------------------------------------------------------------------
// The same recursion with every role spread over NS CTAs (round 2).  CTA c of a role owns the positions
// [c n/NS, (c+1) n/NS) - the warps 32 c n/(256 NS) ... of the single-CTA wavefront - and runs its own period schedule;
// what crosses a CTA boundary travels through global memory behind a progress counter, exactly like the reflection
// coefficients always did between the generator and the lattice CTA:
//     gkap[n]  + prog[0]                      coefficients kappa_j, published by whichever generator CTA currently leads
//     gbnd[role][c][n] + prog[1 + 3 role + c] the values that leave the top position of CTA c (eight per batch), for CTA c+1
// All information flows upward, so there is no feedback between CTAs and a consumer only ever lags.  Why: on ONE SM the
// lattice is bound by the FP64 pipe late in the recursion (all 16 warps live: 236 cycles per step at n = 4096, 2 x 4096 FMAs
// per step on 64 lanes = 128) and the generator early; split, every CTA stays under the coefficient chain's ~95 cycles per
// step.  Requires n to be a multiple of 256 NS (full batches, full warps); the launcher falls back to the single-CTA
// kernel otherwise.  blockIdx.x = system * 2 NS + role * NS + c; the 2 NS CTAs of a system form one thread-block cluster
// (co-scheduled: the consumers spin on their producers).
// Threads: n / (8 NS) workers + ONE PUBLISHER WARP.  Handing a counter over costs a device-scope fence (~750 cycles measured
// here); issued by a worker it sat between two block barriers of the leading CTA - 17 % of its time.  The workers now only
// post the number of finished periods in shared memory after their barrier and the publisher warp, which owns no positions and joins
// no barrier of the recursion (named barrier 1 counts the workers only), fences and stores them.  The generator runs TWO
// batches per barrier period (the ring holds four: two being read by the warp above, two being written), which halves
// what a period costs besides the coefficient chain (barrier, votes, ring loads: ~250 of 1100 cycles).
__global__ void __launch_bounds__(SCHUR_MAX_THREADS / 2 + 32, 1)      // NS >= 2: at most 256 workers
schur_levinson_split_kernel(const double* __restrict__ tab, long long sTab, int n, double jitter, double* __restrict__ g,
                            long long sG, double* __restrict__ half_logdet, long long sLd, int* __restrict__ status,
                            long long sStatus, double* gkap, long long sKap, int* prog, long long sProg, double* gbnd,
                            long long sBnd, int ns, int* guard, int guard_bit0, double guard_min, const int* __restrict__ skip,
                            long long* dbg) {
    if (skip && *skip) return;
    const long long t_begin = dbg ? clock64() : 0;
    long long t_wait = 0;
    extern __shared__ __align__(128) double sm_split[];
    double* kap = sm_split;                                   // [SCHUR_MAX_N]
    double* bin = kap + SCHUR_MAX_N;                          // [SCHUR_MAX_N] values entering this CTA's lowest position
    double (*ring)[SCHUR_RING] = reinterpret_cast<double (*)[SCHUR_RING]>(bin + SCHUR_MAX_N);      // [16][32]
    double* red = reinterpret_cast<double*>(ring + SCHUR_MAX_THREADS / 32);                          // [34]
    int* s_i = reinterpret_cast<int*>(red + 34);                                                    // [8]
    volatile int* s_pub = s_i + 4;                            // [0] periods the workers have finished, [2] abort
    const int per_sys = 2 * ns, sys = blockIdx.x / per_sys, rr = blockIdx.x % per_sys, role = rr / ns, c = rr % ns;
    tab += sys * sTab; g += sys * sG; half_logdet += sys * sLd; status += sys * sStatus; gkap += sys * sKap; prog += sys * sProg;
    gbnd += sys * sBnd;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nworkers = blockDim.x - 32;
    const int nw = nworkers >> 5, gw0 = c * nw, gw = gw0 + warp;
    const int j0t = (gw * 32 + lane) * SCHUR_EPT;
    const int nb = n / SCHUR_EPT;
    const double r0 = tab[0] + jitter;
    volatile int* progK = prog;
    volatile int* progIn = c > 0 ? prog + 1 + role * 3 + (c - 1) : nullptr;
    int* progOut = c < ns - 1 ? prog + 1 + role * 3 + c : nullptr;
    const double* bndIn = c > 0 ? gbnd + (size_t)(role * 3 + c - 1) * n : nullptr;
    double* bndOut = c < ns - 1 ? gbnd + (size_t)(role * 3 + c) * n : nullptr;
    const bool top = warp == nw - 1;
    const int lead_lo = SCHUR_BPW * gw0, lead_hi = min(nb, SCHUR_BPW * (gw0 + nw));          // batches the generator CTA leads
    auto wsync = [&]() { asm volatile("bar.sync 1, %0;" :: "r"(nworkers) : "memory"); };     // the workers' barrier

    if (tid >= nworkers) {
        // ---- publisher warp: hands over the LATEST counts whenever they have moved (a fence takes longer than a lattice period) ----
        if (role == 1 && !progOut) return;                         // the top lattice CTA hands nothing over
        if (tid == nworkers) { s_pub[0] = 0; s_pub[2] = 0; }
        __syncthreads();
        if (lane == 0) {
            // generator: batch m is led by warp m / 32 - gw0 in period m / 2 + that warp, the top warp finishes batches 2 (T - nw + 1), +1
            // in period T;  lattice: the top warp finishes batch T - nw + 1 in period T
            int seen = 0, led = lead_lo, pk = lead_lo, pb = 0;       // periods seen; batches led and finished; published so far
            const int endK = role == 0 ? lead_hi : lead_lo, endB = role == 0 ? lead_hi : nb;
            for (long long spins = 0; spins < (1ll << 28); ++spins) {
                const int per = s_pub[0], ab = s_pub[2];
                if (per != seen) {
                    seen = per;
                    const int T = per - 1;
                    if (role == 0) while (led < lead_hi && led / 2 + (led / SCHUR_BPW - gw0) <= T) ++led;
                    const int vk = led;
                    const int vb = !progOut ? pb : max(pb, min(role == 0 ? 2 * (T - (nw - 1)) + 2 : T - (nw - 1) + 1, endB));
                    if (vk != pk || vb != pb) {
                        __threadfence();                           // cumulative over the workers' stores (ordered by their barrier)
                        if (vk != pk) *reinterpret_cast<volatile int*>(prog) = vk * SCHUR_EPT;
                        if (vb != pb) *reinterpret_cast<volatile int*>(progOut) = vb;
                        pk = vk; pb = vb;
                    }
                }
                if (ab || (pk == endK && (!progOut || pb == endB))) break;
            }
        }
        if (role == 1) return;
        if (c != ns - 1) return;
    }
    int haveK = 0, haveB = 0;
    bool dead = false;
    // waits (uniform over the workers) until needK coefficients and needB boundary batches have been published, then copies the new ones
    auto fetch = [&](int needK, int needB) -> bool {
        if (needK <= haveK && needB <= haveB) return true;
        if (tid == 0) {
            int vK = haveK, vB = haveB;
            long long spins = 0;
            const long long tw = dbg ? clock64() : 0;
            while (spins < (1ll << 24)) {
                vK = *progK; vB = progIn ? *progIn : needB;
                if (vK >= needK && vB >= needB) break;
                ++spins;                                           // no __nanosleep: a hand-over then took ~12 k cycles instead of ~3 k
            }
            if (dbg) t_wait += clock64() - tw;
            s_i[0] = vK; s_i[1] = vB;
        }
        wsync();
        const int nowK = min(s_i[0], n), nowB = min(s_i[1], nb);
        if (nowK < needK || nowB < needB) return false;            // a producer never arrived
        for (int i = haveK + tid; i < nowK; i += nworkers) kap[i] = __ldcg(gkap + i);
        if (bndIn) for (int i = haveB * SCHUR_EPT + tid; i < nowB * SCHUR_EPT; i += nworkers) bin[i] = __ldcg(bndIn + i);
        wsync();
        haveK = max(haveK, nowK); haveB = max(haveB, nowB);
        return true;
    };
    double rin[SCHUR_EPT], kp8[SCHUR_EPT];

    if (role == 0) {
        // ---- generator ----
        if (tid < nworkers) {
            double A[SCHUR_EPT], be[SCHUR_EPT];
#pragma unroll
            for (int i = 0; i < SCHUR_EPT; ++i) {
                const int p = j0t + i;
                const double rp = p == 0 ? r0 : tab[p];
                A[i] = rp;
                be[i] = p == 0 ? 0.0 : rp;
            }
            double cand = 0.0;                                     // kappa_0 = 0
            __syncthreads();                                       // with the publisher: s_pub is initialised
            const int nper = lead_hi / 2 + nw - 1;                 // lead_hi is a multiple of 32
            for (int T = 0; T < nper; ++T) {
                if (c > 0) {       // coefficients and boundary values of the batches the CTAs below lead
                    const int nbat = min(2 * (T + 1), lead_lo);
                    if (!fetch(nbat * SCHUR_EPT, nbat)) { dead = true; break; }
                }
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const int m = 2 * (T - warp) + h, j0 = m * SCHUR_EPT;
                    const bool live = m >= 0 && m < nb && m < SCHUR_BPW * (gw + 1);
                    const bool lead = live && (m / SCHUR_BPW) == gw;
                    if (__all_sync(0xffffffffu, live)) {
                        const int rs = j0 & (SCHUR_RING - 1);
                        if (warp > 0) load8(rin, &ring[warp - 1][rs]);
                        else if (c > 0 && m < lead_lo) load8(rin, bin + j0);
                        else {
#pragma unroll
                            for (int i = 0; i < SCHUR_EPT; ++i) rin[i] = 0.0;                // only reaches dead positions
                        }
                        if (__all_sync(0xffffffffu, lead)) {
                            batch_full<0, true>(cand, m % SCHUR_BPW, lane == m % SCHUR_BPW, lane == 0, lane == 31, kp8, rin, &ring[warp][rs],
                                                kap + j0, gkap + j0, A, be);
                        } else {
                            load8(kp8, kap + j0);
                            batch_full<0, false>(0.0, 0, false, lane == 0, lane == 31, kp8, rin, &ring[warp][rs], nullptr, nullptr, A, be);
                        }
                        cand = neg_div(be[0], A[0]);       // first coefficient of the next batch (meaningful in its owner lane), same
                                                           // basic block as the batch: its chain overlaps the last step's updates
                        if (bndOut && top && lane == 31) {
#pragma unroll
                            for (int i = 0; i < SCHUR_EPT; ++i) bndOut[j0 + i] = ring[warp][rs + i];
                        }
                    }
                }
                wsync();
                if (tid == 0) { __threadfence_block(); s_pub[0] = T + 1; }                   // periods finished, for the publisher warp
            }
            if (dead && tid == 0) s_pub[2] = 1;
        }
        if (dbg && tid == 0) { dbg[2 * blockIdx.x] = clock64() - t_begin; dbg[2 * blockIdx.x + 1] = t_wait; }
        if (c != ns - 1) return;
        // the last generator CTA holds every coefficient: log|K|, first non-positive prediction error, conditioning guard
        if (tid == 0) s_i[2] = 0x7fffffff;
        __syncthreads();                                           // every thread of the CTA, publisher warp included
        double lsum = 0.0, gmin = 1.0;
        for (int k = 1 + tid; k < n; k += blockDim.x) {
            const double kp = kap[k];
            if (!(fabs(kp) < 1.0)) atomicMin(&s_i[2], k);
            lsum += (double)(n - k) * log1p(-kp * kp);
            gmin = fmin(gmin, (1.0 - kp) * (1.0 + kp));
        }
        if (!(r0 > 0.0) && tid == 0) atomicMin(&s_i[2], 0);
        const double ltot = block_sum(lsum, red);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmin = fmin(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        if (lane == 0 && gmin < guard_min && guard) atomicOr(guard, 1 << (guard_bit0 + sys));
        if (tid == 0) {
            const bool aborted = s_pub[2] != 0;
            half_logdet[0] = aborted ? __longlong_as_double(0x7ff8000000000000ll) : 0.5 * ((double)n * log(r0) + ltot);
            if (s_i[2] != 0x7fffffff) status[0] = s_i[2] + 1;
            if (aborted && guard) atomicOr(guard, 1 << (8 + guard_bit0 + sys));
        }
        return;
    }
    // ---- lattice (workers only: the publisher warp has left) ----
    double a[SCHUR_EPT], B[SCHUR_EPT];
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) { a[i] = (j0t + i == 0) ? 1.0 : 0.0; B[i] = a[i]; }
    const int mstart = max(0, SCHUR_BPW * gw0 - 1);            // warp 0 of this CTA becomes non-zero at step 256 gw0 - 1
    haveB = mstart;                                            // older boundary batches are zeros nobody reads (nor writes)
    if (progOut) __syncthreads();                              // with the publisher: s_pub is initialised
    for (int T = mstart; T < nb + nw - 1; ++T) {
        if (!fetch(min((T + 1) * SCHUR_EPT, n), c > 0 ? min(T + 1, nb) : 0)) { dead = true; break; }
        const int m = T - warp, j0 = m * SCHUR_EPT;
        const bool live = m >= 0 && m < nb && (m + 1) * SCHUR_EPT >= gw * SCHUR_WCHUNK;
        if (__all_sync(0xffffffffu, live)) {
            const int rs = j0 & (SCHUR_RING - 1);
            if (warp > 0) load8(rin, &ring[warp - 1][rs]);
            else if (c > 0) load8(rin, bin + j0);
            else {
#pragma unroll
                for (int i = 0; i < SCHUR_EPT; ++i) rin[i] = 0.0;                            // B[-1] = 0
            }
            load8(kp8, kap + j0);
            batch_full<0, false>(0.0, 0, false, lane == 0, lane == 31, kp8, rin, &ring[warp][rs], nullptr, nullptr, B, a);
            if (bndOut && top && lane == 31) {
#pragma unroll
                for (int i = 0; i < SCHUR_EPT; ++i) bndOut[j0 + i] = ring[warp][rs + i];
            }
        }
        wsync();
        if (progOut && tid == 0) { __threadfence_block(); s_pub[0] = T + 1; }                // periods finished, for the publisher warp
    }
    if (dead && progOut && tid == 0) s_pub[2] = 1;
    // g = A_{n-1} / E_{n-1},  E_{n-1} = r0 * prod_k (1 - kappa_k^2): every CTA forms the product over all k in the same order
    double prod = 1.0;
    if (!dead) for (int k = 1 + tid; k < n; k += nworkers) { const double kp = kap[k]; prod *= (1.0 - kp) * (1.0 + kp); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, o);
    wsync();
    if (lane == 0) red[warp] = prod;
    wsync();
    double E = r0;
    for (int w = 0; w < nw; ++w) E *= red[w];
    const double invE = dead ? __longlong_as_double(0x7ff8000000000000ll) : 1.0 / E;
    if (dead && tid == 0 && guard) atomicOr(guard, 1 << (8 + guard_bit0 + sys));
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) g[j0t + i] = a[i] * invE;
    if (dbg && tid == 0) { dbg[2 * blockIdx.x] = clock64() - t_begin; dbg[2 * blockIdx.x + 1] = t_wait; }
}

// The split recursion with the hand-over through DISTRIBUTED SHARED MEMORY instead of global memory: producers keep what they
// hand over (coefficients, boundary values, two counters) in their own shared memory and publish once per period with a
// cluster-scope fence; consumers poll the producer's counter (a ~200-cycle DSMEM load instead of an L2 round trip) and pull what
// is new.  Per-batch granularity, no gpu-scope fence, no flags in global memory.  Same arithmetic, same bits.
__global__ void __launch_bounds__(SCHUR_MAX_THREADS, 1)
schur_levinson_dsmem_kernel(const double* __restrict__ tab, long long sTab, int n, double jitter, double* __restrict__ g,
                            long long sG, double* __restrict__ half_logdet, long long sLd, int* __restrict__ status,
                            long long sStatus, double* gkap, long long sKap, int* prog, long long sProg, double* gbnd,
                            long long sBnd, int ns, int* guard, int guard_bit0, double guard_min, const int* __restrict__ skip) {
    if (skip && *skip) return;
    extern __shared__ __align__(128) double sm_split[];
    double* kap = sm_split;                                   // [SCHUR_MAX_N] coefficients (own + pulled from the leading CTA)
    double* bin = kap + SCHUR_MAX_N;                          // [SCHUR_MAX_N] values entering this CTA's lowest position (pulled)
    double* bout = bin + SCHUR_MAX_N;                         // [SCHUR_MAX_N] values leaving this CTA's top position (read remotely)
    double (*ring)[SCHUR_RING] = reinterpret_cast<double (*)[SCHUR_RING]>(bout + SCHUR_MAX_N);     // [16][32]
    double* red = reinterpret_cast<double*>(ring + SCHUR_MAX_THREADS / 32);                          // [34]
    int* s_i = reinterpret_cast<int*>(red + 34);                                                    // [4] scratch, [4] kdone, [5] bdone
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int per_sys = 2 * ns, sys = blockIdx.x / per_sys, rr = blockIdx.x % per_sys, role = rr / ns, c = rr % ns;
    tab += sys * sTab; g += sys * sG; half_logdet += sys * sLd; status += sys * sStatus; gkap += sys * sKap; prog += sys * sProg;
    gbnd += sys * sBnd;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x >> 5, gw0 = c * nw, gw = gw0 + warp;
    const int j0t = (gw * 32 + lane) * SCHUR_EPT;
    const int nb = n / SCHUR_EPT;
    const double r0 = tab[0] + jitter;
    // hand-over through distributed shared memory: a producer publishes into ITS OWN shared memory (kap / bout + the counters
    // kdone / bdone, one cluster-scope fence per period); a consumer polls the producer's counter and pulls what is new
    volatile int* kdone = s_i + 4;
    volatile int* bdone = s_i + 5;
    if (tid == 0) { *kdone = 0; *bdone = 0; }
    cluster.sync();                                            // every CTA of the system is resident and initialised
    const double* bndIn = c > 0 ? cluster.map_shared_rank(bout, role * ns + c - 1) : nullptr;
    const volatile int* progIn = c > 0 ? cluster.map_shared_rank(s_i + 5, role * ns + c - 1) : nullptr;
    const bool hasOut = c < ns - 1;
    const bool top = warp == nw - 1;
    const int bpc = SCHUR_BPW * nw;                            // batches led per generator CTA
    int haveK = 0, haveB = 0;
    bool dead = false;
    // waits (uniform) until needK coefficients and needB boundary batches have been published, then copies the new ones
    auto fetch = [&](int needK, int needB) -> bool {
        if (needK <= haveK && needB <= haveB) return true;
        // the generator CTA that leads coefficient needK - 1 holds every coefficient below it as well
        const int csrc = needK > haveK ? min(ns - 1, ((needK - 1) / SCHUR_EPT) / bpc) : 0;
        const double* rkap = cluster.map_shared_rank(kap, csrc);
        const volatile int* rkd = cluster.map_shared_rank(s_i + 4, csrc);
        if (tid == 0) {
            int vK = haveK, vB = haveB;
            long long spins = 0;
            while (spins < (1ll << 26)) {
                vK = needK > haveK ? *rkd : haveK; vB = (progIn && needB > haveB) ? *progIn : max(haveB, needB);
                if (vK >= needK && vB >= needB) break;
                ++spins;
            }
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
            s_i[0] = vK; s_i[1] = vB;
        }
        __syncthreads();
        const int nowK = min(s_i[0], n), nowB = min(s_i[1], nb);
        if (nowK < needK || nowB < needB) return false;            // a producer never arrived
        if (nowK > haveK && !(role == 0 && csrc == c))
            for (int i = haveK + tid; i < nowK; i += blockDim.x) kap[i] = rkap[i];
        if (bndIn) for (int i = haveB * SCHUR_EPT + tid; i < nowB * SCHUR_EPT; i += blockDim.x) bin[i] = bndIn[i];
        __syncthreads();
        haveK = max(haveK, nowK); haveB = max(haveB, nowB);
        return true;
    };
    // one thread, after the period barrier: everything the CTA wrote so far becomes visible cluster-wide before the counters move
    auto publish2 = [&](int kcount, int bcount) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        if (kcount >= 0) *kdone = kcount;
        if (bcount >= 0) *bdone = bcount;
    };
    double rin[SCHUR_EPT], kp8[SCHUR_EPT];

    if (role == 0) {
        // ---- generator ----
        double A[SCHUR_EPT], be[SCHUR_EPT];
#pragma unroll
        for (int i = 0; i < SCHUR_EPT; ++i) {
            const int p = j0t + i;
            const double rp = p == 0 ? r0 : tab[p];
            A[i] = rp;
            be[i] = p == 0 ? 0.0 : rp;
        }
        const int lead_lo = SCHUR_BPW * gw0, lead_hi = min(nb, SCHUR_BPW * (gw0 + nw));      // batches led (and the last ones processed) here
        int led = lead_lo;
        __syncthreads();
        for (int T = 0; T < lead_hi + nw - 1; ++T) {
            if (c > 0) {       // coefficients and boundary values of the batches the CTAs below lead
                const int nbat = min(T + 1, lead_lo);
                if (!fetch(nbat * SCHUR_EPT, nbat)) { dead = true; break; }
            }
            const int m = T - warp, j0 = m * SCHUR_EPT;
            const bool live = m >= 0 && m < nb && m < SCHUR_BPW * (gw + 1);
            const bool lead = live && (m / SCHUR_BPW) == gw;
            if (__all_sync(0xffffffffu, live)) {
                const double cand = j0 == 0 ? 0.0 : neg_div(be[0], A[0]);
                const int rs = j0 & (SCHUR_RING - 1);
                if (warp > 0) load8(rin, &ring[warp - 1][rs]);
                else if (c > 0 && m < lead_lo) load8(rin, bin + j0);
                else {
#pragma unroll
                    for (int i = 0; i < SCHUR_EPT; ++i) rin[i] = 0.0;                        // only reaches dead positions
                }
                if (__all_sync(0xffffffffu, lead)) {
                    batch_full<0, true>(cand, m % SCHUR_BPW, lane == m % SCHUR_BPW, lane == 0, lane == 31, kp8, rin, &ring[warp][rs],
                                        kap + j0, kap + j0, A, be);
                } else {
                    load8(kp8, kap + j0);
                    batch_full<0, false>(0.0, 0, false, lane == 0, lane == 31, kp8, rin, &ring[warp][rs], nullptr, nullptr, A, be);
                }
                if (hasOut && top && lane == 31) {
#pragma unroll
                    for (int i = 0; i < SCHUR_EPT; ++i) bout[j0 + i] = ring[warp][rs + i];
                }
            }
            __syncthreads();
            int kc = -1, bc = -1;
            if (led < lead_hi && led + (led / SCHUR_BPW - gw0) == T) { ++led; kc = led * SCHUR_EPT; }
            const int mt = T - (nw - 1);                       // batch the top warp finished in this period
            if (hasOut && mt >= 0 && mt < lead_hi) bc = mt + 1;
            if ((kc >= 0 || bc >= 0) && tid == blockDim.x - 1) publish2(kc, bc);
        }
        if (c != ns - 1) { cluster.sync(); return; }
        // the last generator CTA holds every coefficient: log|K|, first non-positive prediction error, conditioning guard
        if (tid == 0) s_i[2] = 0x7fffffff;
        __syncthreads();
        double lsum = 0.0, gmin = 1.0;
        for (int k = 1 + tid; k < n; k += blockDim.x) {
            const double kp = kap[k];
            if (!(fabs(kp) < 1.0)) atomicMin(&s_i[2], k);
            lsum += (double)(n - k) * log1p(-kp * kp);
            gmin = fmin(gmin, (1.0 - kp) * (1.0 + kp));
        }
        if (!(r0 > 0.0) && tid == 0) atomicMin(&s_i[2], 0);
        const double ltot = block_sum(lsum, red);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmin = fmin(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
        if (lane == 0 && gmin < guard_min && guard) atomicOr(guard, 1 << (guard_bit0 + sys));
        if (tid == 0) {
            half_logdet[0] = dead ? __longlong_as_double(0x7ff8000000000000ll) : 0.5 * ((double)n * log(r0) + ltot);
            if (s_i[2] != 0x7fffffff) status[0] = s_i[2] + 1;
            if (dead && guard) atomicOr(guard, 1 << (8 + guard_bit0 + sys));
        }
        cluster.sync();                                        // consumers may still be pulling from this CTA's shared memory
        return;
    }
    // ---- lattice ----
    double a[SCHUR_EPT], B[SCHUR_EPT];
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) { a[i] = (j0t + i == 0) ? 1.0 : 0.0; B[i] = a[i]; }
    const int mstart = max(0, SCHUR_BPW * gw0 - 1);            // warp 0 of this CTA becomes non-zero at step 256 gw0 - 1
    haveB = mstart;                                            // older boundary batches are zeros nobody reads
    for (int T = mstart; T < nb + nw - 1; ++T) {
        if (!fetch(min((T + 1) * SCHUR_EPT, n), c > 0 ? min(T + 1, nb) : 0)) { dead = true; break; }
        const int m = T - warp, j0 = m * SCHUR_EPT;
        const bool live = m >= 0 && m < nb && (m + 1) * SCHUR_EPT >= gw * SCHUR_WCHUNK;
        if (__all_sync(0xffffffffu, live)) {
            const int rs = j0 & (SCHUR_RING - 1);
            if (warp > 0) load8(rin, &ring[warp - 1][rs]);
            else if (c > 0) load8(rin, bin + j0);
            else {
#pragma unroll
                for (int i = 0; i < SCHUR_EPT; ++i) rin[i] = 0.0;                            // B[-1] = 0
            }
            load8(kp8, kap + j0);
            batch_full<0, false>(0.0, 0, false, lane == 0, lane == 31, kp8, rin, &ring[warp][rs], nullptr, nullptr, B, a);
            if (hasOut && top && lane == 31) {
#pragma unroll
                for (int i = 0; i < SCHUR_EPT; ++i) bout[j0 + i] = ring[warp][rs + i];
            }
        }
        __syncthreads();
        const int mt = T - (nw - 1);
        if (hasOut && mt >= 0 && mt < nb && (mt + 1) * SCHUR_EPT >= (gw0 + nw - 1) * SCHUR_WCHUNK && tid == blockDim.x - 1)
            publish2(-1, mt + 1);
    }
    // g = A_{n-1} / E_{n-1},  E_{n-1} = r0 * prod_k (1 - kappa_k^2): every CTA forms the product over all k in the same order
    double prod = 1.0;
    if (!dead) for (int k = 1 + tid; k < n; k += blockDim.x) { const double kp = kap[k]; prod *= (1.0 - kp) * (1.0 + kp); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) prod *= __shfl_xor_sync(0xffffffffu, prod, o);
    __syncthreads();
    if (lane == 0) red[warp] = prod;
    __syncthreads();
    double E = r0;
    for (int w = 0; w < nw; ++w) E *= red[w];
    const double invE = dead ? __longlong_as_double(0x7ff8000000000000ll) : 1.0 / E;
    if (dead && tid == 0 && guard) atomicOr(guard, 1 << (8 + guard_bit0 + sys));
#pragma unroll
    for (int i = 0; i < SCHUR_EPT; ++i) g[j0t + i] = a[i] * invE;
    cluster.sync();
}

// One CTA per axis.  spec[0..3][L] (bit-reversed order, scaled like launch_toeplitz_spectrum):
//   0: conj(G)/L   (v -> L(g)^T v)        2:  G / (L g0)   (v -> L(g) v / g0)
//   1: conj(H)/L   (v -> L(h)^T v)        3: -H / (L g0)   (v -> -L(h) v / g0)
// sKinv[d] = sum over |i-j| = d of K^-1[i][j]  (both triangles for d > 0).
__global__ void __launch_bounds__(FFT_THREADS, 1)
gs_prepare_kernel(const double* __restrict__ g, long long sG, int n, int L, int logL, const double2* __restrict__ W,
                  double2* __restrict__ spec, long long sSpec, double* __restrict__ sKinv, long long sS, const int* __restrict__ skip) {
    if (skip && *skip) return;
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
    static DeviceOnce once;
    if (!once.needed()) return GPHM_OK;
    GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(gs_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)fft_smem_bytes(FFT_MAX_L)));
    once.done();
    return GPHM_OK;
}

constexpr size_t kSplitSmem = (2 * (size_t)SCHUR_MAX_N + (SCHUR_MAX_THREADS / 32) * SCHUR_RING + 34) * sizeof(double) + 8 * sizeof(int);
constexpr size_t kDsmemSmem = (3 * (size_t)SCHUR_MAX_N + (SCHUR_MAX_THREADS / 32) * SCHUR_RING + 34) * sizeof(double) + 8 * sizeof(int);
constexpr int kSchurDefaultDsmem = 0;           // 1: hand-over through distributed shared memory (schur_levinson_dsmem_kernel); GPHM_SCHUR_DSMEM

int schur_split_factor(int n) {           // CTAs per role: GPHM_SCHUR_SPLIT (1, 2 or 4); needs n to be a multiple of 256 * split
    static const int want = [] { const char* e = getenv("GPHM_SCHUR_SPLIT"); return e ? atoi(e) : kSchurDefaultSplit; }();
    int ns = want >= 4 ? 4 : (want >= 2 ? 2 : 1);
    while (ns > 1 && (n % (SCHUR_WCHUNK * ns) != 0 || n / (SCHUR_EPT * ns) < 64)) ns >>= 1;
    return ns;
}

int launch_schur_levinson(const double* tabK, long long sTab, int n, double jitter, double* g, long long sG,
                          double* half_logdet, long long sLd, int* status, long long sStatus, double* gkap, long long sKap,
                          int* prog, long long sProg, int nsys, cudaStream_t st, long long* dbg, int* guard, int guard_bit0,
                          double* gbnd, long long sBnd, const int* skip) {
    if (n < 1 || n > SCHUR_MAX_N) { set_last_error("schur: n=%d outside [1,%d]", n, SCHUR_MAX_N); return GPHM_EINVAL; }
    if (nsys < 1 || nsys > 2) { set_last_error("schur: nsys=%d", nsys); return GPHM_EINVAL; }
    const int ns = gbnd && !dbg ? schur_split_factor(n) : 1;
    if (ns > 1) {
        static DeviceOnce once;
        if (once.needed()) {
            GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(schur_levinson_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplitSmem));
            GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(schur_levinson_dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDsmemSmem));
            once.done();
        }
        static const bool dsmem = [] { const char* e = getenv("GPHM_SCHUR_DSMEM"); return e ? atoi(e) != 0 : kSchurDefaultDsmem != 0; }();
        for (int s = 0; s < nsys; ++s) GPHM_CUDA_OK(cudaMemsetAsync(prog + s * sProg, 0, 8 * sizeof(int), st));
        LaunchScope scope(CAT_CHOL_DIAG, st, 8.0 * (double)n * n * nsys);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * ns * nsys); cfg.blockDim = dim3(n / (SCHUR_EPT * ns) + (dsmem ? 0 : 32));      // + the publisher warp
        cfg.dynamicSmemBytes = dsmem ? kDsmemSmem : kSplitSmem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2 * ns; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const double gmin = toeplitz_guard_min();
        if (dsmem)
            GPHM_CUDA_OK(cudaLaunchKernelEx(&cfg, schur_levinson_dsmem_kernel, tabK, sTab, n, jitter, g, sG, half_logdet, sLd, status, sStatus,
                                            gkap, sKap, prog, sProg, gbnd, sBnd, ns, guard, guard_bit0, gmin, skip));
        else {
            static const bool want_dbg = getenv("GPHM_SCHUR_DBG") != nullptr;      // diagnostics: cycles per CTA, total and waiting for a producer
            static long long* dbgbuf = nullptr;
            if (want_dbg && !dbgbuf) GPHM_CUDA_OK(cudaMalloc(&dbgbuf, sizeof(long long) * 64));
            GPHM_CUDA_OK(cudaLaunchKernelEx(&cfg, schur_levinson_split_kernel, tabK, sTab, n, jitter, g, sG, half_logdet, sLd, status, sStatus,
                                            gkap, sKap, prog, sProg, gbnd, sBnd, ns, guard, guard_bit0, gmin, skip, want_dbg ? dbgbuf : nullptr));
            if (want_dbg && nsys * 2 * ns <= 32) {
                long long h[64];
                GPHM_CUDA_OK(cudaStreamSynchronize(st));
                GPHM_CUDA_OK(cudaMemcpy(h, dbgbuf, sizeof(h), cudaMemcpyDeviceToHost));
                for (int i = 0; i < 2 * ns * nsys; ++i)
                    fprintf(stderr, "schur split: cta %d (%s %d) %lld cycles, %lld waiting\n", i, (i % (2 * ns)) / ns ? "lattice" : "generator",
                            i % ns, h[2 * i], h[2 * i + 1]);
            }
        }
        GPHM_LAUNCH_OK();
        return GPHM_OK;
    }
    const int threads = std::min(SCHUR_MAX_THREADS, ((n + SCHUR_EPT - 1) / SCHUR_EPT + 31) / 32 * 32);
    for (int s = 0; s < nsys; ++s) GPHM_CUDA_OK(cudaMemsetAsync(prog + s * sProg, 0, sizeof(int), st));
    {
        // The lattice CTA consumes what the generator CTA of the same system produces: launched as a thread-block CLUSTER
        // of 2, which the hardware co-schedules - a plain <<<>>> launch does not guarantee that both are resident together
        // (16-stream CUDA-graph ensembles, MPS, a saturated device).
        LaunchScope scope(CAT_CHOL_DIAG, st, 8.0 * (double)n * n * nsys);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * nsys); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const double gmin = toeplitz_guard_min();
        GPHM_CUDA_OK(cudaLaunchKernelEx(&cfg, schur_levinson_kernel, tabK, sTab, n, jitter, g, sG, half_logdet, sLd, status, sStatus,
                                        gkap, sKap, prog, sProg, guard, guard_bit0, gmin, dbg, skip));
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_gs_prepare(const double* g, long long sG, int n, int L, const double* W, double* spec, long long sSpec,
                      double* sKinv, long long sS, int nsys, cudaStream_t st, const int* skip) {
    GPHM_TRY(toeplitz_inv_init());
    if (L > FFT_MAX_L || L < 2 * n) { set_last_error("gs_prepare: L=%d does not fit n=%d", L, n); return GPHM_EINVAL; }
    {
        LaunchScope scope(CAT_FFT, st);
        gs_prepare_kernel<<<nsys, FFT_THREADS, fft_smem_bytes(L), st>>>(
            g, sG, n, L, ilog2i(L), reinterpret_cast<const double2*>(W), reinterpret_cast<double2*>(spec), sSpec / 2, sKinv, sS, skip);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
