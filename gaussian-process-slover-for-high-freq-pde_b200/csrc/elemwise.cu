// HBM-bound pieces of the log-joint: residual / boundary / quadratic-form reductions, the
// scalar assembly of the loss, dL/dU assembly, theta-gradient reductions and the fused Adam.
//
//   residual + eq_gap, boundary_gap          model_GP_solver_2d.py:123-143 (1-D: _1d.py:101-121,
//                                            advection: _advection.py:123-139)
//   loss assembly                            model_GP_solver_2d.py:157-174 (_1d.py:133-149)
//   reverse pass of the above                jax.value_and_grad at :179 (hand-derived)
//   Adam                                     optax.adam(lr) at :60,180-182 (b1=.9,b2=.999,eps=1e-8)
// All reductions are two-stage with a fixed tree, so results are run-to-run identical.
#include <algorithm>
#include <type_traits>
#include "common.cuh"
#include "kernfun.cuh"
#include "kernels.h"

namespace gphm {

// ---------------------------------------------------------------------------------------------
// residual: R <- e^{log_v} * (R + nl(U) - F);   part[2b] = sum r^2, part[2b+1] = sum A*Bt
// (on entry R holds c1*D1*A + Bt*D2^T from the GEMMs)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_kernel(double* __restrict__ R, const double* __restrict__ U, const double* __restrict__ F,
                const double* __restrict__ A, const double* __restrict__ Bt, size_t n, int allencahn,
                const double* __restrict__ base, const double* __restrict__ small, int Q, double* __restrict__ part) {
    __shared__ double red[33];
    const double ev = exp(small[6 * Q + 1]);
    double e = 0.0, q = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double r = R[i] - F[i];
        if (allencahn) { const double u = U[i] + (base ? base[i] : 0.0); r += u * (u * u - 1.0); }
        e += r * r;
        q += A[i] * Bt[i];
        R[i] = ev * r;
    }
    e = block_sum(e, red);
    q = block_sum(q, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = e; part[2 * blockIdx.x + 1] = q; }
}

int launch_residual(double* R, const double* U, const double* F, const double* A, const double* Bt, size_t n,
                    int eq_type, const double* base, const double* small, int Q, double* part, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); residual_kernel<<<kRedBlocks, 256, 0, st>>>(R, U, F, A, Bt, n, eq_type == 1, base, small, Q, part); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// ---------------------------------------------------------------------------------------------
// finalize (one CTA): boundary gap, log-dets, loss and the two scalar gradients.
// terms = [loss, logdet1, logdet2, quad, bgap, eqgap, dL/dlog_tau, dL/dlog_v]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t boundary_index(int e, int dim, int n1, int n2, const int* xind) {
    if (dim == 1) return (size_t)xind[e];
    if (e < n2) return (size_t)e;                                    // U[0, :]
    if (e < 2 * n2) return (size_t)(n1 - 1) * n2 + (e - n2);         // U[-1, :]
    if (e < 2 * n2 + n1) return (size_t)(e - 2 * n2) * n2;           // U[:, 0]
    return (size_t)(e - 2 * n2 - n1) * n2 + (n2 - 1);                // U[:, -1]
}

__global__ void __launch_bounds__(1024)
finalize_kernel(LossConsts c, const double* __restrict__ U, const double* __restrict__ bvals,
                const int* __restrict__ xind, const double* __restrict__ part,
                const double* __restrict__ ldp1, int nblk1, const double* __restrict__ ldp2, int nblk2,
                const double* __restrict__ small, double* __restrict__ eb, double* __restrict__ terms,
                double* __restrict__ gsmall, int* __restrict__ status) {
    __shared__ double red[33];
    const int tid = threadIdx.x;
    double e = 0.0, q = 0.0, b = 0.0, l1 = 0.0, l2 = 0.0;
    for (int i = tid; i < kRedBlocks; i += blockDim.x) { e += part[2 * i]; q += part[2 * i + 1]; }
    for (int i = tid; i < c.nb; i += blockDim.x) {
        const double d = U[boundary_index(i, c.dim, c.n1, c.n2, xind)] - bvals[i];
        eb[i] = d;
        b += d * d;
    }
    for (int i = tid; i < nblk1; i += blockDim.x) l1 += ldp1[i];
    for (int i = tid; i < nblk2; i += blockDim.x) l2 += ldp2[i];
    e = block_sum(e, red);
    q = block_sum(q, red);
    b = block_sum(b, red);
    l1 = 2.0 * block_sum(l1, red);
    l2 = 2.0 * block_sum(l2, red);
    if (tid == 0) {
        const double tau = small[6 * c.Q], v = small[6 * c.Q + 1];
        const double Nc = (double)c.n1 * (double)c.n2, Nb = (double)c.nb;
        const double prior = 0.5 * c.logdet * ((double)c.n2 * l1 + (double)c.n1 * l2) + 0.5 * q;
        const double loss = prior - c.llk_weight * (0.5 * Nb * tau - 0.5 * exp(tau) * b)
                                  - (0.5 * Nc * v - 0.5 * exp(v) * e);
        const double gtau = -c.llk_weight * (0.5 * Nb - 0.5 * exp(tau) * b);
        const double gv = -(0.5 * Nc - 0.5 * exp(v) * e);
        terms[0] = loss; terms[1] = l1; terms[2] = l2; terms[3] = q; terms[4] = b; terms[5] = e;
        terms[6] = gtau; terms[7] = gv;
        if (gsmall) { gsmall[6 * c.Q] = gtau; gsmall[6 * c.Q + 1] = gv; }
        if (status && !isfinite(loss)) status[2] = 1;
    }
}

int launch_finalize(const LossConsts& c, const double* U, const double* bvals, const int* xind,
                    const double* part, const double* ldp1, int nblk1, const double* ldp2, int nblk2,
                    const double* small, double* eb, double* terms, double* gsmall, int* status, cudaStream_t st) {
    {
        LaunchScope scope(CAT_ELEMWISE, st);
        finalize_kernel<<<1, 1024, 0, st>>>(c, U, bvals, xind, part, ldp1, nblk1, ldp2, nblk2, small, eb, terms, gsmall,
                                            status);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// Sharded step: loss terms and the two scalar gradients from the ALL-REDUCED partial sums
// sums = [eq_gap, quad, boundary_gap], ld = [log|K1|, log|K2|] (model_GP_solver_2d.py:158-174).
__global__ void mg_finalize_kernel(LossConsts c, const double* __restrict__ sums, const double* __restrict__ ld,
                                   const double* __restrict__ small, double* __restrict__ terms, double* __restrict__ gsmall,
                                   int* __restrict__ status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double e = sums[0], q = sums[1], b = sums[2], l1 = ld[0], l2 = ld[1];
    const double tau = small[6 * c.Q], v = small[6 * c.Q + 1];
    const double Nc = (double)c.n1 * (double)c.n2, Nb = (double)c.nb;
    const double prior = 0.5 * c.logdet * ((double)c.n2 * l1 + (double)c.n1 * l2) + 0.5 * q;
    const double loss = prior - c.llk_weight * (0.5 * Nb * tau - 0.5 * exp(tau) * b) - (0.5 * Nc * v - 0.5 * exp(v) * e);
    const double gtau = -c.llk_weight * (0.5 * Nb - 0.5 * exp(tau) * b);
    const double gv = -(0.5 * Nc - 0.5 * exp(v) * e);
    terms[0] = loss; terms[1] = l1; terms[2] = l2; terms[3] = q; terms[4] = b; terms[5] = e; terms[6] = gtau; terms[7] = gv;
    if (gsmall) { gsmall[6 * c.Q] = gtau; gsmall[6 * c.Q + 1] = gv; }
    if (status && !isfinite(loss)) status[2] = 1;
}

int launch_mg_finalize(const LossConsts& c, const double* sums, const double* ld, const double* small, double* terms,
                       double* gsmall, int* status, cudaStream_t st) {
    {
        LaunchScope scope(CAT_ELEMWISE, st);
        mg_finalize_kernel<<<1, 32, 0, st>>>(c, sums, ld, small, terms, gsmall, status);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// ---------------------------------------------------------------------------------------------
// dL/dU = W + S1 + S2 [+ G*(3U^2-1)] + lambda*e^{tau}*E_b ;  V1 = S1 + W/2, V2 = S2 + W/2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grad_u_kernel(size_t n, int allencahn, const double* __restrict__ base, const double* __restrict__ U, const double* __restrict__ G,
              const double* __restrict__ W, const double* __restrict__ S1, const double* __restrict__ S2,
              double* __restrict__ gU, double* __restrict__ V1, double* __restrict__ V2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double w = W[i], s1 = S1[i];
        const double s2 = S2 ? S2[i] : 0.0;
        double g = w + s1 + s2;
        if (allencahn) { const double u = U[i] + (base ? base[i] : 0.0); g += G[i] * (3.0 * u * u - 1.0); }
        gU[i] = g;
        if (V1) V1[i] = s1 + 0.5 * w;
        if (V2) V2[i] = s2 + 0.5 * w;
    }
}

// one CTA, phases separated by barriers so corner points accumulate deterministically
__global__ void __launch_bounds__(1024)
boundary_scatter_kernel(LossConsts c, const double* __restrict__ eb, const int* __restrict__ xind,
                        const double* __restrict__ small, double* __restrict__ gU) {
    const double s = c.llk_weight * exp(small[6 * c.Q]);
    if (c.dim == 1) {
        if (threadIdx.x == 0)
            for (int e = 0; e < c.nb; ++e) gU[xind[e]] += s * eb[e];
        return;
    }
    const int bounds[5] = {0, c.n2, 2 * c.n2, 2 * c.n2 + c.n1, 2 * c.n2 + 2 * c.n1};
    for (int ph = 0; ph < 4; ++ph) {
        for (int e = bounds[ph] + threadIdx.x; e < bounds[ph + 1]; e += blockDim.x)
            gU[boundary_index(e, 2, c.n1, c.n2, nullptr)] += s * eb[e];
        __syncthreads();
        __threadfence_block();
    }
}

int launch_grad_u(const LossConsts& c, const double* base, const double* U, const double* G, const double* W, const double* S1,
                  const double* S2, const double* eb, const int* xind, const double* small, double* gU,
                  double* V1, double* V2, cudaStream_t st) {
    const size_t n = (size_t)c.n1 * c.n2;
    { LaunchScope scope(CAT_ELEMWISE, st); grad_u_kernel<<<kRedBlocks, 256, 0, st>>>(n, c.eq_type == 1, base, U, G, W, S1, S2, gU, V1, V2); }
    GPHM_LAUNCH_OK();
    { LaunchScope scope(CAT_ELEMWISE, st); boundary_scatter_kernel<<<1, 1024, 0, st>>>(c, eb, xind, small, gU); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// ---------------------------------------------------------------------------------------------
// Toeplitz theta-gradient: diagonal sums of Kbar / Dbar, then an n x 3Q table contraction.
// ---------------------------------------------------------------------------------------------
constexpr int DS_ROWS = 16;
size_t diag_sums_part_doubles(int n) { return (size_t)2 * ((n + DS_ROWS - 1) / DS_ROWS) * (2 * (size_t)n); }

// part[(chunk*2 + which)*2n + (lag + n - 1)] = sum over the chunk's rows i of M[i, i+lag]
__global__ void __launch_bounds__(256)
diag_sums_kernel(const double* __restrict__ Kbar, const double* __restrict__ Dbar, int n, int ld,
                 double* __restrict__ part) {
    const int r0 = blockIdx.x * DS_ROWS, r1 = min(r0 + DS_ROWS, n);
    double* pk = part + ((size_t)blockIdx.x * 2 + 0) * (2 * (size_t)n);
    double* pd = part + ((size_t)blockIdx.x * 2 + 1) * (2 * (size_t)n);
    for (int l = threadIdx.x; l < 2 * n - 1; l += blockDim.x) {
        const int lag = l - (n - 1);
        double sk = 0.0, sd = 0.0;
        for (int i = r0; i < r1; ++i) {
            const int j = i + lag;
            if (j >= 0 && j < n) { sk += Kbar[(size_t)i * ld + j]; if (Dbar) sd += Dbar[(size_t)i * ld + j]; }
        }
        pk[l] = sk; pd[l] = sd;
    }
}

// sK[m] = sum_{|i-j|=m} Kbar ; sD[m] = same (ORDER 2) or dirsign*(lower - upper) (ORDER 1)
__global__ void __launch_bounds__(256)
diag_sums_reduce_kernel(const double* __restrict__ part, int n, int nchunks, int antisym, double dirsign,
                        double* __restrict__ sK, double* __restrict__ sD) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    double ku = 0.0, kl = 0.0, du = 0.0, dl = 0.0;
    for (int c = 0; c < nchunks; ++c) {
        const double* pk = part + ((size_t)c * 2 + 0) * (2 * (size_t)n);
        const double* pd = part + ((size_t)c * 2 + 1) * (2 * (size_t)n);
        ku += pk[n - 1 + m]; du += pd[n - 1 + m];          // j - i = m  (upper)
        kl += pk[n - 1 - m]; dl += pd[n - 1 - m];          // i - j = m  (lower)
    }
    if (m == 0) { sK[0] = ku; if (sD) sD[0] = antisym ? 0.0 : du; }
    else { sK[m] = ku + kl; if (sD) sD[m] = antisym ? dirsign * (dl - du) : (du + dl); }
}

int launch_diag_sums(const double* Kbar, const double* Dbar, int n, int ld, bool antisym, double dirsign,
                     double* part, double* sK, double* sD, cudaStream_t st) {
    const int nchunks = (n + DS_ROWS - 1) / DS_ROWS;
    { LaunchScope scope(CAT_ELEMWISE, st); diag_sums_kernel<<<nchunks, 256, 0, st>>>(Kbar, Dbar, n, ld, part); }
    GPHM_LAUNCH_OK();
    { LaunchScope scope(CAT_ELEMWISE, st); diag_sums_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, n, nchunks, antisym ? 1 : 0, dirsign, sK, sD); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// one CTA per mixture component q: g[t*Q+q] = sum_m sK[m]*d k/d theta_t + sD[m]*d k^(o)/d theta_t
struct ThetaGradArgs { const double* x[2]; int n[2]; const double* theta[2]; const double* sK[2]; const double* sD[2]; double* gtheta[2]; };
template <int KID, int ORDER>
__global__ void __launch_bounds__(256)
theta_grad_toeplitz_kernel(ThetaGradArgs A, int Q) {           // blockIdx.y = axis
    const int ax = blockIdx.y;
    const double* __restrict__ x = A.x[ax];
    const double* __restrict__ theta = A.theta[ax];
    const double* __restrict__ sK = A.sK[ax];
    const double* __restrict__ sD = A.sD[ax];
    double* __restrict__ gtheta = A.gtheta[ax];
    const int n = A.n[ax];
    __shared__ double red[33];
    const int q = blockIdx.x;
    const CompConst c = make_comp(KID, theta[q], theta[Q + q], theta[2 * Q + q]);
    const double x0 = x[0];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        double p0[3], pd[3];
        comp_partials<KID, ORDER>(fabs(x[m] - x0), c, p0, pd);
        const double k = sK[m], d = sD[m];
        a0 += k * p0[0] + d * pd[0];
        a1 += k * p0[1] + d * pd[1];
        a2 += k * p0[2] + d * pd[2];
    }
    a0 = block_sum(a0, red); a1 = block_sum(a1, red); a2 = block_sum(a2, red);
    if (threadIdx.x == 0) { gtheta[q] = a0; gtheta[Q + q] = a1; gtheta[2 * Q + q] = a2; }
}

int launch_theta_grad_toeplitz_multi(int kid, int order, const ThetaGradJob* jobs, int count, int Q, cudaStream_t st) {
    if (count < 1 || count > 2) { set_last_error("theta_grad: %d jobs", count); return GPHM_EINVAL; }
    ThetaGradArgs A = {};
    for (int i = 0; i < count; ++i) {
        A.x[i] = jobs[i].x; A.n[i] = jobs[i].n; A.theta[i] = jobs[i].theta; A.sK[i] = jobs[i].sK; A.sD[i] = jobs[i].sD;
        A.gtheta[i] = jobs[i].gtheta;
    }
    LaunchScope scope(CAT_ELEMWISE, st);
    int rc = GPHM_DISPATCH_KID_ORDER(kid, order,
        theta_grad_toeplitz_kernel<KID, ORDER><<<dim3(Q, count), 256, 0, st>>>(A, Q));
    if (rc != 0) { set_last_error("theta_grad: bad kernel id %d / order %d", kid, order); return GPHM_EINVAL; }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_theta_grad_toeplitz(int kid, int order, const double* x, int n, const double* theta, int Q,
                               const double* sK, const double* sD, double* gtheta, cudaStream_t st) {
    const ThetaGradJob job = {x, n, theta, sK, sD, gtheta};
    return launch_theta_grad_toeplitz_multi(kid, order, &job, 1, Q, st);
}

// ---------------------------------------------------------------------------------------------
// General (non-uniform grid) theta-gradient: sum_ij Kbar*dK/dtheta + Dbar*dD/dtheta directly.
// CTA = 256 threads x TG_E elements; part[block][3Q]; second kernel reduces over blocks.
// ---------------------------------------------------------------------------------------------
constexpr int TG_E = 4;
size_t theta_general_part_doubles(int n, int Q) {
    const size_t nblk = ((size_t)n * n + 256 * TG_E - 1) / (256 * TG_E);
    return nblk * 3 * (size_t)Q;
}

template <int KID, int ORDER>
__global__ void __launch_bounds__(256)
theta_grad_general_kernel(const double* __restrict__ x, int n, const double* __restrict__ theta, int Q,
                          const double* __restrict__ Kbar, const double* __restrict__ Dbar, int ld,
                          double* __restrict__ part) {
    __shared__ double red[33];
    double d[TG_E], kb[TG_E], db[TG_E];
    const size_t base = ((size_t)blockIdx.x * 256 + threadIdx.x) * TG_E;
#pragma unroll
    for (int e = 0; e < TG_E; ++e) {
        const size_t idx = base + e;
        if (idx < (size_t)n * n) {
            const int i = (int)(idx / n), j = (int)(idx - (size_t)i * n);
            const double df = x[i] - x[j];
            d[e] = fabs(df);
            kb[e] = Kbar[(size_t)i * ld + j];
            db[e] = Dbar[(size_t)i * ld + j];
            if (ORDER == 1 && df < 0.0) db[e] = -db[e];
        } else { d[e] = 0.0; kb[e] = 0.0; db[e] = 0.0; }
    }
    for (int q = 0; q < Q; ++q) {
        const CompConst c = make_comp(KID, theta[q], theta[Q + q], theta[2 * Q + q]);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int e = 0; e < TG_E; ++e) {
            double p0[3], pd[3];
            comp_partials<KID, ORDER>(d[e], c, p0, pd);
            a0 += kb[e] * p0[0] + db[e] * pd[0];
            a1 += kb[e] * p0[1] + db[e] * pd[1];
            a2 += kb[e] * p0[2] + db[e] * pd[2];
        }
        a0 = block_sum(a0, red); a1 = block_sum(a1, red); a2 = block_sum(a2, red);
        if (threadIdx.x == 0) {
            double* p = part + (size_t)blockIdx.x * 3 * Q;
            p[q] = a0; p[Q + q] = a1; p[2 * Q + q] = a2;
        }
    }
}

__global__ void __launch_bounds__(256)
column_reduce_kernel(const double* __restrict__ part, size_t nrows, int ncols, double* __restrict__ out) {
    __shared__ double red[33];
    const int c = blockIdx.x;
    double s = 0.0;
    for (size_t r = threadIdx.x; r < nrows; r += blockDim.x) s += part[r * ncols + c];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[c] = s;
}

int launch_theta_grad_general(int kid, int order, const double* x, int n, const double* theta, int Q,
                              const double* Kbar, const double* Dbar, int ld, double* part, double* gtheta,
                              cudaStream_t st) {
    const size_t nblk = ((size_t)n * n + 256 * TG_E - 1) / (256 * TG_E);
    int rc;
    { LaunchScope scope(CAT_ELEMWISE, st);
    rc = GPHM_DISPATCH_KID_ORDER(kid, order,
        theta_grad_general_kernel<KID, ORDER><<<(unsigned)nblk, 256, 0, st>>>(x, n, theta, Q, Kbar, Dbar, ld, part)); }
    if (rc != 0) { set_last_error("theta_grad: bad kernel id %d / order %d", kid, order); return GPHM_EINVAL; }
    GPHM_LAUNCH_OK();
    { LaunchScope scope(CAT_ELEMWISE, st); column_reduce_kernel<<<3 * Q, 256, 0, st>>>(part, nblk, 3 * Q, gtheta); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// ---------------------------------------------------------------------------------------------
// Adam (optax 0.1.4 defaults).  The step count lives on the device so the loop is graph-capturable.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m, double* __restrict__ v,
            size_t n, const long long* __restrict__ count, double lr) {
    constexpr double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double t = (double)(*count + 1);
    const double c1 = 1.0 - pow(b1, t), c2 = 1.0 - pow(b2, t);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double gi = g[i];
        const double mi = b1 * m[i] + (1.0 - b1) * gi;
        const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= lr * (mi / c1) / (sqrt(vi / c2) + eps);
    }
}
__global__ void count_inc_kernel(long long* count) { *count += 1; }

int launch_adam(double* p, const double* g, double* m, double* v, size_t n, const long long* count, double lr,
                cudaStream_t st) {
    if (n == 0) return GPHM_OK;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 8);
    { LaunchScope scope(CAT_ADAM, st); adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, count, lr); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}
// Adam on a short vector (the 6Q+2 kernel parameters) and ++count in ONE launch: every thread reads the count before
// the barrier, thread 0 bumps it after.
__global__ void __launch_bounds__(1024)
adam_inc_kernel(double* __restrict__ p, const double* __restrict__ g, double* __restrict__ m, double* __restrict__ v, int n,
                long long* count, double lr) {
    constexpr double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double t = (double)(*count + 1);
    __syncthreads();
    const double c1 = 1.0 - pow(b1, t), c2 = 1.0 - pow(b2, t);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double gi = g[i];
        const double mi = b1 * m[i] + (1.0 - b1) * gi;
        const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= lr * (mi / c1) / (sqrt(vi / c2) + eps);
    }
    if (threadIdx.x == 0) *count += 1;
}
int launch_adam_inc(double* p, const double* g, double* m, double* v, size_t n, long long* count, double lr, cudaStream_t st) {
    { LaunchScope scope(CAT_ADAM, st); adam_inc_kernel<<<1, 1024, 0, st>>>(p, g, m, v, (int)n, count, lr); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}
__global__ void __launch_bounds__(1024)
adam_out_kernel(const double* __restrict__ p, double* __restrict__ p_out, const double* __restrict__ g, double* __restrict__ m,
                double* __restrict__ v, int n, const long long* __restrict__ count, double lr) {
    constexpr double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double t = (double)(*count + 1);
    const double c1 = 1.0 - pow(b1, t), c2 = 1.0 - pow(b2, t);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double gi = g[i];
        const double mi = b1 * m[i] + (1.0 - b1) * gi;
        const double vi = b2 * v[i] + (1.0 - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p_out[i] = p[i] - lr * (mi / c1) / (sqrt(vi / c2) + eps);
    }
}
int launch_adam_out(const double* p, double* p_out, const double* g, double* m, double* v, size_t n, const long long* count,
                    double lr, cudaStream_t st) {
    { LaunchScope scope(CAT_ADAM, st); adam_out_kernel<<<1, 1024, 0, st>>>(p, p_out, g, m, v, (int)n, count, lr); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}
// flags[1] (skip) = the look-ahead result is valid (flags[0]) AND was computed for exactly this theta; flags[0] is consumed
__global__ void __launch_bounds__(256)
lk_compare_kernel(const double* __restrict__ small, const double* __restrict__ lk_small, int n, int* __restrict__ flags) {
    __shared__ int differ;
    if (threadIdx.x == 0) differ = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (__double_as_longlong(small[i]) != __double_as_longlong(lk_small[i])) differ = 1;
    __syncthreads();
    if (threadIdx.x == 0) { flags[1] = (flags[0] != 0 && !differ) ? 1 : 0; flags[0] = 0; }
}
__global__ void lk_set_kernel(int* flags, int value) { flags[0] = value; }
int launch_lk_compare(const double* small, const double* lk_small, int n, int* flags, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); lk_compare_kernel<<<1, 256, 0, st>>>(small, lk_small, n, flags); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}
int launch_lk_set(int* flags, int value, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); lk_set_kernel<<<1, 1, 0, st>>>(flags, value); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}
int launch_count_inc(long long* count, cudaStream_t st) {
    { LaunchScope scope(CAT_ADAM, st); count_inc_kernel<<<1, 1, 0, st>>>(count); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// ---------------------------------------------------------------------------------------------
// relative L2 error ||pred - truth|| / ||truth||   (model_GP_solver_2d.py:297-300)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rel_l2_part_kernel(const double* __restrict__ a, const double* __restrict__ b, size_t n, double* __restrict__ part) {
    __shared__ double red[33];
    double num = 0.0, den = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double t = b[i], d = a[i] - t;
        num += d * d; den += t * t;
    }
    num = block_sum(num, red); den = block_sum(den, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = num; part[2 * blockIdx.x + 1] = den; }
}
__global__ void __launch_bounds__(1024)
rel_l2_final_kernel(const double* __restrict__ part, double* __restrict__ out) {
    __shared__ double red[33];
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < kRedBlocks; i += blockDim.x) { num += part[2 * i]; den += part[2 * i + 1]; }
    num = block_sum(num, red); den = block_sum(den, red);
    if (threadIdx.x == 0) *out = sqrt(num) / sqrt(den);
}
int launch_rel_l2(const double* pred, const double* truth, size_t n, double* part, double* out, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); rel_l2_part_kernel<<<kRedBlocks, 256, 0, st>>>(pred, truth, n, part); }
    GPHM_LAUNCH_OK();
    { LaunchScope scope(CAT_ELEMWISE, st); rel_l2_final_kernel<<<1, 1024, 0, st>>>(part, out); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

__global__ void __launch_bounds__(256)
copy_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}
int launch_copy(double* dst, const double* src, size_t n, cudaStream_t st) {
    if (n == 0) return GPHM_OK;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 8);
    { LaunchScope scope(CAT_ELEMWISE, st); copy_kernel<<<blocks, 256, 0, st>>>(dst, src, n); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// ---- building blocks of the sharded (multi-GPU) step: the same math on a rank-local block ----
__global__ void __launch_bounds__(1024)
pair_reduce_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ out2) {
    __shared__ double red[33];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { a += part[2 * i]; b += part[2 * i + 1]; }
    a = block_sum(a, red); b = block_sum(b, red);
    if (threadIdx.x == 0) { out2[0] = a; out2[1] = b; }
}
int launch_pair_reduce(const double* part, double* out2, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); pair_reduce_kernel<<<1, 1024, 0, st>>>(part, kRedBlocks, out2); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// eb[e] = U[bidx[e]] - bvals[e];  out[0] = sum eb^2      (rank-local boundary points)
__global__ void __launch_bounds__(1024)
boundary_indexed_kernel(const double* __restrict__ U, const int* __restrict__ bidx, const double* __restrict__ bvals,
                        int nb, double* __restrict__ eb, double* __restrict__ out) {
    __shared__ double red[33];
    double b = 0.0;
    for (int e = threadIdx.x; e < nb; e += blockDim.x) {
        const double d = U[bidx[e]] - bvals[e];
        eb[e] = d;
        b += d * d;
    }
    b = block_sum(b, red);
    if (threadIdx.x == 0) out[0] = b;
}
int launch_boundary_indexed(const double* U, const int* bidx, const double* bvals, int nb, double* eb, double* out,
                            cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); boundary_indexed_kernel<<<1, 1024, 0, st>>>(U, bidx, bvals, nb, eb, out); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// gU[bidx[e]] += llk_weight * exp(log_tau) * eb[e]; indices are unique inside each of the two
// segments [0,nseg0) and [nseg0,nb) (row edges, column edges), which run one after the other.
__global__ void __launch_bounds__(1024)
boundary_scatter_indexed_kernel(double* __restrict__ gU, const int* __restrict__ bidx, const double* __restrict__ eb,
                                int nseg0, int nb, double llk_weight, const double* __restrict__ log_tau) {
    const double s = llk_weight * exp(*log_tau);
    for (int e = threadIdx.x; e < nseg0; e += blockDim.x) gU[bidx[e]] += s * eb[e];
    __syncthreads();
    __threadfence_block();
    for (int e = nseg0 + threadIdx.x; e < nb; e += blockDim.x) gU[bidx[e]] += s * eb[e];
}
int launch_boundary_scatter_indexed(double* gU, const int* bidx, const double* eb, int nseg0, int nb, double llk_weight,
                                    const double* log_tau, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st);
      boundary_scatter_indexed_kernel<<<1, 1024, 0, st>>>(gU, bidx, eb, nseg0, nb, llk_weight, log_tau); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

__global__ void __launch_bounds__(256)
lincomb_kernel(double* __restrict__ out, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
               size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}
int launch_lincomb(double* out, double a, const double* x, double b, const double* y, size_t n, cudaStream_t st) {
    if (n == 0) return GPHM_OK;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 8);
    { LaunchScope scope(CAT_ELEMWISE, st); lincomb_kernel<<<blocks, 256, 0, st>>>(out, a, x, b, y, n); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_grad_u_local(size_t n, int allencahn, const double* U, const double* G, const double* W, const double* S1,
                        const double* S2, double* gU, double* V1, double* V2, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); grad_u_kernel<<<kRedBlocks, 256, 0, st>>>(n, allencahn, nullptr, U, G, W, S1, S2, gU, V1, V2); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

__global__ void __launch_bounds__(256)
sum_scaled_kernel(const double* __restrict__ v, int n, double scale, double* __restrict__ out) {
    __shared__ double red[33];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = scale * s;
}
int launch_sum_scaled(const double* v, int n, double scale, double* out, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); sum_scaled_kernel<<<1, 256, 0, st>>>(v, n, scale, out); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

// gphm_toeplitz_solve: status 0 (SPD) becomes -1 when the conditioning guard fired
__global__ void status_merge_guard_kernel(int* __restrict__ status, const int* __restrict__ guard) {
    if (*status == 0 && *guard != 0) *status = -1;
}
int launch_status_merge_guard(int* status, const int* guard, cudaStream_t st) {
    { LaunchScope scope(CAT_ELEMWISE, st); status_merge_guard_kernel<<<1, 1, 0, st>>>(status, guard); }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
