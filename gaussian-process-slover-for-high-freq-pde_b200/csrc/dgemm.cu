// FP64 GEMM engine for the Kronecker contractions, triangular applications of L^-1 and the
// blocked factorisations:  C = alpha * op(A) * op(B) + beta * C, row-major, batched.
//
// These are the dense contractions of the reference's jitted step - jnp.linalg.solve and
// jnp.matmul at model_GP_solver_2d.py:104-105,112,119 and their reverse-mode counterparts
// (:179) - and carry ~99.9% of the 28 N^3 FLOPs per iteration at N=4096.
//
// Design (sm_100a): tcgen05 has no f64 kind, so native FP64 runs on the DMMA pipe
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  One CTA computes a BM x BN tile with a
// 3-stage cp.async (LDGSTS, zero-fill predicated) shared-memory pipeline over BK=16 slices;
// 16 warps each own a 32x32 sub-tile (32 FP64 accumulators per thread).  ptxas spaces dependent
// DMMAs of one warp with NOPs, so 2 warps per scheduler left the pipe 78 % busy (ncu, round 1);
// 4 warps per scheduler at <=128 registers keep it fed.  Shared-memory rows are
// padded by 4 doubles so every fragment LDS.64 is bank-conflict free.  Roofline: FP64 compute
// (DMMA issue rate); operand traffic per tile is 32 KB per 0.5 MFLOP, far below L2 bandwidth.
// Triangular operands (L^-1 applications, LAUUM-style products) restrict the k-range per tile
// (kmode flags) so the zero halves are never loaded or multiplied.
#include <algorithm>
#include <cstdint>
#include "common.cuh"
#include "kernels.h"

namespace gphm {

constexpr int BK = 16;
constexpr int STAGES = 3;
constexpr int PAD = 4;

template <int VEC>
__device__ __forceinline__ void cp_async(double* smem, const double* gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    if (VEC == 2)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Per-thread descriptors for copying one operand tile (MN of the m- or n-index by BK of k) into
// shared memory with cp.async.  Everything that does not depend on the k-tile (row/column
// validity, shared-memory offset, global pointer) is computed once; per k-tile a chunk costs a
// clamp, one LDGSTS and a pointer bump, and the chunks are interleaved with the DMMAs.
// KCONTIG: global element (mn, k) at G[mn*ld + k]  -> S[mn][k], row pitch BK+PAD
// else   : global element (k, mn) at G[k*ld + mn]  -> S[k][mn],  row pitch MN+PAD
// Everything outside [0,mn_max) x [kbegin,kend) is zero-filled by cp.async's src-size operand.
template <int MN, bool KCONTIG, int VEC, int NT>
struct TileLoader {
    static constexpr int CPR = KCONTIG ? BK / VEC : MN / VEC;        // chunks per shared-memory row
    static constexpr int TOTAL = MN * BK / VEC;
    static constexpr int NCH = TOTAL / NT;
    static_assert(TOTAL % NT == 0, "tile chunks must divide evenly over the threads");
    const double* g[NCH];     // next global address of each chunk
    int soff[NCH];            // offset inside a stage buffer (doubles)
    int aux[NCH];             // KCONTIG: k offset of the chunk (huge if the row is out of range)
                              // else   : k row of the chunk | valid element count << 8
    const double* base;
    long long step;

    __device__ __forceinline__ void init(const double* G, int ld, int mn0, int mn_max, int kbegin) {
        base = G;
        step = KCONTIG ? (long long)BK : (long long)BK * ld;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = threadIdx.x + i * NT;
            const int r = c / CPR, cc = (c % CPR) * VEC;
            if (KCONTIG) {
                const int gr = mn0 + r;
                soff[i] = r * (BK + PAD) + cc;
                aux[i] = gr < mn_max ? cc : (1 << 28);
                g[i] = G + (size_t)min(gr, mn_max - 1) * ld + kbegin + cc;
            } else {
                const int gc = mn0 + cc;
                const int nv = min(max(mn_max - gc, 0), VEC);
                soff[i] = r * (MN + PAD) + cc;
                aux[i] = r | (nv << 8);
                g[i] = G + (size_t)(kbegin + r) * ld + min(gc, mn_max - 1);
            }
        }
    }
    // issue chunk i of the k-tile starting at k0 into the stage buffer S, then advance
    __device__ __forceinline__ void issue(int i, double* S, int k0, int kend) {
        int nv;
        if (KCONTIG) nv = min(max(kend - k0 - aux[i], 0), VEC);
        else nv = (k0 + (aux[i] & 255) < kend) ? (aux[i] >> 8) : 0;
        cp_async<VEC>(S + soff[i], nv > 0 ? g[i] : base, nv * 8);
        g[i] += step;
    }
};

template <int BM, int BN, int WARPS_M, int WARPS_N, bool TA, bool TB, int VEC, int MINB>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MINB)
dgemm_kernel(const GemmArgs p) {
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
    constexpr int MI = WTM / 8, NI = WTN / 8;
    constexpr int A_TILE = TA ? BK * (BM + PAD) : BM * (BK + PAD);
    constexpr int B_TILE = TB ? BN * (BK + PAD) : BK * (BN + PAD);
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_TILE;

    // ---- tile coordinates: groups of 8 tile-rows share B columns in L2; heavy tiles first ----
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
    int pid = blockIdx.x;
    if (p.kmode & (KM_A_LOWER | KM_B_UPPER)) pid = tiles_m * tiles_n - 1 - pid;
    constexpr int GROUP = 8;
    const int per_group = GROUP * tiles_n;
    const int group_id = pid / per_group;
    const int first_m = group_id * GROUP;
    const int gsz = min(tiles_m - first_m, GROUP);
    const int tm = first_m + (pid % per_group) % gsz;
    const int tn = (pid % per_group) / gsz;
    const int m0 = tm * BM, n0 = tn * BN;
    if ((p.kmode & KM_C_LOWER) && n0 >= m0 + BM) return;

    const double* __restrict__ A = p.A + (size_t)blockIdx.y * p.sA;
    const double* __restrict__ B = p.B + (size_t)blockIdx.y * p.sB;
    double* __restrict__ C = p.C + (size_t)blockIdx.y * p.sC;

    int kbegin = 0, kend = p.K;
    if (p.kmode & KM_A_LOWER) kend = min(kend, m0 + BM);
    if (p.kmode & KM_A_UPPER) kbegin = max(kbegin, m0);
    if (p.kmode & KM_B_LOWER) kbegin = max(kbegin, n0);
    if (p.kmode & KM_B_UPPER) kend = min(kend, n0 + BN);
    const int KT = kend > kbegin ? (kend - kbegin + BK - 1) / BK : 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, t = lane & 3;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    using LoaderA = TileLoader<BM, !TA, VEC, NT>;
    using LoaderB = TileLoader<BN, TB, VEC, NT>;
    constexpr int NCH = LoaderA::NCH + LoaderB::NCH;
    constexpr int KSTEPS = BK / 4;
    static_assert(NCH % KSTEPS == 0, "chunks are spread evenly over the k4-steps");
    LoaderA la;
    LoaderB lb;
    if (KT > 0) {
        la.init(A, p.lda, m0, p.M, kbegin);
        lb.init(B, p.ldb, n0, p.N, kbegin);
    }
    auto issue_chunk = [&](int c, int slot, int k0) {      // c is a compile-time constant after unrolling
        if (c < LoaderA::NCH) la.issue(c, As + slot * A_TILE, k0, kend);
        else lb.issue(c - LoaderA::NCH, Bs + slot * B_TILE, k0, kend);
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) issue_chunk(c, s, kbegin + s * BK);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int ktn = kt + STAGES - 1;
        const bool more = ktn < KT;
        const int slot_n = ktn % STAGES, k0_n = kbegin + ktn * BK;

        const double* as = As + (kt % STAGES) * A_TILE;
        const double* bs = Bs + (kt % STAGES) * B_TILE;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int r = wm * WTM + i * 8 + g;
                af[i] = TA ? as[(kk * 4 + t) * (BM + PAD) + r] : as[r * (BK + PAD) + kk * 4 + t];
            }
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                const int c = wn * WTN + j * 8 + g;
                bf[j] = TB ? bs[c * (BK + PAD) + kk * 4 + t] : bs[(kk * 4 + t) * (BN + PAD) + c];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            // the stage being refilled was consumed in iteration kt-1 (barrier above)
            if (more) {
#pragma unroll
                for (int c = kk * (NCH / KSTEPS); c < (kk + 1) * (NCH / KSTEPS); ++c) issue_chunk(c, slot_n, k0_n);
            }
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    // ---- epilogue: each thread owns 2 adjacent columns per 8x8 fragment -> 16 B stores ----
    const bool cvec = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = m0 + wm * WTM + i * 8 + g;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int col = n0 + wn * WTN + j * 8 + 2 * t;
            if (col >= p.N) continue;
            double* ptr = C + (size_t)row * p.ldc + col;
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            const bool two = (col + 1 < p.N);
            if (two && cvec) {
                if (p.beta != 0.0) {
                    const double2 o = *reinterpret_cast<const double2*>(ptr);
                    v0 += p.beta * o.x; v1 += p.beta * o.y;
                }
                *reinterpret_cast<double2*>(ptr) = make_double2(v0, v1);
            } else {
                if (p.beta != 0.0) { v0 += p.beta * ptr[0]; if (two) v1 += p.beta * ptr[1]; }
                ptr[0] = v0;
                if (two) ptr[1] = v1;
            }
        }
    }
}

template <int BM, int BN, bool TA, bool TB>
constexpr size_t gemm_smem_bytes() {
    return (size_t)STAGES * ((TA ? BK * (BM + PAD) : BM * (BK + PAD)) + (TB ? BN * (BK + PAD) : BK * (BN + PAD))) *
           sizeof(double);
}

// FLOPs actually issued for this launch (2*m*n*k over the computed tiles with their k-ranges).
static double gemm_flops(const GemmArgs& g, int BM, int BN) {
    if (!profiling_enabled()) return 0.0;
    double f = 0.0;
    for (int m0 = 0; m0 < g.M; m0 += BM)
        for (int n0 = 0; n0 < g.N; n0 += BN) {
            if ((g.kmode & KM_C_LOWER) && n0 >= m0 + BM) continue;
            int kb = 0, ke = g.K;
            if (g.kmode & KM_A_LOWER) ke = std::min(ke, m0 + BM);
            if (g.kmode & KM_A_UPPER) kb = std::max(kb, m0);
            if (g.kmode & KM_B_LOWER) kb = std::max(kb, n0);
            if (g.kmode & KM_B_UPPER) ke = std::min(ke, n0 + BN);
            if (ke > kb) f += 2.0 * std::min(BM, g.M - m0) * std::min(BN, g.N - n0) * (double)(ke - kb);
        }
    return f * g.batch;
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB, int VEC, int MINB>
static int launch_cfg(const GemmArgs& g, cudaStream_t st) {
    const int tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    dim3 grid(tiles, g.batch), block(WM * WN * 32);
    LaunchScope scope(CAT_DGEMM, st, gemm_flops(g, BM, BN));
    dgemm_kernel<BM, BN, WM, WN, TA, TB, VEC, MINB><<<grid, block, gemm_smem_bytes<BM, BN, TA, TB>(), st>>>(g);
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB, int VEC, int MINB>
static int set_attr() {
    GPHM_CUDA_OK(cudaFuncSetAttribute(dgemm_kernel<BM, BN, WM, WN, TA, TB, VEC, MINB>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)gemm_smem_bytes<BM, BN, TA, TB>()));
    return GPHM_OK;
}

#define GPHM_FOR_ALL_GEMM(F)                                                                   \
    F(false, false, 2) F(false, true, 2) F(true, false, 2) F(true, true, 2)                    \
    F(false, false, 1) F(false, true, 1) F(true, false, 1) F(true, true, 1)

int dgemm_init() {
    static DeviceOnce once;
    if (!once.needed()) return GPHM_OK;
#define GPHM_SET(TA, TB, VEC)                                                                              \
    if (set_attr<128, 128, 4, 4, TA, TB, VEC, 1>() != GPHM_OK) { once.done(false); return GPHM_ECUDA; }    \
    if (set_attr<64, 64, 2, 2, TA, TB, VEC, 3>() != GPHM_OK) { once.done(false); return GPHM_ECUDA; }
    GPHM_FOR_ALL_GEMM(GPHM_SET)
#undef GPHM_SET
    once.done();
    return GPHM_OK;
}

int launch_dgemm(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return GPHM_OK;
    if (g.K < 0 || g.lda < 1 || g.ldb < 1 || g.ldc < 1) { set_last_error("dgemm: bad shape"); return GPHM_EINVAL; }
    GPHM_TRY(dgemm_init());
    auto aligned = [](const double* p, int ld, long long s) {
        return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 1) == 0 && (s & 1) == 0;
    };
    const int vec = (aligned(g.A, g.lda, g.sA) && aligned(g.B, g.ldb, g.sB)) ? 2 : 1;
    const long long big_tiles = (long long)((g.M + 127) / 128) * ((g.N + 127) / 128) * g.batch;
    const bool big = big_tiles >= 96;   // otherwise 64x64 tiles to fill the 148 SMs
#define GPHM_GO(TA, TB, VEC)                                                                       \
    if ((g.transA != 0) == TA && (g.transB != 0) == TB && vec == VEC)                              \
        return big ? launch_cfg<128, 128, 4, 4, TA, TB, VEC, 1>(g, st)                             \
                   : launch_cfg<64, 64, 2, 2, TA, TB, VEC, 3>(g, st);
    GPHM_FOR_ALL_GEMM(GPHM_GO)
#undef GPHM_GO
    set_last_error("dgemm: no kernel variant");
    return GPHM_EINVAL;
}

}  // namespace gphm
