// FP64-accurate GEMM on the 5th-generation tensor cores: Ozaki splitting into int8 slices, tcgen05.mma kind::i8 with
// TMA-staged operands and int32 accumulators in TMEM (sm_100a only).
//
// Replaces the plain dense contractions of the general (non-uniform grid) path - jnp.matmul(K_dxx1, K1inv_U) and its
// reverse-mode products, model_GP_solver_2d.py:112,119 under :179 - when the plan sets force_general bit 6.  The solves
// (jnp.linalg.solve, :104-105) stay on native FP64: their error is amplified by cond(K).
//
//   C = alpha * op(A) op(B) + beta * C          op(A): M x K,  op(B): K x N,  row-major FP64
//
// 1. ozaki_split_kernel.  Every row i of op(A) is scaled by 2^-eA[i] (eA[i] = exponent of its largest magnitude), every
//    column j of op(B) by 2^-eB[j], and cut into S signed digits:  a' = sum_s dA_s 2^-w_s + rho,  w_s = 6 + 7 (s-1),
//    |dA_s| <= 64 (int8), |rho| <= 2^-(w_S + 1).  All of it is exact in FP64 (power-of-two scalings, rint, subtraction).
//    The slices are written K-major ([slice][row][k], zero padded to the tile sizes) - the layout the tensor core reads.
// 2. ozaki_gemm_kernel.  One CTA per 128 x 64 tile of C, 6 warps:
//      warp 0   TMA producer: per 64-wide k-block, S A-slices (128 x 64 B) and S B-slices (64 x 64 B) into a 2-stage
//               ring of 64-byte-swizzled shared-memory tiles (cp.async.bulk.tensor.2d, mbarrier complete_tx)
//      warp 1   one thread issues tcgen05.mma.cta_group::1.kind::i8 (M128 N64 K32) for every slice pair with s + t <= S + 1;
//               pairs of equal s + t have equal weight and share ONE int32 accumulator in TMEM (S accumulators x 64
//               columns = all 512 columns for S = 8), so a single pass over K feeds all diagonals;
//               tcgen05.commit releases the stage / signals the epilogue
//      warps 2-5  epilogue: tcgen05.ld the S accumulators, recombine in FP64  sum_d 2^-(12 + 7 d) I_d, rescale by
//               2^(eA[i] + eB[j]), apply alpha / beta, store.
//    int32 sums are exact: K * 64 * 64 * S <= 2^31 for K <= 65536.
//
// Error bound (the "stated looser bound" of north_star; derivation in DESIGN section 4): with a_i = max_k |op(A)[i,k]|,
// b_j = max_k |op(B)[k,j]|,
//      | C_ozaki - C_exact |_ij  <=  |alpha| * 4 K (S + 1.1) 2^(-7 S) * a_i b_j        ( + FP64 rounding of the recombination )
// i.e. 5.2e-13 a_i b_j at K = 4096, S = 8 - the size of DGEMM's own rounding bound K eps sum |a||b|.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include "common.cuh"
#include "kernels.h"

namespace gphm {

namespace {

constexpr int OZ_BM = 128, OZ_BN = 64, OZ_BK = 64, OZ_SMAX = 8, OZ_STAGES = 2;
constexpr int OZ_THREADS = 192;
constexpr int OZ_A_SLICE_BYTES = OZ_BM * OZ_BK;      // 8 KB
constexpr int OZ_B_SLICE_BYTES = OZ_BN * OZ_BK;      // 4 KB

__host__ __device__ inline int oz_round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------------------
// splitting
// ---------------------------------------------------------------------------------------------------------------------
// expo[r] = exponent e with max_k |Y[r,k]| < 2^e (0 for an all-zero row).  Y[r,k] = X[r*ld + k] (!TRANS) or X[k*ld + r].
template <bool TRANS>
__global__ void __launch_bounds__(256)
ozaki_rowmax_kernel(const double* __restrict__ X, int ld, int R, int K, int* __restrict__ expo) {
    __shared__ double red[8][33];
    if (!TRANS) {                                   // one warp per row
        const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
        double m = 0.0;
        if (r < R)
            for (int k = lane; k < K; k += 32) m = fmax(m, fabs(X[(size_t)r * ld + k]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && r < R) { int e = 0; if (m > 0.0) frexp(m, &e); expo[r] = e; }
    } else {                                        // 32 rows (contiguous in memory) per block, 8 k-lanes
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        const int r = blockIdx.x * 32 + tx;
        double m = 0.0;
        if (r < R)
            for (int k = ty; k < K; k += 8) m = fmax(m, fabs(X[(size_t)k * ld + r]));
        red[ty][tx] = m;
        __syncthreads();
        if (ty == 0 && r < R) {
#pragma unroll
            for (int j = 1; j < 8; ++j) m = fmax(m, red[j][tx]);
            int e = 0; if (m > 0.0) frexp(m, &e); expo[r] = e;
        }
    }
}

// slices[s][r][k] (int8, rows padded to Rpad, k to Kpad; the buffer is zeroed beforehand) for a 32 x 32 tile per block
template <bool TRANS>
__global__ void __launch_bounds__(256)
ozaki_split_kernel(const double* __restrict__ X, int ld, int R, int K, int S, const int* __restrict__ expo,
                   int8_t* __restrict__ slices, int Rpad, int Kpad) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    // load: coalesced along the contiguous direction of X
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int a = ty + 8 * j;
        if (!TRANS) { const int r = r0 + a, k = k0 + tx; tile[a][tx] = (r < R && k < K) ? X[(size_t)r * ld + k] : 0.0; }
        else        { const int k = k0 + a, r = r0 + tx; tile[tx][a] = (r < R && k < K) ? X[(size_t)k * ld + r] : 0.0; }
    }
    __syncthreads();
    // digits: thread (row a, four consecutive k) -> one 4-byte store per slice
    const int a = threadIdx.x >> 3, kq = (threadIdx.x & 7) * 4;
    const int r = r0 + a;
    if (r >= R || k0 + kq >= Kpad) return;
    const int e = expo[r];
    double v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = ldexp(tile[a][kq + i], 6 - e);          // |v| < 64
    for (int s = 0; s < S; ++s) {
        char4 d;
        int8_t* dp = reinterpret_cast<int8_t*>(&d);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double q = rint(v[i]);
            dp[i] = (int8_t)(int)q;
            v[i] = (v[i] - q) * 128.0;                                            // remainder in [-64, 64]
        }
        *reinterpret_cast<char4*>(slices + ((size_t)s * Rpad + r) * Kpad + k0 + kq) = d;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "OZ_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra OZ_DONE;\n\t"
        "bra OZ_WAIT;\n\t"
        "OZ_DONE:\n\t"
        "}\n" :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, M128 N64 K32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major tile with 64-byte rows, 64-byte swizzle (cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO (unused for swizzled K-major, 1) << 16 | SBO (8 rows x 64 B = 512 B >> 4) << 32 | version 1 << 46 |
// layout SWIZZLE_64B (4) << 61.  Advancing by one K32 slab = +32 bytes on the start address.
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kOzIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);

struct OzSmem {
    uint64_t full[OZ_STAGES], empty[OZ_STAGES], acc_full;
    uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int S, int Mpad, int Npad,
                  int nkb, const int* __restrict__ eA, const int* __restrict__ eB, double* __restrict__ C, int ldc, int M, int N,
                  double alpha, double beta) {
    extern __shared__ uint8_t oz_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_raw) + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 512 B
    uint8_t* sA = base;                                                  // [stage][slice][128 rows][64 B]
    uint8_t* sB = base + (size_t)OZ_STAGES * S * OZ_A_SLICE_BYTES;       // [stage][slice][ 64 rows][64 B]
    OzSmem* sm = reinterpret_cast<OzSmem*>(sB + (size_t)OZ_STAGES * S * OZ_B_SLICE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * OZ_BM, n0 = blockIdx.x * OZ_BN;
    const uint32_t tmem_cols = S <= 4 ? 256u : 512u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < OZ_STAGES; ++i) { mbar_init(&sm->full[i], 1); mbar_init(&sm->empty[i], 1); }
        mbar_init(&sm->acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1) tmem_alloc(&sm->tmem_base, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm->tmem_base;

    if (warp == 0) {
        if (lane == 0) {                                                 // ---- TMA producer ----
            const uint32_t stage_bytes = (uint32_t)S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % OZ_STAGES;
                mbar_wait(&sm->empty[st], ((kb / OZ_STAGES) & 1) ^ 1);
                mbar_expect_tx(&sm->full[st], stage_bytes);
                for (int s = 0; s < S; ++s) {
                    tma_load_2d(sA + ((size_t)st * S + s) * OZ_A_SLICE_BYTES, &tmA, &sm->full[st], kb * OZ_BK, s * Mpad + m0);
                    tma_load_2d(sB + ((size_t)st * S + s) * OZ_B_SLICE_BYTES, &tmB, &sm->full[st], kb * OZ_BK, s * Npad + n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                                 // ---- MMA issuer ----
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % OZ_STAGES;
                mbar_wait(&sm->full[st], (kb / OZ_STAGES) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + (size_t)st * S * OZ_A_SLICE_BYTES);
                const uint32_t b0 = smem_u32(sB + (size_t)st * S * OZ_B_SLICE_BYTES);
#pragma unroll 1
                for (int ks = 0; ks < OZ_BK / 32; ++ks) {
                    for (int s = 0; s < S; ++s) {
                        const uint64_t da = smem_desc_sw64(a0 + s * OZ_A_SLICE_BYTES + ks * 32);
                        for (int t = 0; s + t < S; ++t) {                // slice pair (s, t), 0-based: diagonal d = s + t
                            const uint64_t db = smem_desc_sw64(b0 + t * OZ_B_SLICE_BYTES + ks * 32);
                            // the first product of a diagonal overwrites: k-block 0, slab 0, s == 0 (pair (0, d))
                            mma_i8(tmem + (uint32_t)(s + t) * OZ_BN, da, db, kOzIdesc, (kb | ks | s) != 0 ? 1u : 0u);
                        }
                    }
                }
                tc_commit(&sm->empty[st]);                               // stage free once these MMAs have read it
            }
            tc_commit(&sm->acc_full);
        }
    } else {                                                             // ---- epilogue: warps 2..5 -> TMEM lane quarters 2,3,0,1 ----
        mbar_wait(&sm->acc_full, 0);
        tc_fence_after();
        const int quarter = warp & 3;
        const int row = m0 + quarter * 32 + lane;
        const int ea = row < M ? eA[row] : 0;
        for (int c = 0; c < OZ_BN / 16; ++c) {
            double acc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0;
            for (int d = S - 1; d >= 0; --d) {                           // smallest weights first
                uint32_t r[16];
                tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * OZ_BN + c * 16), r);
                const double w = __longlong_as_double((long long)(1023 - (12 + 7 * d)) << 52);      // 2^-(12 + 7 d)
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fma((double)(int)r[j], w, acc[j]);
            }
            if (row < M) {
                const int col0 = n0 + c * 16;
                double* crow = C + (size_t)row * ldc;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = col0 + j;
                    if (col < N) {
                        const double v = alpha * ldexp(acc[j], ea + eB[col]);
                        crow[col] = beta != 0.0 ? fma(beta, crow[col], v) : v;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// int8 [rows][Kpad] row-major, box {64 bytes along k, box_rows}, 64-byte swizzle
int make_map(CUtensorMap* map, const int8_t* ptr, size_t rows, int Kpad, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_last_error("ozaki: cuTensorMapEncodeTiled is not available in this driver"); return GPHM_ECUDA; }
    const cuuint64_t gdim[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)Kpad};
    const cuuint32_t box[2] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("ozaki: cuTensorMapEncodeTiled failed (%d)", (int)r); return GPHM_ECUDA; }
    return GPHM_OK;
}

size_t oz_smem_bytes(int S) {
    return 1024 + (size_t)OZ_STAGES * S * (OZ_A_SLICE_BYTES + OZ_B_SLICE_BYTES) + sizeof(OzSmem) + 64;
}

}  // namespace

int ozaki_default_slices() {
    static const int s = [] { const char* e = getenv("GPHM_OZAKI_SLICES"); const int v = e ? atoi(e) : 8; return std::min(OZ_SMAX, std::max(2, v)); }();
    return s;
}

size_t ozaki_work_bytes(int M, int N, int K, int S) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    S = std::min(OZ_SMAX, std::max(2, S));
    const size_t Mp = oz_round_up(M, OZ_BM), Np = oz_round_up(N, OZ_BN), Kp = oz_round_up(K, OZ_BK);
    return (size_t)S * (Mp + Np) * Kp + sizeof(int) * (Mp + Np) + 1024;
}

double ozaki_error_factor(int K, int S) {                 // |dC_ij| <= |alpha| * factor * a_i * b_j
    return 4.0 * (double)K * (S + 1.1) * exp2(-7.0 * S);
}

int launch_ozaki_dgemm(bool transA, bool transB, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
                       int ldb, double beta, double* C, int ldc, int S, void* work, size_t work_bytes, cudaStream_t st) {
    if (M <= 0 || N <= 0) return GPHM_OK;
    if (K <= 0) { set_last_error("ozaki: K must be positive"); return GPHM_EINVAL; }
    if (K > 65536) { set_last_error("ozaki: K=%d exceeds the exact int32 accumulation range (65536)", K); return GPHM_EINVAL; }
    S = std::min(OZ_SMAX, std::max(2, S));
    if (work_bytes < ozaki_work_bytes(M, N, K, S)) { set_last_error("ozaki: workspace too small"); return GPHM_ENOMEM; }
    static DeviceOnce once;
    if (once.needed()) {
        GPHM_ONCE_CUDA_OK(once, cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz_smem_bytes(OZ_SMAX)));
        once.done();
    }
    const int Mp = oz_round_up(M, OZ_BM), Np = oz_round_up(N, OZ_BN), Kp = oz_round_up(K, OZ_BK);
    uint8_t* w = static_cast<uint8_t*>(work);
    w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(w) + 255) & ~(uintptr_t)255);
    int8_t* slA = reinterpret_cast<int8_t*>(w);
    int8_t* slB = slA + (size_t)S * Mp * Kp;
    int* eA = reinterpret_cast<int*>(slB + (size_t)S * Np * Kp);
    int* eB = eA + Mp;
    GPHM_CUDA_OK(cudaMemsetAsync(slA, 0, (size_t)S * (Mp + Np) * Kp + sizeof(int) * (Mp + Np), st));
    // op(A) rows: A row-major M x K (no trans) or stored K x M (trans);  op(B)^T rows: B stored K x N (no trans -> TRANS) or N x K
    const dim3 ga((K + 31) / 32, (M + 31) / 32), gb((K + 31) / 32, (N + 31) / 32);
    {
        LaunchScope scope(CAT_GRAM, st, 0.0, 8.0 * (double)M * K);
        if (!transA) ozaki_rowmax_kernel<false><<<(M + 7) / 8, 256, 0, st>>>(A, lda, M, K, eA);
        else ozaki_rowmax_kernel<true><<<(M + 31) / 32, 256, 0, st>>>(A, lda, M, K, eA);
    }
    {
        LaunchScope scope(CAT_GRAM, st, 0.0, 8.0 * (double)N * K);
        if (transB) ozaki_rowmax_kernel<false><<<(N + 7) / 8, 256, 0, st>>>(B, ldb, N, K, eB);
        else ozaki_rowmax_kernel<true><<<(N + 31) / 32, 256, 0, st>>>(B, ldb, N, K, eB);
    }
    {
        LaunchScope scope(CAT_GRAM, st, 0.0, (8.0 + S) * (double)M * K);
        if (!transA) ozaki_split_kernel<false><<<ga, 256, 0, st>>>(A, lda, M, K, S, eA, slA, Mp, Kp);
        else ozaki_split_kernel<true><<<ga, 256, 0, st>>>(A, lda, M, K, S, eA, slA, Mp, Kp);
    }
    {
        LaunchScope scope(CAT_GRAM, st, 0.0, (8.0 + S) * (double)N * K);
        if (transB) ozaki_split_kernel<false><<<gb, 256, 0, st>>>(B, ldb, N, K, S, eB, slB, Np, Kp);
        else ozaki_split_kernel<true><<<gb, 256, 0, st>>>(B, ldb, N, K, S, eB, slB, Np, Kp);
    }
    GPHM_LAUNCH_OK();
    CUtensorMap tmA, tmB;
    GPHM_TRY(make_map(&tmA, slA, (size_t)S * Mp, Kp, OZ_BM));
    GPHM_TRY(make_map(&tmB, slB, (size_t)S * Np, Kp, OZ_BN));
    {
        LaunchScope scope(CAT_DGEMM, st, 2.0 * M * (double)N * K, 0.0);     // FP64-equivalent FLOPs (int8 MACs: x S (S + 1) / 2)
        ozaki_gemm_kernel<<<dim3(Np / OZ_BN, Mp / OZ_BM), OZ_THREADS, oz_smem_bytes(S), st>>>(tmA, tmB, S, Mp, Np, Kp / OZ_BK, eA, eB, C,
                                                                                                  ldc, M, N, alpha, beta);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
