// Layout exchange of the sharded 2-D step as ONE kernel over NVLink peer memory (no NCCL on the data path).
//
// The sharded step alternates between row blocks (h x N2) and transposed column blocks (w x N1) of N1 x N2 fields
// (dist.py; SURVEY 8e: left-multiplications by axis-1 operators act on columns, right-multiplications by axis-2 operators
// on rows).  Round 1 did every exchange as  transpose-pack kernel -> ncclAllToAll -> segment-unpack kernel.  Here the
// transposing kernel stores each 32 x 32 tile STRAIGHT into the destination rank's result buffer, in its final layout,
// through a peer pointer (cudaIpc-mapped device memory, NVLink 5 / NVSwitch):
//
//     out_d[a][c][me * rows + r] = X_a[r][d * pc + c]          d = destination rank, a = array, pc = columns per rank
//
// Completion: every CTA fences system-wide and bumps a counter; the last CTA publishes the exchange's sequence number
// into flag slot [me] of every peer.  The consumer side is `mg_wait_flags_kernel`: one warp polling its own flag slots
// (acquire, system scope) until every source has published - the next kernel in the stream then reads the data.
// One buffer set per exchange of the step and in-order streams make reuse safe: a peer can only write exchange e of step
// t+1 after this rank has signalled exchange e-1 of step t+1, which it enqueues behind all its readers of step t.
#include <cstdint>
#include <cstring>
#include "common.cuh"
#include "kernels.h"

namespace gphm {

namespace {

constexpr int kFlagSlots = 64;                          // sequence numbers at the head of every exchange buffer
constexpr size_t kFlagBytes = kFlagSlots * sizeof(unsigned long long);
constexpr int kMaxPeers = 16;
constexpr int kMaxArrays = 4;

struct PeerTable { double* base[kMaxPeers]; };
struct ArrayTable { const double* in[kMaxArrays]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid: (ceil(cols / 32), ceil(rows / 32), k); block (32, 8)
__global__ void __launch_bounds__(256)
a2a_transpose_peer_kernel(ArrayTable arrays, int rows, int cols, int pc, PeerTable peers, int P, int me,
                          unsigned long long seq, unsigned int* done, unsigned int nblocks) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const double* __restrict__ X = arrays.in[blockIdx.z];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + ty + 8 * j, c = c0 + tx;
        tile[ty + 8 * j][tx] = (r < rows && c < cols) ? X[(size_t)r * cols + c] : 0.0;
    }
    __syncthreads();
    const size_t ldo = (size_t)P * rows;                // leading dimension of every destination array
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cg = c0 + ty + 8 * j, r = r0 + tx;    // 32 consecutive r = 256 contiguous bytes at the destination
        if (cg < cols && r < rows) {
            const int d = cg / pc, c = cg - d * pc;
            double* out = peers.base[d] + kFlagSlots + ((size_t)blockIdx.z * pc + c) * ldo + (size_t)me * rows + r;
            *out = tile[tx][ty + 8 * j];
        }
    }
    // completion: all stores of this CTA are visible system-wide before its ticket; the last CTA publishes the sequence number
    __threadfence_system();
    __syncthreads();
    if (tx == 0 && ty == 0) {
        const unsigned int ticket = atomicAdd(done, 1u);
        if (ticket == nblocks - 1) {
            *done = 0;                                  // ready for the next launch (stream order)
            __threadfence_system();
            for (int d = 0; d < P; ++d)
                st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[d]) + me, seq);
        }
    }
}

// One warp: lane s waits until source s has published `seq` (or later) into this rank's flag slot.  status[0] |= 1 on time-out.
__global__ void mg_wait_flags_kernel(const unsigned long long* flags, int P, unsigned long long seq, int* status) {
    const int s = threadIdx.x;
    if (s >= P) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + s) < seq) {
        __nanosleep(200);
        if (clock64() - t0 > (20ll << 30)) { if (status) atomicOr(status, 1); break; }     // ~10 s: a peer died
    }
}

}  // namespace

size_t peer_flag_bytes() { return kFlagBytes; }

int launch_a2a_transpose_peer(const double* const* in, int k, int rows, int cols, int pc, double* const* peer_bases, int P, int me,
                              unsigned long long seq, unsigned int* done, cudaStream_t st) {
    if (k < 1 || k > kMaxArrays || P < 1 || P > kMaxPeers || pc < 1 || cols != P * pc || me < 0 || me >= P) {
        set_last_error("a2a_transpose_peer: bad argument (k=%d P=%d cols=%d pc=%d)", k, P, cols, pc);
        return GPHM_EINVAL;
    }
    ArrayTable at; PeerTable pt;
    memset(&at, 0, sizeof(at)); memset(&pt, 0, sizeof(pt));
    for (int a = 0; a < k; ++a) at.in[a] = in[a];
    for (int d = 0; d < P; ++d) pt.base[d] = peer_bases[d];
    const dim3 grid((cols + 31) / 32, (rows + 31) / 32, k), block(32, 8);
    {
        LaunchScope scope(CAT_ELEMWISE, st, 0.0, 16.0 * k * rows * (double)cols);
        a2a_transpose_peer_kernel<<<grid, block, 0, st>>>(at, rows, cols, pc, pt, P, me, seq, done, grid.x * grid.y * grid.z);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

int launch_mg_wait_flags(const void* local_base, int P, unsigned long long seq, int* status, cudaStream_t st) {
    if (P < 1 || P > 32) { set_last_error("mg_wait_flags: P=%d", P); return GPHM_EINVAL; }
    {
        LaunchScope scope(CAT_ELEMWISE, st);
        mg_wait_flags_kernel<<<1, 32, 0, st>>>(static_cast<const unsigned long long*>(local_base), P, seq, status);
    }
    GPHM_LAUNCH_OK();
    return GPHM_OK;
}

}  // namespace gphm
