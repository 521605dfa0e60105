// One-shot all-reduce (SUM, int64) of the solved / produced counters over NVLink peer memory (sm_100a).
//
// The path's only collective ("NCCL used only for the final solved-count and reward reductions")
// moves a few hundred bytes, so its cost is pure latency: an ncclAllReduce of 640 bytes is a launch plus a
// multi-hop low-latency protocol.  Here every rank owns an EXCHANGE BUFFER that all ranks of the box have mapped
// (torch's symmetric memory: cuMem allocations exchanged once at start-up); one small kernel per rank
//   1. stores its values into slot [rank] of EVERY rank's buffer (plain stores through NVLink / NVSwitch),
//   2. fences and raises its flag in every rank's buffer (release, system scope),
//   3. waits until all flags in its OWN buffer have reached this call's epoch (acquire), and
//   4. sums the slots locally -- every rank adds the same numbers in the same order: identical results.
// One NVLink store latency instead of a ring / tree: the whole exchange is one hop.  Flags carry a call
// epoch that only grows, so nothing is ever reset; slots are double-buffered by the epoch's parity, because a
// fast rank may already be pushing call e + 1 while a slow one still sums call e (it cannot reach e + 2 before
// the slow rank has signalled e + 1, which it does after its sums).
#include <cstdint>
#include <cuda_runtime.h>
#include "cube_kernels.h"

namespace {

constexpr int kMaxRanks = CUBE_PEER_MAX_RANKS;
constexpr int kFlagStride = 32;                                  // flags on separate 128-byte lines
constexpr int kFlagWords = kMaxRanks * kFlagStride;

struct PeerBufs { unsigned long long base[kMaxRanks]; };

__device__ __forceinline__ void store_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned load_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256, 1)
peer_allreduce_i64_kernel(PeerBufs bufs, int world, int rank, long long* __restrict__ values, int n, int cap, unsigned epoch)
{
    const int tid = threadIdx.x;
    const size_t slots_at = (size_t)kFlagWords * 4 + (size_t)(epoch & 1u) * kMaxRanks * cap * 8;
    // 1. my values into slot [rank] of every rank's buffer
    for (int p = 0; p < world; ++p) {
        long long* dst = reinterpret_cast<long long*>(bufs.base[p] + slots_at) + (size_t)rank * cap;
        for (int i = tid; i < n; i += blockDim.x) dst[i] = values[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. tell every rank, 3. wait for every rank
    if (tid < world) {
        store_release_sys(reinterpret_cast<unsigned*>(bufs.base[tid]) + rank * kFlagStride, epoch);
        const unsigned* mine = reinterpret_cast<const unsigned*>(bufs.base[rank]) + tid * kFlagStride;
        while ((int)(load_acquire_sys(mine) - epoch) < 0) {}
    }
    __syncthreads();
    // 4. the same sum on every rank
    const long long* slots = reinterpret_cast<const long long*>(bufs.base[rank] + slots_at);
    for (int i = tid; i < n; i += blockDim.x) {
        long long s = 0;
        for (int p = 0; p < world; ++p) s += __ldcv(slots + (size_t)p * cap + i);
        values[i] = s;
    }
}

}  // namespace

namespace cube {

size_t peer_buffer_bytes(int cap) { return (size_t)kFlagWords * 4 + 2 * (size_t)kMaxRanks * cap * 8; }

int launch_peer_allreduce_i64(int world, int rank, const uint64_t* peer_bufs, long long* values, int n, int cap,
                              unsigned epoch, cudaStream_t stream)
{
    PeerBufs b{};
    for (int p = 0; p < world; ++p) b.base[p] = peer_bufs[p];
    peer_allreduce_i64_kernel<<<1, 256, 0, stream>>>(b, world, rank, values, n, cap, epoch);
    return (int)cudaGetLastError();
}

}  // namespace cube
