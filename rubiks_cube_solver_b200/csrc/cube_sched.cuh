// Tile scheduling for the persistent kernels (K1p, K2p): static bulk, dynamic tail.
//
// SMs do not get equal shares of HBM bandwidth (die-local L2, channel contention): with a purely
// static split the fastest SM of a bandwidth-bound launch is done after ~80 % of the kernel's time.
// A purely dynamic split does not work either -- same-address atomics retire at ~0.4 per ns while
// the kernels retire up to 2 tiles per ns.  So the first (1 - 1/kTailDiv) of the tiles is strided
// over the warps statically (neighbours in time are neighbours in memory), and the tail is claimed
// by whoever gets there, kClaimChunk tiles at a time, from kRanges counters on separate cache
// lines (a warp starts in its home sub-range and then steals from the others).  Claims are issued
// one chunk ahead, so the atomic's round trip is not waited for, and the kernel's last CTA puts the
// counters back to zero: a launch needs no memset and leaves no state behind.  Launches take their
// counters from a ring of slots, so kernels in flight on different streams never share one.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sched {

constexpr int kClaimChunk = 2;         // tiles per claim (the default; a kernel with long tiles claims one at a time)
constexpr int kRanges = 8;
constexpr int kTailDiv = 4;            // the last 1/4 of the tiles is claimed dynamically (measured: 8 -> +10 %, 4 -> +13 % on K2p)
// Slots are handed out round-robin from a ring: two launches share a slot only if kSlots launches of persistent
// kernels are in flight AT THE SAME TIME on one device (across streams, host threads and concurrently replayed
// CUDA graphs, which bake their slot in).  A kernel holds its slot for its own duration only, launches of one
// stream serialise, so that takes > 1024 concurrently running streams / graphs -- far beyond what fits on 148
// SMs; the ring costs 1.2 MB of device memory.
constexpr int kSlots = 1024;
constexpr int kExhausted = 0x7fffff00;

struct Slot { unsigned next[kRanges][32]; unsigned ctas_done; unsigned tail_div; unsigned pad[30]; };

// host: the device address of a fresh slot for the current device (nullptr on error)
Slot* claim_slot();
// host: tail divisor (kTailDiv; CUBE_TAIL_DIV overrides it for A/B runs, 0 = fully static)
int tail_div();

// per-warp state, identical in all lanes except `pending` (lane 0 only)
struct WarpTiles {
    Slot* slot;
    int n_tiles, total_warps, tail0;   // tiles [tail0, n_tiles) are dynamic
    int next_static;
    int r, visited;                    // tail sub-range being worked on; visited == kRanges: nothing left anywhere
    int cur, end;                      // the claimed chunk in hand
    int chunk;                         // tiles per claim
    unsigned pending;                  // lane 0: raw result of the claim issued ahead in sub-range r

    __device__ __forceinline__ int lo(int range) const
    {
        return tail0 + (int)((long long)range * (n_tiles - tail0) / kRanges);
    }
    __device__ __forceinline__ void claim_ahead(int lane)
    {
        if (lane == 0) pending = atomicAdd(&slot->next[r][0], (unsigned)chunk);
    }

    __device__ __forceinline__ void init(Slot* s, int n, int warps_per_cta, int warp, int lane, int tail_div,
                                         int claim_chunk = kClaimChunk)
    {
        slot = s; n_tiles = n; total_warps = (int)gridDim.x * warps_per_cta; chunk = claim_chunk;
        const int gid = (int)blockIdx.x * warps_per_cta + warp;
        tail0 = tail_div > 0 ? n - n / tail_div : n;
        next_static = gid;
        r = gid % kRanges; visited = 0; cur = end = 0; pending = 0;
        if (next_static >= tail0) claim_ahead(lane);                  // no static share at all
    }

    // next tile of this warp, kExhausted once everything is handed out
    __device__ __forceinline__ int pop(int lane)
    {
        if (next_static < tail0) {
            const int t = next_static;
            next_static += total_warps;
            if (next_static >= tail0) claim_ahead(lane);              // last static tile: first claim goes out now
            return t;
        }
        while (cur >= end) {
            if (visited >= kRanges) return kExhausted;
            const unsigned got = __shfl_sync(0xffffffffu, pending, 0);
            const int hi = lo(r + 1);
            const long long base = (long long)lo(r) + got;
            if (base < hi) {
                cur = (int)base;
                end = cur + chunk < hi ? cur + chunk : hi;
            } else {
                // sub-range used up: look at ALL counters at once (lanes 0..7, one round trip) and move to the
                // next one that still has tiles; none left = done.  (Probing them one claim after the other cost
                // up to kRanges atomic round trips per warp at the end of a kernel: ~8 us during which the CTA
                // holds its SM and the next launch of the stream waits for the grid to complete.)
                bool left = false;
                if (lane < kRanges) left = lo(lane) + (long long)__ldcg(&slot->next[lane][0]) < lo(lane + 1);
                const unsigned open = __ballot_sync(0xffffffffu, left) & ((1u << kRanges) - 1u);
                if (open == 0) { visited = kRanges; return kExhausted; }
                const unsigned rot = (open >> (r + 1)) | (open << (kRanges - r - 1));     // bit 0 = range r + 1
                r = (r + 1 + (__ffs(rot & ((1u << kRanges) - 1u)) - 1)) % kRanges;
            }
            claim_ahead(lane);
        }
        return cur++;
    }
};

// end of kernel, after a __syncthreads(): the last CTA of the grid re-arms the slot
__device__ __forceinline__ void release(Slot* slot)
{
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&slot->ctas_done, 1u) == gridDim.x - 1) {
            for (int i = 0; i < kRanges; ++i) slot->next[i][0] = 0;
            slot->ctas_done = 0;
            __threadfence();
        }
    }
}

}  // namespace sched
