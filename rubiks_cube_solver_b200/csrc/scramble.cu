// K1 -- fused scramble from the solved cube (sm_100a).
//
// Replaces the reference's per-cube Python loop  init_state(); for a in seq: step(a)
// (cube_env.py:62-67 reset, :187-191 get_random_samples) for a whole batch: every
// thread owns one instance, keeps the cube as five registers of cubie bytes and
// applies the whole move sequence with PRMT byte permutes driven by per-move
// selector words in shared memory (see gen_tables.py) -- no branch on the move and
// no intermediate state in HBM.  The final state is expanded to the reference's
// sticker row (identical to chaining s[moveDefs[m]], py333.py:220-222), the
// face-uniformity verdict (py333.py:229-233) and the +-1 reward (cube_env.py:89-104).
//
// HBM traffic per instance: depth (moves in) + S (stickers out) + 1 + 4 bytes.
// One CTA handles one tile of 256 instances.  The tile's move bytes arrive in
// shared memory through one 1-D bulk asynchronous copy (TMA, mbarrier-signalled) and
// its sticker rows leave through one bulk store, so neither costs LSU wavefronts or
// address arithmetic (rows are 54 / 24 bytes: per-thread global stores would touch a
// different sector per lane).  Several CTAs per SM overlap load, compute and store.
// The kernel is bound by the ALU pipe (PRMT/LOP3) and the shared-memory pipe (six
// table words per move), not by HBM -- see DESIGN.md.
#include <cuda_runtime.h>
#include "cube_bulk.cuh"
#include "cube_kernels.h"
#include "cube_threads.cuh"

namespace {

#ifndef CUBE_SCRAMBLE_MIN_BLOCKS
#define CUBE_SCRAMBLE_MIN_BLOCKS 6      // 40 registers/thread, no spills; 6 CTAs x 8 warps per SM
#endif
constexpr int kTile = 256;              // instances per tile == threads per CTA
constexpr int kMaxStagedDepth = 160;    // deeper sequences are read straight from global

__host__ __device__ constexpr int round16(int x) { return (x + 15) & ~15; }

template <int SIZE>
struct ScrambleSmem {                   // carve-up of the dynamic shared memory
    using G = CubeGeom<SIZE>;
    static constexpr int kTable = 0;
    static constexpr int kCornerLut = kTable + G::MW * CUBE_MOVE_ROWS * 4;
    static constexpr int kEdgeLut = kCornerLut + 32 * 4;
    static constexpr int kBarrier = kEdgeLut + 64 * 4;
    static constexpr int kOut = round16(kBarrier + 16);
    static constexpr int kMoves = kOut + round16(kTile * G::S);
    __host__ __device__ static constexpr int bytes(int depth, bool staged)
    {
        return kMoves + (staged ? round16(kTile * depth) + 32 : 0);
    }
};

template <int SIZE>
__device__ __forceinline__ void load_tables(uint8_t* smem, int tid)
{
    using L = ScrambleSmem<SIZE>;
    uint32_t* s_tbl = reinterpret_cast<uint32_t*>(smem + L::kTable);
    uint32_t* s_clut = reinterpret_cast<uint32_t*>(smem + L::kCornerLut);
    uint32_t* s_elut = reinterpret_cast<uint32_t*>(smem + L::kEdgeLut);
    if (tid < CubeGeom<SIZE>::MW * CUBE_MOVE_ROWS) s_tbl[tid] = (SIZE == 3) ? kMoveWords3[tid] : kMoveWords2[tid];
    if (tid >= 128 && tid < 160) s_clut[tid - 128] = (SIZE == 3) ? kCornerColour3[tid - 128] : kCornerColour2[tid - 128];
    if (tid >= 160 && tid < 224) s_elut[tid - 160] = (SIZE == 3) ? kEdgeColour3[tid - 160] : 0u;
}

// one tile per CTA; move bytes staged in shared memory
template <int SIZE>
__global__ void __launch_bounds__(kTile, CUBE_SCRAMBLE_MIN_BLOCKS)
scramble_tile_kernel(const uint8_t* __restrict__ moves, long long n, int depth, uint8_t* __restrict__ out,
                     uint8_t* __restrict__ solved, float* __restrict__ reward,
                     unsigned long long* __restrict__ counters)
{
    using G = CubeGeom<SIZE>;
    using L = ScrambleSmem<SIZE>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t* s_tbl = reinterpret_cast<const uint32_t*>(smem + L::kTable);
    const uint32_t* s_clut = reinterpret_cast<const uint32_t*>(smem + L::kCornerLut);
    const uint32_t* s_elut = reinterpret_cast<const uint32_t*>(smem + L::kEdgeLut);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrier);
    uint8_t* s_out = smem + L::kOut;
    uint8_t* s_moves = smem + L::kMoves;

    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * kTile;
    const int cnt = (int)((n - base) < (long long)kTile ? (n - base) : (long long)kTile);
    const bool full = cnt == kTile;                       // bulk copies need 16-byte multiples: full tiles only
    const uint32_t move_bytes = (uint32_t)(kTile * depth);

    if (full && depth > 0 && tid == 0) {
        bulk::mbar_init(s_bar, 1);
        bulk::mbar_expect_tx(s_bar, move_bytes);
        bulk::load(s_moves, moves + base * depth, move_bytes, s_bar);
    }
    load_tables<SIZE>(smem, tid);
    if (!full) {                                          // ragged last tile: plain copies
        const long long byte0 = base * depth;
        const int nbytes = cnt * depth;
        for (int i = tid; i < nbytes; i += kTile) s_moves[i] = moves[byte0 + i];
    }
    __syncthreads();                                      // tables + barrier init visible to every thread
    if (full && depth > 0) bulk::mbar_wait(s_bar, 0);

    bool ok = false;
    if (tid < cnt) {
        CubieState st;
        cubie_init(st);
        scramble_run_staged<SIZE>(st, tid, depth, s_moves, s_tbl);
        ok = scramble_finish<SIZE>(st, tid, s_clut, s_elut, s_out);
        if (solved) solved[base + tid] = ok ? 1 : 0;
        if (reward) reward[base + tid] = ok ? 1.0f : -1.0f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if ((tid & 31) == 0 && bal && counters) atomicAdd(&counters[0], (unsigned long long)__popc(bal));
    if (blockIdx.x == 0 && tid == 0 && counters) atomicAdd(&counters[1], (unsigned long long)n);

    if (full) {
        bulk::fence_smem_writes();                        // rows written above -> visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            bulk::store(out + base * G::S, s_out, (uint32_t)(kTile * G::S));
            bulk::commit();
            bulk::wait_read_all();                        // shared memory must outlive the copy's reads
        }
    } else {
        __syncthreads();
        const long long byte0 = base * G::S;
        const int nbytes = cnt * G::S;
        for (int i = tid; i < nbytes; i += kTile) out[byte0 + i] = s_out[i];
    }
}

// deep sequences (depth > kMaxStagedDepth): persistent CTAs, moves read straight from global
template <int SIZE>
__global__ void __launch_bounds__(kTile, 4)
scramble_deep_kernel(const uint8_t* __restrict__ moves, long long n, int depth, uint8_t* __restrict__ out,
                     uint8_t* __restrict__ solved, float* __restrict__ reward,
                     unsigned long long* __restrict__ counters)
{
    using G = CubeGeom<SIZE>;
    using L = ScrambleSmem<SIZE>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t* s_tbl = reinterpret_cast<const uint32_t*>(smem + L::kTable);
    const uint32_t* s_clut = reinterpret_cast<const uint32_t*>(smem + L::kCornerLut);
    const uint32_t* s_elut = reinterpret_cast<const uint32_t*>(smem + L::kEdgeLut);
    uint8_t* s_out = smem + L::kOut;
    const int tid = threadIdx.x;
    load_tables<SIZE>(smem, tid);

    const long long n_tiles = (n + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * kTile;
        const int cnt = (int)((n - base) < (long long)kTile ? (n - base) : (long long)kTile);
        __syncthreads();            // tables visible; previous tile's copy-out finished
        bool ok = false;
        if (tid < cnt) {
            CubieState st;
            cubie_init(st);
            const uint8_t* row = moves + (base + tid) * depth;
            for (int k = 0; k < depth; ++k) {
                cubie_move<SIZE>(st, s_tbl, (uint32_t)__ldg(row + k) & 0xfu);
                if ((k & 7) == 7) { st.c0 = cubie_fold_twist(st.c0); st.c1 = cubie_fold_twist(st.c1); }
            }
            ok = scramble_finish<SIZE>(st, tid, s_clut, s_elut, s_out);
            if (solved) solved[base + tid] = ok ? 1 : 0;
            if (reward) reward[base + tid] = ok ? 1.0f : -1.0f;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0 && bal && counters) atomicAdd(&counters[0], (unsigned long long)__popc(bal));
        __syncthreads();
        const long long byte0 = base * G::S;               // multiple of 16
        const int nbytes = cnt * G::S;
        const int nvec = nbytes >> 4;
        int4* dst = reinterpret_cast<int4*>(out + byte0);
        for (int i = tid; i < nvec; i += kTile) __stcs(dst + i, reinterpret_cast<const int4*>(s_out)[i]);
        for (int i = (nvec << 4) + tid; i < nbytes; i += kTile) out[byte0 + i] = s_out[i];
    }
    if (blockIdx.x == 0 && tid == 0 && counters) atomicAdd(&counters[1], (unsigned long long)n);
}

template <int SIZE>
int launch_one(const uint8_t* moves, long long n, int depth, uint8_t* out, uint8_t* solved, float* reward,
               unsigned long long* counters, cudaStream_t stream)
{
    const bool staged = depth <= kMaxStagedDepth;
    const int smem = ScrambleSmem<SIZE>::bytes(depth, staged);
    const long long n_tiles = (n + kTile - 1) / kTile;
    if (staged) {
        auto kern = scramble_tile_kernel<SIZE>;
        static int configured_smem = -1;
        if (smem > configured_smem) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            // largest shared-memory carve-out, so occupancy is set by registers, not by the default split
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            configured_smem = smem;
        }
        kern<<<(unsigned)n_tiles, kTile, smem, stream>>>(moves, n, depth, out, solved, reward, counters);
    } else {
        auto kern = scramble_deep_kernel<SIZE>;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTile, smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        long long grid = (long long)cube::sm_count() * per_sm;
        if (grid > n_tiles) grid = n_tiles;
        kern<<<(unsigned)grid, kTile, smem, stream>>>(moves, n, depth, out, solved, reward, counters);
    }
    return (int)cudaGetLastError();
}

}  // namespace

namespace cube {

int launch_scramble(int size, const uint8_t* moves, long long n, int depth, uint8_t* states_out,
                    uint8_t* solved, float* reward, unsigned long long* counters, cudaStream_t stream)
{
    if (n == 0) return 0;
    if (size == 3) return launch_one<3>(moves, n, depth, states_out, solved, reward, counters, stream);
    return launch_one<2>(moves, n, depth, states_out, solved, reward, counters, stream);
}

}  // namespace cube
