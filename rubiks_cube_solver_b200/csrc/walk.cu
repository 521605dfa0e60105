// K2 -- moves applied to resident sticker rows (sm_100a): the reference's
// CubeEnv.step (cube_env.py:71-111) for a batch, for ANY byte content of the
// rows: new[i] = old[moveDefs[a][i]] (py333.py:220-222; py222 doMove), then
// face uniformity (py333.py:229-233) and the +-1 reward (cube_env.py:89-104).
//
// One CTA turns one tile of 256 rows.  The tile comes in through one 1-D bulk
// asynchronous copy (TMA) and leaves through one bulk store; in between each
// thread turns its own row in place in shared memory -- a face turn is five
// (3x3x3) or three (2x2x2) sticker 4-cycles whose byte offsets come from a
// [cycle][move] table, so there is no branch on the move.  3x3x3 rows are 13.5
// words, so even rows go to warps 0-3 and odd rows to warps 4-7: lanes of a warp
// are then 27 words apart and row-relative accesses spread over all 32 banks.
// depth = 1 is `step`; depth > 1 walks several moves without leaving shared
// memory; depth = 0 with no output is `cube_solved`.
// HBM traffic per instance: 2*S + depth + 1 + 4 bytes.
#include <cuda_runtime.h>
#include "cube_bulk.cuh"
#include "cube_kernels.h"
#include "cube_threads.cuh"

namespace {

constexpr int kTile = 256;

__host__ __device__ constexpr int round16(int x) { return (x + 15) & ~15; }

// FULL: all tiles of the launch have 256 rows (bulk copies); !FULL: the ragged last tile.
template <int SIZE, bool FULL>
__global__ void __launch_bounds__(kTile, 6)
walk_tile_kernel(const uint8_t* in, const uint8_t* __restrict__ moves, long long n, long long tile0, int depth,
                 uint8_t* out, uint8_t* __restrict__ solved, float* __restrict__ reward,
                 unsigned long long* __restrict__ counters)
{
    using G = CubeGeom<SIZE>;
    __shared__ __align__(128) uint8_t s_rows[round16(kTile * G::S)];
    __shared__ uint32_t s_cyc[G::NCYC * CUBE_MOVE_ROWS];
    __shared__ uint8_t s_flags[kTile];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x;
    const int row = (SIZE == 3) ? 2 * (((tid >> 5) & 3) * 32 + (tid & 31)) + (tid >> 7) : tid;
    const long long base = (tile0 + blockIdx.x) * kTile;
    const int cnt = FULL ? kTile : (int)(n - base);
    constexpr uint32_t kTileBytes = kTile * G::S;

    if (FULL && tid == 0) {
        bulk::mbar_init(&s_bar, 1);
        bulk::mbar_expect_tx(&s_bar, kTileBytes);
        bulk::load(s_rows, in + base * G::S, kTileBytes, &s_bar);
    }
    if (tid < G::NCYC * CUBE_MOVE_ROWS) s_cyc[tid] = (SIZE == 3) ? kCycles3[tid] : kCycles2[tid];
    if (!FULL) {
        const long long byte0 = base * G::S;
        for (int i = tid; i < cnt * G::S; i += kTile) s_rows[i] = in[byte0 + i];
    }
    // the moves of this thread's row, fetched while the tile is in flight (depth 1: one byte)
    uint32_t first_move = CUBE_NOOP_MOVE;
    const uint8_t* mrow = moves + (base + row) * depth;
    if (row < cnt && depth > 0) first_move = (uint32_t)__ldcs(mrow) & 0xfu;
    __syncthreads();
    if (FULL) bulk::mbar_wait(&s_bar, 0);

    bool ok = false;
    if (row < cnt) {
        uint8_t* r = s_rows + G::S * row;
        if (depth > 0) walk_turn<SIZE>(r, first_move, s_cyc);
        for (int k = 1; k < depth; ++k) walk_turn<SIZE>(r, (uint32_t)__ldcs(mrow + k) & 0xfu, s_cyc);
        ok = row_solved<SIZE>(r);
        s_flags[row] = ok ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if ((tid & 31) == 0 && bal && counters) atomicAdd(&counters[0], (unsigned long long)__popc(bal));
    if (blockIdx.x == 0 && tid == 0 && counters)
        atomicAdd(&counters[1], (unsigned long long)(FULL ? (long long)gridDim.x * kTile : (long long)cnt));

    if (FULL && out) bulk::fence_smem_writes();
    __syncthreads();
    if (out) {
        if (FULL) {
            if (tid == 0) {
                bulk::store(out + base * G::S, s_rows, kTileBytes);
                bulk::commit();
            }
        } else {
            const long long byte0 = base * G::S;
            for (int i = tid; i < cnt * G::S; i += kTile) out[byte0 + i] = s_rows[i];
        }
    }
    if (tid < cnt) {
        const bool row_ok = s_flags[tid] != 0;
        if (solved) solved[base + tid] = row_ok ? 1 : 0;
        if (reward) reward[base + tid] = row_ok ? 1.0f : -1.0f;
    }
    if (FULL && out && tid == 0) bulk::wait_read_all();
}

// out-of-range action scan (the reference raises IndexError, cube_env.py:86,96)
__global__ void __launch_bounds__(256)
validate_kernel(const uint8_t* __restrict__ actions, long long count, unsigned n_actions,
                unsigned long long* __restrict__ counters)
{
    unsigned bad = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nvec = count >> 4;
    const uint4* v = reinterpret_cast<const uint4*>(actions);
    const uint32_t k = (0x80u - n_actions) * 0x01010101u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 q = __ldcs(v + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            bad += __popc((((w[j] & 0x7f7f7f7fu) + k) | w[j]) & 0x80808080u);
    }
    for (long long i = (nvec << 4) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        bad += actions[i] >= n_actions;
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&counters[2], (unsigned long long)bad);
}

template <int SIZE>
int launch_tiles(const uint8_t* in, const uint8_t* moves, long long n, int depth, uint8_t* out, uint8_t* solved,
                 float* reward, unsigned long long* counters, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(walk_tile_kernel<SIZE, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    const long long full_tiles = n / kTile;
    if (full_tiles > 0)
        walk_tile_kernel<SIZE, true><<<(unsigned)full_tiles, kTile, 0, stream>>>(in, moves, n, 0, depth, out, solved,
                                                                               reward, counters);
    if (full_tiles * kTile < n)
        walk_tile_kernel<SIZE, false><<<1, kTile, 0, stream>>>(in, moves, n, full_tiles, depth, out, solved, reward,
                                                              counters);
    return (int)cudaGetLastError();
}

}  // namespace

namespace cube {

int launch_walk(int size, const uint8_t* states_in, const uint8_t* moves, long long n, int depth,
                uint8_t* states_out, uint8_t* solved, float* reward, unsigned long long* counters,
                cudaStream_t stream)
{
    if (n == 0) return 0;
    if (size == 3) return launch_tiles<3>(states_in, moves, n, depth, states_out, solved, reward, counters, stream);
    return launch_tiles<2>(states_in, moves, n, depth, states_out, solved, reward, counters, stream);
}

int launch_solved(int size, const uint8_t* states, long long n, uint8_t* solved, float* reward,
                  unsigned long long* counters, cudaStream_t stream)
{
    if (n == 0) return 0;
    if (size == 3) return launch_tiles<3>(states, nullptr, n, 0, nullptr, solved, reward, counters, stream);
    return launch_tiles<2>(states, nullptr, n, 0, nullptr, solved, reward, counters, stream);
}

int launch_validate(int size, const uint8_t* actions, long long count, unsigned long long* counters,
                    cudaStream_t stream)
{
    if (count == 0) return 0;
    long long blocks = (count / 16 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    validate_kernel<<<(unsigned)blocks, 256, 0, stream>>>(actions, count, size == 3 ? 12u : 6u, counters);
    return (int)cudaGetLastError();
}

}  // namespace cube
