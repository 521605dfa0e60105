// K2 -- moves applied to resident sticker rows (sm_100a): the reference's
// CubeEnv.step (cube_env.py:71-111) for a batch, for ANY byte content of the
// rows: new[i] = old[moveDefs[a][i]] (py333.py:220-222; py222 doMove), then
// face uniformity (py333.py:229-233) and the +-1 reward (cube_env.py:89-104).
//
// A tile of 256 rows is staged in shared memory with coalesced 16-byte
// loads, each thread turns its own row in place -- a face turn is five
// (3x3x3) or three (2x2x2) sticker 4-cycles whose byte offsets come from a
// [cycle][move] table, so there is no branch on the move -- and the tile goes
// back with coalesced 16-byte stores.  depth = 1 is `step`; depth > 1 walks
// several moves without leaving shared memory.  HBM traffic per instance:
// 2*S + depth + 1 + 4 bytes.
#include <cuda_runtime.h>
#include "cube_threads.cuh"
#include "cube_kernels.h"

namespace {

constexpr int kTile = 256;

__host__ __device__ constexpr int round16(int x) { return (x + 15) & ~15; }

template <int SIZE>
__global__ void __launch_bounds__(kTile, 4)
walk_kernel(const uint8_t* in, const uint8_t* __restrict__ moves, long long n, int depth,
            uint8_t* out, uint8_t* __restrict__ solved, float* __restrict__ reward,
            unsigned long long* __restrict__ counters)
{
    using G = CubeGeom<SIZE>;
    __shared__ uint32_t s_cyc[G::NCYC * CUBE_MOVE_ROWS];
    __shared__ __align__(16) uint8_t s_rows[round16(kTile * G::S)];
    __shared__ unsigned int s_solved_count;

    const int tid = threadIdx.x;
    for (int i = tid; i < G::NCYC * CUBE_MOVE_ROWS; i += kTile) s_cyc[i] = (SIZE == 3) ? kCycles3[i] : kCycles2[i];
    if (tid == 0) s_solved_count = 0;

    const long long n_tiles = (n + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * kTile;
        const int cnt = (int)((n - base) < (long long)kTile ? (n - base) : (long long)kTile);
        const long long byte0 = base * G::S;               // multiple of 16
        const int nbytes = cnt * G::S;
        const int nvec = nbytes >> 4;
        __syncthreads();
        {
            const int4* src = reinterpret_cast<const int4*>(in + byte0);
            for (int i = tid; i < nvec; i += kTile) reinterpret_cast<int4*>(s_rows)[i] = __ldcs(src + i);
            for (int i = (nvec << 4) + tid; i < nbytes; i += kTile) s_rows[i] = in[byte0 + i];
        }
        __syncthreads();

        bool ok = false;
        if (tid < cnt) {
            uint8_t* row = s_rows + G::S * tid;
            const uint8_t* mrow = moves + (base + tid) * depth;
            for (int k = 0; k < depth; ++k) {
                const uint32_t m = (uint32_t)__ldcs(mrow + k) & 0xfu;    // rows >= A: trivial cycles
                walk_turn<SIZE>(row, m, s_cyc);
            }
            ok = row_solved<SIZE>(row);
            if (solved) solved[base + tid] = ok ? 1 : 0;
            if (reward) reward[base + tid] = ok ? 1.0f : -1.0f;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0 && bal) atomicAdd(&s_solved_count, (unsigned)__popc(bal));
        __syncthreads();
        {
            int4* dst = reinterpret_cast<int4*>(out + byte0);
            for (int i = tid; i < nvec; i += kTile) __stcs(dst + i, reinterpret_cast<const int4*>(s_rows)[i]);
            for (int i = (nvec << 4) + tid; i < nbytes; i += kTile) out[byte0 + i] = s_rows[i];
        }
    }
    __syncthreads();
    if (tid == 0 && counters) {
        if (s_solved_count) atomicAdd(&counters[0], (unsigned long long)s_solved_count);
        if (blockIdx.x == 0) atomicAdd(&counters[1], (unsigned long long)n);
    }
}

// solved / reward of resident rows, nothing else (cube_solved)
template <int SIZE>
__global__ void __launch_bounds__(kTile, 4)
solved_kernel(const uint8_t* __restrict__ in, long long n, uint8_t* __restrict__ solved,
              float* __restrict__ reward, unsigned long long* __restrict__ counters)
{
    using G = CubeGeom<SIZE>;
    __shared__ __align__(16) uint8_t s_rows[round16(kTile * G::S)];
    __shared__ unsigned int s_solved_count;
    const int tid = threadIdx.x;
    if (tid == 0) s_solved_count = 0;
    const long long n_tiles = (n + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * kTile;
        const int cnt = (int)((n - base) < (long long)kTile ? (n - base) : (long long)kTile);
        const long long byte0 = base * G::S;
        const int nbytes = cnt * G::S;
        const int nvec = nbytes >> 4;
        __syncthreads();
        const int4* src = reinterpret_cast<const int4*>(in + byte0);
        for (int i = tid; i < nvec; i += kTile) reinterpret_cast<int4*>(s_rows)[i] = __ldcs(src + i);
        for (int i = (nvec << 4) + tid; i < nbytes; i += kTile) s_rows[i] = in[byte0 + i];
        __syncthreads();
        bool ok = false;
        if (tid < cnt) {
            ok = row_solved<SIZE>(s_rows + G::S * tid);
            if (solved) solved[base + tid] = ok ? 1 : 0;
            if (reward) reward[base + tid] = ok ? 1.0f : -1.0f;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0 && bal) atomicAdd(&s_solved_count, (unsigned)__popc(bal));
    }
    __syncthreads();
    if (tid == 0 && counters) {
        if (s_solved_count) atomicAdd(&counters[0], (unsigned long long)s_solved_count);
        if (blockIdx.x == 0) atomicAdd(&counters[1], (unsigned long long)n);
    }
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

template <typename K>
long long grid_for(K kern, long long n_tiles, int threads)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    long long grid = (long long)cube::sm_count() * per_sm;
    return grid < n_tiles ? grid : n_tiles;
}

}  // namespace

namespace cube {

int launch_walk(int size, const uint8_t* states_in, const uint8_t* moves, long long n, int depth,
                uint8_t* states_out, uint8_t* solved, float* reward, unsigned long long* counters,
                cudaStream_t stream)
{
    if (n == 0) return 0;
    const long long n_tiles = (n + kTile - 1) / kTile;
    if (size == 3) {
        const long long grid = grid_for(walk_kernel<3>, n_tiles, kTile);
        walk_kernel<3><<<(unsigned)grid, kTile, 0, stream>>>(states_in, moves, n, depth, states_out, solved,
                                                            reward, counters);
    } else {
        const long long grid = grid_for(walk_kernel<2>, n_tiles, kTile);
        walk_kernel<2><<<(unsigned)grid, kTile, 0, stream>>>(states_in, moves, n, depth, states_out, solved,
                                                            reward, counters);
    }
    return (int)cudaGetLastError();
}

int launch_solved(int size, const uint8_t* states, long long n, uint8_t* solved, float* reward,
                  unsigned long long* counters, cudaStream_t stream)
{
    if (n == 0) return 0;
    const long long n_tiles = (n + kTile - 1) / kTile;
    if (size == 3) {
        const long long grid = grid_for(solved_kernel<3>, n_tiles, kTile);
        solved_kernel<3><<<(unsigned)grid, kTile, 0, stream>>>(states, n, solved, reward, counters);
    } else {
        const long long grid = grid_for(solved_kernel<2>, n_tiles, kTile);
        solved_kernel<2><<<(unsigned)grid, kTile, 0, stream>>>(states, n, solved, reward, counters);
    }
    return (int)cudaGetLastError();
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
