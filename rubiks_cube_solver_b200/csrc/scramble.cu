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
// Moves of a 256-instance tile are staged through shared memory with coalesced
// 16-byte loads; sticker rows leave through a shared-memory tile with coalesced
// 16-byte stores (rows are 54 / 24 bytes, so per-thread global stores would
// touch a different sector per lane).
#include <cuda_runtime.h>
#include "cube_threads.cuh"
#include "cube_kernels.h"

namespace {

constexpr int kTile = 256;              // instances per tile == threads per CTA
constexpr int kMaxStagedDepth = 128;    // deeper sequences are read straight from global

__host__ __device__ constexpr int round16(int x) { return (x + 15) & ~15; }

template <int SIZE, bool STAGED>
__global__ void __launch_bounds__(kTile, 4)
scramble_kernel(const uint8_t* __restrict__ moves, long long n, int depth, uint8_t* __restrict__ out,
                uint8_t* __restrict__ solved, float* __restrict__ reward,
                unsigned long long* __restrict__ counters)
{
    using G = CubeGeom<SIZE>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* s_tbl = reinterpret_cast<uint32_t*>(smem);
    uint32_t* s_clut = s_tbl + G::MW * CUBE_MOVE_ROWS;
    uint32_t* s_elut = s_clut + 32;
    uint8_t* s_out = reinterpret_cast<uint8_t*>(s_elut + 32);
    uint8_t* s_moves = s_out + round16(kTile * G::S);
    __shared__ unsigned int s_solved_count;

    const int tid = threadIdx.x;
    for (int i = tid; i < G::MW * CUBE_MOVE_ROWS; i += kTile)
        s_tbl[i] = (SIZE == 3) ? kMoveWords3[i] : kMoveWords2[i];
    if (tid < 32) {
        s_clut[tid] = (SIZE == 3) ? kCornerColour3[tid] : kCornerColour2[tid];
        s_elut[tid] = (SIZE == 3) ? kEdgeColour3[tid] : 0u;
    }
    if (tid == 0) s_solved_count = 0;

    const long long n_tiles = (n + kTile - 1) / kTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * kTile;
        const int cnt = (int)((n - base) < (long long)kTile ? (n - base) : (long long)kTile);
        __syncthreads();            // tables visible; previous tile's shared-memory reads finished

        if (STAGED) {
            const long long byte0 = base * depth;          // multiple of 16 (kTile = 256)
            const int nbytes = cnt * depth;
            const int nvec = nbytes >> 4;
            const int4* src = reinterpret_cast<const int4*>(moves + byte0);
            for (int i = tid; i < nvec; i += kTile) reinterpret_cast<int4*>(s_moves)[i] = __ldcs(src + i);
            for (int i = (nvec << 4) + tid; i < nbytes; i += kTile) s_moves[i] = moves[byte0 + i];
            __syncthreads();
        }

        bool ok = false;
        if (tid < cnt) {
            CubieState st;
            cubie_init(st);
            if (STAGED) {
                scramble_run_staged<SIZE>(st, tid, depth, s_moves, s_tbl);
            } else {
                const uint8_t* row = moves + (base + tid) * depth;
                for (int k = 0; k < depth; ++k) {
                    cubie_move<SIZE>(st, s_tbl, (uint32_t)__ldg(row + k) & 0xfu);
                    if ((k & 7) == 7) { st.c0 = cubie_fold_twist(st.c0); st.c1 = cubie_fold_twist(st.c1); }
                }
            }
            ok = scramble_finish<SIZE>(st, tid, s_clut, s_elut, s_out);
            if (solved) solved[base + tid] = ok ? 1 : 0;
            if (reward) reward[base + tid] = ok ? 1.0f : -1.0f;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0 && bal) atomicAdd(&s_solved_count, (unsigned)__popc(bal));
        __syncthreads();

        {   // coalesced copy-out of the sticker tile
            const long long byte0 = base * G::S;           // multiple of 16
            const int nbytes = cnt * G::S;
            const int nvec = nbytes >> 4;
            int4* dst = reinterpret_cast<int4*>(out + byte0);
            for (int i = tid; i < nvec; i += kTile) __stcs(dst + i, reinterpret_cast<const int4*>(s_out)[i]);
            for (int i = (nvec << 4) + tid; i < nbytes; i += kTile) out[byte0 + i] = s_out[i];
        }
    }
    __syncthreads();
    if (tid == 0 && counters) {
        if (s_solved_count) atomicAdd(&counters[0], (unsigned long long)s_solved_count);
        if (blockIdx.x == 0) atomicAdd(&counters[1], (unsigned long long)n);
    }
}

template <int SIZE, bool STAGED>
int launch_one(const uint8_t* moves, long long n, int depth, uint8_t* out, uint8_t* solved, float* reward,
               unsigned long long* counters, cudaStream_t stream)
{
    using G = CubeGeom<SIZE>;
    auto kern = scramble_kernel<SIZE, STAGED>;
    const int smem = (G::MW * CUBE_MOVE_ROWS + 64) * 4 + round16(kTile * G::S)
                   + (STAGED ? round16(kTile * depth) + 32 : 0);
    static int configured_smem = -1;
    if (smem > configured_smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured_smem = smem;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTile, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    const long long n_tiles = (n + kTile - 1) / kTile;
    long long grid = (long long)cube::sm_count() * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, kTile, smem, stream>>>(moves, n, depth, out, solved, reward, counters);
    return (int)cudaGetLastError();
}

}  // namespace

namespace cube {

int launch_scramble(int size, const uint8_t* moves, long long n, int depth, uint8_t* states_out,
                    uint8_t* solved, float* reward, unsigned long long* counters, cudaStream_t stream)
{
    if (n == 0) return 0;
    const bool staged = depth <= kMaxStagedDepth;
    if (size == 3)
        return staged ? launch_one<3, true>(moves, n, depth, states_out, solved, reward, counters, stream)
                      : launch_one<3, false>(moves, n, depth, states_out, solved, reward, counters, stream);
    return staged ? launch_one<2, true>(moves, n, depth, states_out, solved, reward, counters, stream)
                  : launch_one<2, false>(moves, n, depth, states_out, solved, reward, counters, stream);
}

}  // namespace cube
