// One-hot -> sticker rows for 2x2x2: CubeEnv.state_to_sim_state (cube_env.py:154-175), i.e.
// np.where(row == 1)[0][0] per cubelet row, then py222 getStickers.  The reference raises
// NotImplementedError for 3x3x3 (cube_env.py:171-172) -- its corner encoding is lossy -- and so
// does this library for that encoding; the opt-in exact encoding has an inverse (below).
#include <cuda_runtime.h>
#include "cube_threads.cuh"
#include "cube_kernels.h"

namespace {

template <typename T> __device__ __forceinline__ bool is_one(T v);
template <> __device__ __forceinline__ bool is_one<uint8_t>(uint8_t v) { return v == 1; }
template <> __device__ __forceinline__ bool is_one<uint16_t>(uint16_t v) { return v == 0x3f80; }     // bf16 1.0
template <> __device__ __forceinline__ bool is_one<float>(float v) { return v == 1.0f; }

// A block takes 64 instances: their one-hot rows (64 x 147 elements, contiguous) arrive in shared memory by
// coalesced 16-byte loads, one thread per (instance, cubelet) finds the column of its row's 1, one thread per
// (instance, sticker) fills the 64 x 24 sticker tile, which leaves coalesced.  (Round 1's version, one thread per
// instance reading 147 elements one by one, was instruction-queue-bound: ncu lg_throttle 113 stall cycles per
// issue, 17 % of DRAM throughput.)
constexpr int kDec2Tile = 64;

template <typename T>
__global__ void __launch_bounds__(256)
decode2_kernel(const T* __restrict__ onehot, long long n, uint8_t* __restrict__ out)
{
    __shared__ __align__(16) T s_oh[kDec2Tile * 147];
    __shared__ uint8_t s_col[kDec2Tile * 7];
    __shared__ uint8_t s_at[kDec2Tile * 7];
    __shared__ __align__(16) uint8_t s_rows[kDec2Tile * 24];
    __shared__ uint8_t s_piece[24];                        // sticker -> position | k << 4 (0x80: the fixed DBL cubie's colour)
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * kDec2Tile;
    const int cnt = (int)((n - base) < (long long)kDec2Tile ? (n - base) : (long long)kDec2Tile);
    if (tid < 24) {
        uint8_t v = 0x80 | (tid == 14 ? 3 : tid == 18 ? 4 : 5);          // stickers 14, 18, 23 of the fixed cubie
        for (int pos = 0; pos < 7; ++pos)
            for (int k = 0; k < 3; ++k)
                if (kPieceDefs2[pos * 3 + k] == tid) v = (uint8_t)(pos | k << 4);
        s_piece[tid] = v;
    }
    {   // 64 * 147 elements: the tile starts on a 16-byte boundary (64 * 147 * sizeof(T) is a multiple of 16)
        const long long e0 = base * 147;
        const int n_el = cnt * 147;
        constexpr int V = 16 / (int)sizeof(T);
        const int nvec = n_el / V;
        const int4* src = reinterpret_cast<const int4*>(onehot + e0);
        for (int i = tid; i < nvec; i += 256) reinterpret_cast<int4*>(s_oh)[i] = __ldcs(src + i);
        for (int i = nvec * V + tid; i < n_el; i += 256) s_oh[i] = onehot[e0 + i];
    }
    __syncthreads();
    for (int it = tid; it < cnt * 7; it += 256) {
        const T* row = s_oh + it * 21;
        int col = 0;                                       // argmax semantics: first 1, else column 0
        for (int c = 20; c >= 0; --c) if (is_one<T>(row[c])) col = c;
        s_col[it] = (uint8_t)col;
    }
    __syncthreads();
    // which cubelet sits at every position, and how it is turned: cubelet | ori << 4 (0xff: nobody; the LAST
    // cubelet that names a position wins, like the reference's loop over cubelets)
    for (int it = tid; it < cnt * 7; it += 256) {
        const int r = it / 7, pos = it - 7 * r;
        uint32_t at = 0xffu;
        for (int cubelet = 0; cubelet < 7; ++cubelet) {
            const uint32_t col = s_col[r * 7 + cubelet], p3 = (col * 11u) >> 5;       // col / 3 for col < 32
            if ((int)p3 == pos) at = (uint32_t)cubelet | (col - 3u * p3) << 4;
        }
        s_at[it] = (uint8_t)at;
    }
    __syncthreads();
    for (int it = tid; it < cnt * 24; it += 256) {
        const int r = it / 24, st = it - 24 * r;
        const uint32_t pk = s_piece[st];
        uint32_t colour = pk & 7u;
        if (!(pk & 0x80u)) {
            // sticker st is the k-th sticker of position pos
            const uint32_t at = s_at[r * 7 + (pk & 15u)], k = pk >> 4;
            uint32_t j = k + 3u - (at >> 4);                                          // np.roll(home, ori)[k] = home[(k - ori) mod 3]
            j -= j >= 3u ? 3u : 0u;
            colour = at == 0xffu ? 0u : kHomeColour2[(at & 15u) * 3 + j];
        }
        s_rows[it] = (uint8_t)colour;
    }
    __syncthreads();
    uint8_t* dst = out + base * 24;
    const int nbytes = cnt * 24, nvec = nbytes >> 4;
    for (int i = tid; i < nvec; i += 256) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(s_rows)[i];
    for (int i = (nvec << 4) + tid; i < nbytes; i += 256) dst[i] = s_rows[i];
}

// ---- 3x3x3, EXACT encoding (opt-in): one-hot [n, 20, 24] -> sticker rows [n, 54] ----
// The reference has no inverse for 3x3x3 (cube_env.py:171-172) because its corner table is lossy; the exact
// encoding (cube_b200.h, gen_tables.py) is a bijection, so the inverse exists: row q's column names the
// (piece, ori) in slot q, whose stickers show np.roll(home colours of the piece, ori); centres never move.
// A block takes 64 instances: phase 1 finds the column of the 1 of each of its 64 x 20 one-hot rows (one
// row = 24 elements = a few 16-byte vectors per thread, consecutive threads read consecutive rows), phase 2
// fills the 64 x 54 sticker tile from the columns, phase 3 stores the tile coalesced.
constexpr int kDec3Tile = 64;

template <typename T>
__device__ __forceinline__ int first_one_24(const T* row)
{
    constexpr int V = 16 / (int)sizeof(T);                  // elements per 16-byte vector: 8 / 4 / 16
    int col = 0;                                           // argmax semantics: first 1, else column 0
    if (sizeof(T) == 1) {
        const uint2* p = reinterpret_cast<const uint2*>(row);   // 24-byte rows: 8-byte aligned
#pragma unroll
        for (int j = 2; j >= 0; --j) {
            const uint2 v = __ldcs(p + j);
            const uint32_t w[2] = {v.x, v.y};
#pragma unroll
            for (int k = 7; k >= 0; --k) if (((w[k >> 2] >> (8 * (k & 3))) & 0xffu) == 1u) col = 8 * j + k;
        }
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(row);   // 48 / 96-byte rows: 16-byte aligned
#pragma unroll
        for (int j = 24 / V - 1; j >= 0; --j) {
            const uint4 v = __ldcs(p + j);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = V - 1; k >= 0; --k) {
                const bool one = (sizeof(T) == 2) ? (((w[k >> 1] >> (16 * (k & 1))) & 0xffffu) == 0x3f80u)
                                                  : (w[k] == 0x3f800000u);
                if (one) col = V * j + k;
            }
        }
    }
    return col;
}

template <typename T>
__global__ void __launch_bounds__(256)
decode3_exact_kernel(const T* __restrict__ onehot, long long n, uint8_t* __restrict__ out)
{
    __shared__ uint8_t s_col[kDec3Tile * 20];
    __shared__ __align__(16) uint8_t s_rows[kDec3Tile * 54];
    __shared__ uint32_t s_home[48];                        // [0, 24) corners, [24, 48) edges: colours by column
    __shared__ uint8_t s_slot[54];
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * kDec3Tile;
    const int cnt = (int)((n - base) < (long long)kDec3Tile ? (n - base) : (long long)kDec3Tile);
    if (tid < 24) { s_home[tid] = kCornerHome3x[tid]; s_home[24 + tid] = kEdgeHome3[tid]; }
    if (tid < 54) s_slot[tid] = kStickerSlot3x[tid];
    for (int it = tid; it < cnt * 20; it += 256) s_col[it] = (uint8_t)first_one_24<T>(onehot + (base * 20 + it) * 24);
    __syncthreads();
    for (int it = tid; it < cnt * 54; it += 256) {
        const int r = it / 54, k = it - 54 * r;
        const uint32_t sl = s_slot[k];
        uint32_t colour = sl & 7u;                          // centres: 0x80 | colour
        if (!(sl & 0x80u)) {
            const uint32_t q = sl & 31u, j = sl >> 5;
            const uint32_t col = s_col[r * 20 + q];
            colour = (s_home[(q < 8 ? 0 : 24) + (col < 24 ? col : 0)] >> (8 * j)) & 0xffu;
        }
        s_rows[it] = (uint8_t)colour;
    }
    __syncthreads();
    uint8_t* dst = out + base * 54;                         // 64 * 54 bytes per tile: 16-byte aligned
    const int nbytes = cnt * 54, nvec = nbytes >> 4;
    for (int i = tid; i < nvec; i += 256) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(s_rows)[i];
    for (int i = (nvec << 4) + tid; i < nbytes; i += 256) dst[i] = s_rows[i];
}

}  // namespace

namespace cube {

int launch_decode3_exact(const void* onehot, int dtype, long long n, uint8_t* out, cudaStream_t stream)
{
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n + kDec3Tile - 1) / kDec3Tile);
    if (dtype == 0) decode3_exact_kernel<uint16_t><<<blocks, 256, 0, stream>>>((const uint16_t*)onehot, n, out);
    else if (dtype == 1) decode3_exact_kernel<float><<<blocks, 256, 0, stream>>>((const float*)onehot, n, out);
    else decode3_exact_kernel<uint8_t><<<blocks, 256, 0, stream>>>((const uint8_t*)onehot, n, out);
    return (int)cudaGetLastError();
}

int launch_decode2(const void* onehot, int dtype, long long n, uint8_t* out, cudaStream_t stream)
{
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n + kDec2Tile - 1) / kDec2Tile);
    if (dtype == 0) decode2_kernel<uint16_t><<<blocks, 256, 0, stream>>>((const uint16_t*)onehot, n, out);
    else if (dtype == 1) decode2_kernel<float><<<blocks, 256, 0, stream>>>((const float*)onehot, n, out);
    else decode2_kernel<uint8_t><<<blocks, 256, 0, stream>>>((const uint8_t*)onehot, n, out);
    return (int)cudaGetLastError();
}

}  // namespace cube
