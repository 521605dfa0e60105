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

template <typename T>
__global__ void __launch_bounds__(128)
decode2_kernel(const T* __restrict__ onehot, long long n, uint8_t* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* row = onehot + i * 147;
    uint8_t s[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) s[k] = 0;
    s[14] = 3; s[18] = 4; s[23] = 5;                       // the fixed DBL cubie
    for (int cubelet = 0; cubelet < 7; ++cubelet) {
        int col = 0;                                       // argmax semantics: first 1, else column 0
        for (int c = 20; c >= 0; --c) if (is_one<T>(row[cubelet * 21 + c])) col = c;
        const int position = col / 3, ori = col - 3 * position;
#pragma unroll
        for (int k = 0; k < 3; ++k)                        // np.roll(home, ori)[k] == home[(k - ori) mod 3]
            s[kPieceDefs2[position * 3 + k]] = kHomeColour2[cubelet * 3 + (k + 3 - ori) % 3];
    }
    uint8_t* o = out + i * 24;
    for (int k = 0; k < 24; ++k) o[k] = s[k];
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
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (dtype == 0) decode2_kernel<uint16_t><<<blocks, 128, 0, stream>>>((const uint16_t*)onehot, n, out);
    else if (dtype == 1) decode2_kernel<float><<<blocks, 128, 0, stream>>>((const float*)onehot, n, out);
    else decode2_kernel<uint8_t><<<blocks, 128, 0, stream>>>((const uint8_t*)onehot, n, out);
    return (int)cudaGetLastError();
}

}  // namespace cube
