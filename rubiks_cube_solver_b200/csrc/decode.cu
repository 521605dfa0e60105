// One-hot -> sticker rows for 2x2x2: CubeEnv.state_to_sim_state (cube_env.py:154-175), i.e.
// np.where(row == 1)[0][0] per cubelet row, then py222 getStickers.  The reference raises
// NotImplementedError for 3x3x3 (cube_env.py:171-172) -- its corner encoding is lossy -- and so
// does this library.  One thread per instance; a latency-trivial helper, not a hot kernel.
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

}  // namespace

namespace cube {

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
