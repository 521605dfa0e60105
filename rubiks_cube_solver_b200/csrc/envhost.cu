// Single-cube host front end: what the reference's per-cube callers do through the drop-in CubeEnv
// (cube_env.py:56-111: reset, step, get_obs -- train.py:155, :186-191, mcts.py:80, test.py:123).
//
// One call = one cube, host buffers in and out, ONE stream synchronisation and NO copy calls: the
// handle owns a page of mapped pinned memory that the kernels read and write directly over PCIe
// (sticker row and moves in; sticker row, solved flag and the uint8 one-hot out).  A step is ONE launch
// of the expansion kernel (all A children of the cube with their verdicts and encodings land in the
// page, the host picks the child the action names); reset and encode are two launches.  The
// torch-level path (upload, two kernels, three downloads with their synchronisations) took ~120 us per
// step; the reference's own NumPy step takes ~25 us on the same host.
#include <cstring>
#include <cuda_runtime.h>
#include <new>

#include "../../include/cube_b200.h"
#include "cube_common.cuh"
#include "cube_kernels.h"

struct cube_env_host {
    int cube_size, max_depth;
    uint8_t* pin;        // mapped pinned page, laid out below
    uint8_t* dev;        // the same page as the device sees it
    int off_moves, off_out, off_solved, off_onehot, bytes;
    int off_children, off_child_onehot, off_child_solved;            // [A, S], [A, D] uint8, [A]
};

namespace {

inline int round16(int x) { return (x + 15) & ~15; }

}  // namespace

extern "C" {

int cube_env_host_create(int cube_size, int max_depth, cube_env_host** out)
{
    if (cube_size != 2 && cube_size != 3) return CUBE_ERR_SIZE;
    if (!out || max_depth < 1 || max_depth > 4096) return CUBE_ERR_ARG;
    cube_env_host* h = new (std::nothrow) cube_env_host();
    if (!h) return (int)cudaErrorMemoryAllocation;
    const int S = cube_size == 3 ? 54 : 24, D = cube_size == 3 ? 480 : 147;
    h->cube_size = cube_size;
    h->max_depth = max_depth;
    h->off_moves = round16(S);
    h->off_out = h->off_moves + round16(max_depth);
    h->off_solved = h->off_out + round16(S);
    h->off_onehot = h->off_solved + 16;
    const int A = cube_size == 3 ? 12 : 6;
    h->off_children = h->off_onehot + round16(D);
    h->off_child_onehot = h->off_children + round16(A * S);
    h->off_child_solved = h->off_child_onehot + round16(A * D);
    h->bytes = h->off_child_solved + 16;
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, (size_t)h->bytes, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) { delete h; return (int)e; }
    h->pin = static_cast<uint8_t*>(p);
    memset(h->pin, 0, (size_t)h->bytes);
    void* d = nullptr;
    e = cudaHostGetDevicePointer(&d, p, 0);
    if (e != cudaSuccess) { cudaFreeHost(p); delete h; return (int)e; }
    h->dev = static_cast<uint8_t*>(d);
    *out = h;
    return CUBE_OK;
}

int cube_env_host_destroy(cube_env_host* h)
{
    if (!h) return CUBE_OK;
    if (h->pin) cudaFreeHost(h->pin);
    delete h;
    return CUBE_OK;
}

// shared tail: encode the row at off_out, wait, hand the results back
static int finish(cube_env_host* h, uint8_t* stickers_out_host, uint8_t* onehot_u8_host, int* solved_host,
                  cudaStream_t stream)
{
    const int S = h->cube_size == 3 ? 54 : 24, D = h->cube_size == 3 ? 480 : 147;
    int rc = 0;
    if (onehot_u8_host)
        rc = cube::launch_expand(h->cube_size, h->dev + h->off_out, 1, nullptr, nullptr, h->dev + h->off_onehot,
                                 CUBE_DTYPE_U8, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    const cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return (int)e;
    if (stickers_out_host) memcpy(stickers_out_host, h->pin + h->off_out, (size_t)S);
    if (onehot_u8_host) memcpy(onehot_u8_host, h->pin + h->off_onehot, (size_t)D);
    if (solved_host) *solved_host = h->pin[h->off_solved] != 0;
    return CUBE_OK;
}

int cube_env_host_step(cube_env_host* h, const uint8_t* stickers_host, int action, uint8_t* stickers_out_host,
                       uint8_t* onehot_u8_host, int* solved_host, void* stream)
{
    if (!h || !stickers_host || action < 0 || action > 255) return CUBE_ERR_ARG;
    const int S = h->cube_size == 3 ? 54 : 24, D = h->cube_size == 3 ? 480 : 147, A = h->cube_size == 3 ? 12 : 6;
    memcpy(h->pin, stickers_host, (size_t)S);
    if (action < A) {
        // one launch: every child of the cube (cube_expand), the host keeps child `action`
        const int rc = cube::launch_expand(h->cube_size, h->dev, 1, h->dev + h->off_children,
                                           onehot_u8_host ? h->dev + h->off_child_onehot : nullptr, nullptr, CUBE_DTYPE_U8,
                                           h->dev + h->off_child_solved, nullptr, nullptr, (cudaStream_t)stream);
        if (rc) return rc;
        const cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
        if (stickers_out_host) memcpy(stickers_out_host, h->pin + h->off_children + action * S, (size_t)S);
        if (onehot_u8_host) memcpy(onehot_u8_host, h->pin + h->off_child_onehot + action * D, (size_t)D);
        if (solved_host) *solved_host = h->pin[h->off_child_solved + action] != 0;
        return CUBE_OK;
    }
    h->pin[h->off_moves] = (uint8_t)action;                           // A..255: the kernels' no-op / unspecified range
    const int rc = cube::launch_walk(h->cube_size, h->dev, h->dev + h->off_moves, 1, 1, h->dev + h->off_out,
                                     h->dev + h->off_solved, nullptr, nullptr, (cudaStream_t)stream);
    if (rc) return rc;
    return finish(h, stickers_out_host, onehot_u8_host, solved_host, (cudaStream_t)stream);
}

int cube_env_host_scramble(cube_env_host* h, const uint8_t* moves_host, int depth, uint8_t* stickers_out_host,
                           uint8_t* onehot_u8_host, int* solved_host, void* stream)
{
    if (!h || depth < 0 || depth > h->max_depth || (depth > 0 && !moves_host)) return CUBE_ERR_ARG;
    if (depth > 0) memcpy(h->pin + h->off_moves, moves_host, (size_t)depth);
    const int rc = cube::launch_scramble(h->cube_size, h->dev + h->off_moves, 1, depth, h->dev + h->off_out,
                                         h->dev + h->off_solved, nullptr, nullptr, (cudaStream_t)stream);
    if (rc) return rc;
    return finish(h, stickers_out_host, onehot_u8_host, solved_host, (cudaStream_t)stream);
}

int cube_env_host_encode(cube_env_host* h, const uint8_t* stickers_host, uint8_t* onehot_u8_host, void* stream)
{
    if (!h || !stickers_host || !onehot_u8_host) return CUBE_ERR_ARG;
    const int S = h->cube_size == 3 ? 54 : 24;
    memcpy(h->pin + h->off_out, stickers_host, (size_t)S);
    return finish(h, nullptr, onehot_u8_host, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
