// Host-buffer front end of the fused scramble: the end-to-end path a caller with NumPy /
// pinned host arrays uses (bench.py's `e2e` figure).  The batch is cut into chunks that
// rotate over a few stages, each with its own stream and device buffers, so the H2D copy of
// chunk c+1, the kernel of chunk c and the D2H copy of chunk c-1 overlap (the two copy
// engines run in opposite directions at the same time).  All device memory is owned by
// the handle; the per-call entry point allocates nothing.
#include <cuda_runtime.h>
#include <new>

#include "../../include/cube_b200.h"
#include "cube_common.cuh"
#include "cube_kernels.h"

struct cube_pipeline {
    int cube_size, depth, n_stages;
    long long chunk;                 // instances per chunk (multiple of 256)
    cudaStream_t stream[4];
    uint8_t* d_moves[4];
    uint8_t* d_states[4];
    uint8_t* d_solved[4];
    float* d_reward[4];
    unsigned long long* d_counters;  // [n_stages][4]
    unsigned long long* h_counters;  // pinned mirror
};

namespace {

void destroy(cube_pipeline* p)
{
    if (!p) return;
    for (int s = 0; s < p->n_stages; ++s) {
        if (p->d_moves[s]) cudaFree(p->d_moves[s]);
        if (p->d_states[s]) cudaFree(p->d_states[s]);
        if (p->d_solved[s]) cudaFree(p->d_solved[s]);
        if (p->d_reward[s]) cudaFree(p->d_reward[s]);
        if (p->stream[s]) cudaStreamDestroy(p->stream[s]);
    }
    if (p->d_counters) cudaFree(p->d_counters);
    if (p->h_counters) cudaFreeHost(p->h_counters);
    delete p;
}

}  // namespace

extern "C" {

int cube_pipeline_create(int cube_size, int depth, int64_t chunk_instances, int n_stages, cube_pipeline** out)
{
    if (cube_size != 2 && cube_size != 3) return CUBE_ERR_SIZE;
    if (!out || depth < 0 || chunk_instances < 1 || n_stages < 1 || n_stages > 4) return CUBE_ERR_ARG;
    cube_pipeline* p = new (std::nothrow) cube_pipeline();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->cube_size = cube_size;
    p->depth = depth;
    p->n_stages = n_stages;
    p->chunk = (chunk_instances + 255) / 256 * 256;
    const size_t S = cube_size == 3 ? 54 : 24;
    cudaError_t e = cudaSuccess;
    for (int s = 0; s < n_stages && e == cudaSuccess; ++s) {
        e = cudaStreamCreateWithFlags(&p->stream[s], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&p->d_moves[s], (size_t)p->chunk * (depth > 0 ? depth : 1));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_states[s], (size_t)p->chunk * S);
        if (e == cudaSuccess) e = cudaMalloc(&p->d_solved[s], (size_t)p->chunk);
        if (e == cudaSuccess) e = cudaMalloc(&p->d_reward[s], (size_t)p->chunk * sizeof(float));
    }
    if (e == cudaSuccess) e = cudaMalloc(&p->d_counters, sizeof(unsigned long long) * 4 * n_stages);
    if (e == cudaSuccess) e = cudaMallocHost(&p->h_counters, sizeof(unsigned long long) * 4 * n_stages);
    if (e != cudaSuccess) {
        destroy(p);
        return (int)e;
    }
    *out = p;
    return 0;
}

int cube_pipeline_destroy(cube_pipeline* p)
{
    destroy(p);
    return 0;
}

// moves_host [n, depth]; states_out_host [n, S]; solved_host [n] / reward_host [n] may be NULL.
// Blocks until the host buffers are filled.  Host buffers should be page-locked for the
// copies to overlap; pageable memory works but serialises.
int cube_pipeline_scramble_host(cube_pipeline* p, const uint8_t* moves_host, int64_t n, uint8_t* states_out_host,
                                uint8_t* solved_host, float* reward_host, int64_t* solved_count)
{
    if (!p || n < 0 || (n > 0 && (!states_out_host || (p->depth > 0 && !moves_host)))) return CUBE_ERR_ARG;
    const size_t S = p->cube_size == 3 ? 54 : 24;
    cudaError_t e = cudaMemsetAsync(p->d_counters, 0, sizeof(unsigned long long) * 4 * p->n_stages, p->stream[0]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream[0]);
    int rc = 0;
    long long c = 0;
    for (long long off = 0; off < n && e == cudaSuccess && rc == 0; off += p->chunk, ++c) {
        const int s = (int)(c % p->n_stages);
        const long long cnt = (n - off) < p->chunk ? (n - off) : p->chunk;
        cudaStream_t st = p->stream[s];
        if (p->depth > 0)
            e = cudaMemcpyAsync(p->d_moves[s], moves_host + off * p->depth, (size_t)cnt * p->depth,
                                cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        rc = cube::launch_scramble(p->cube_size, p->d_moves[s], cnt, p->depth, p->d_states[s], p->d_solved[s],
                                   p->d_reward[s], p->d_counters + 4 * s, st);
        if (rc) break;
        e = cudaMemcpyAsync(states_out_host + off * S, p->d_states[s], (size_t)cnt * S, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && solved_host)
            e = cudaMemcpyAsync(solved_host + off, p->d_solved[s], (size_t)cnt, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && reward_host)
            e = cudaMemcpyAsync(reward_host + off, p->d_reward[s], (size_t)cnt * sizeof(float),
                                cudaMemcpyDeviceToHost, st);
    }
    for (int s = 0; s < p->n_stages; ++s) {
        cudaError_t e2 = cudaStreamSynchronize(p->stream[s]);
        if (e == cudaSuccess) e = e2;
    }
    if (rc) return rc;
    if (e != cudaSuccess) return (int)e;
    if (solved_count) {
        e = cudaMemcpy(p->h_counters, p->d_counters, sizeof(unsigned long long) * 4 * p->n_stages,
                       cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return (int)e;
        long long total = 0;
        for (int s = 0; s < p->n_stages; ++s) total += (long long)p->h_counters[4 * s];
        *solved_count = total;
    }
    return 0;
}

}  // extern "C"
