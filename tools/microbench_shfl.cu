// Micro-benchmark: do warp shuffles (SHFL.IDX with a data-dependent source lane) share the
// shared-memory data pipe?  Times 8 table lookups per iteration done as (a) 8 LDS, (b) 8 SHFL,
// (c) 4 LDS + 4 SHFL, at 8 CTAs x 256 threads per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int N_LDS, int N_SHFL>
__global__ void __launch_bounds__(256) bench(const uint32_t* idx_in, uint32_t* out, long long* cycles, int iters)
{
    __shared__ uint32_t tbl[16 * 16];
    for (int i = threadIdx.x; i < 256; i += 256) tbl[i] = i * 2654435761u;
    __syncthreads();
    uint32_t m = idx_in[blockIdx.x * 256 + threadIdx.x];
    uint32_t r[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) r[u] = tbl[u * 16 + (threadIdx.x & 15)];
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < N_LDS; ++u) acc ^= tbl[u * 16 + m];
#pragma unroll
        for (int u = 0; u < N_SHFL; ++u) acc ^= __shfl_sync(0xffffffffu, r[u], m);
        m = (m + 5 + (acc & 0)) % 12;
    }
    long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int N_LDS, int N_SHFL>
void run(const char* name, const uint32_t* d_idx, uint32_t* d_out, long long* d_cyc, int blocks)
{
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<N_LDS, N_SHFL><<<blocks, 256>>>(d_idx, d_out, d_cyc, iters);
    cudaEventRecord(e0);
    bench<N_LDS, N_SHFL><<<blocks, 256>>>(d_idx, d_out, d_cyc, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    const double warp_iters_per_sm = 8.0 * 8.0 * iters;
    printf("%-16s %7.3f cycles per warp-iteration per SM (clock64), kernel %.3f ms -> %.3f cycles @1.965GHz\n", name,
           avg / warp_iters_per_sm, ms, ms * 1e-3 * 1.965e9 / warp_iters_per_sm);
    delete[] h;
}

int main()
{
    int blocks = 148 * 8;
    uint32_t* h_idx = new uint32_t[blocks * 256];
    uint32_t s = 12345;
    for (int i = 0; i < blocks * 256; ++i) { s = s * 1664525u + 1013904223u; h_idx[i] = (s >> 16) % 12; }
    uint32_t *d_idx, *d_out; long long* d_cyc;
    cudaMalloc(&d_idx, blocks * 256 * 4); cudaMalloc(&d_out, blocks * 256 * 4); cudaMalloc(&d_cyc, blocks * 8);
    cudaMemcpy(d_idx, h_idx, blocks * 256 * 4, cudaMemcpyHostToDevice);
    run<8, 0>("8 LDS", d_idx, d_out, d_cyc, blocks);
    run<4, 0>("4 LDS", d_idx, d_out, d_cyc, blocks);
    run<0, 8>("8 SHFL", d_idx, d_out, d_cyc, blocks);
    run<0, 4>("4 SHFL", d_idx, d_out, d_cyc, blocks);
    run<4, 4>("4 LDS + 4 SHFL", d_idx, d_out, d_cyc, blocks);
    run<8, 8>("8 LDS + 8 SHFL", d_idx, d_out, d_cyc, blocks);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
