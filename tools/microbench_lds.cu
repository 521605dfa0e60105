// Micro-benchmark: cost of shared-memory table lookups where the 32 lanes of a warp hit only
// 12 distinct entries (the move-table access pattern of the scramble kernel), for 32-, 64- and
// 128-bit loads and two table layouts.  Prints SM cycles per warp-level load instruction at
// full occupancy, i.e. the reciprocal shared-memory throughput the kernel design can count on.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_lds microbench_lds.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int WIDTH, int STRIDE_WORDS>
__global__ void __launch_bounds__(256) bench(const uint32_t* idx_in, uint32_t* out, long long* cycles, int iters)
{
    __shared__ __align__(16) uint32_t tbl[16 * 64];
    for (int i = threadIdx.x; i < 16 * 64; i += 256) tbl[i] = i * 2654435761u;
    __syncthreads();
    uint32_t m = idx_in[blockIdx.x * 256 + threadIdx.x];      // 0..11
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t* p = tbl + m * STRIDE_WORDS + u * (WIDTH / 32) ;
            if (WIDTH == 32) acc ^= *p;
            if (WIDTH == 64) { uint2 v = *reinterpret_cast<const uint2*>(p); acc ^= v.x + v.y; }
            if (WIDTH == 128) { uint4 v = *reinterpret_cast<const uint4*>(p); acc ^= v.x + v.y + v.z + v.w; }
        }
        m = (m + 5 + (acc & 0)) % 12;
    }
    long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int WIDTH, int STRIDE_WORDS>
void run(const char* name, const uint32_t* d_idx, uint32_t* d_out, long long* d_cyc, int blocks)
{
    const int iters = 2000;
    bench<WIDTH, STRIDE_WORDS><<<blocks, 256>>>(d_idx, d_out, d_cyc, iters);
    bench<WIDTH, STRIDE_WORDS><<<blocks, 256>>>(d_idx, d_out, d_cyc, iters);
    cudaDeviceSynchronize();
    long long* h = new long long[blocks];
    cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    // 8 resident CTAs/SM x 8 warps issue iters*8 loads each over `avg` cycles
    const double per_sm_loads = 8.0 * 8.0 * iters * 8.0;
    printf("%-28s %8.3f cycles per warp-load per SM (%.0f cycles)\n", name, avg / per_sm_loads, avg);
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
    run<32, 1>("LDS.32  stride 1 word", d_idx, d_out, d_cyc, blocks);   // [word u][move]: u*? -> here p = tbl+m+u: same-bank-ish
    run<32, 16>("LDS.32  stride 16 words", d_idx, d_out, d_cyc, blocks);
    run<32, 17>("LDS.32  stride 17 words", d_idx, d_out, d_cyc, blocks);
    run<64, 2>("LDS.64  stride 2 words", d_idx, d_out, d_cyc, blocks);
    run<64, 16>("LDS.64  stride 16 words", d_idx, d_out, d_cyc, blocks);
    run<64, 18>("LDS.64  stride 18 words", d_idx, d_out, d_cyc, blocks);
    run<128, 4>("LDS.128 stride 4 words", d_idx, d_out, d_cyc, blocks);
    run<128, 32>("LDS.128 stride 32 words", d_idx, d_out, d_cyc, blocks);
    run<128, 36>("LDS.128 stride 36 words", d_idx, d_out, d_cyc, blocks);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
